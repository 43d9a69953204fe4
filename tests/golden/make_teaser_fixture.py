"""Writes tests/golden/teaser_robot.npz.xz from the reference's own test case
(`/root/reference/test_cases/teaser_robot/inputmesh.obj`: V 269 026, F 499 981, 268 818 UVs -- BASELINE.json config 4).

Category-(b) test DATA, not source: the geometry is stored losslessly (float32 bit patterns of the parsed OBJ, faces as
deltas, UV faces as their difference from the position faces) inside an xz-compressed npz, ~3 MB.  The GPU box has no
/root/reference, so the -m gpu tests and bench.py read this file (tests/bake_meshes.py::teaser_robot).

    python tests/golden/make_teaser_fixture.py
"""
import io
import lzma
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
SRC = "/root/reference/test_cases/teaser_robot/inputmesh.obj"
DST = Path(__file__).resolve().parent / "teaser_robot.npz.xz"


def main():
    from unitex_b200.bake import load_obj
    V, F, UV, Ft = load_obj(SRC)
    assert V.shape == (269026, 3) and F.shape == (499981, 3) and Ft.shape == F.shape
    buf = io.BytesIO()
    np.savez(buf, V=V.view(np.uint32), UV=UV.view(np.uint32),
             F_delta=np.diff(F.reshape(-1).astype(np.int64), prepend=0).astype(np.int32),
             Ft_minus_F=(Ft.astype(np.int64) - F).astype(np.int32))
    DST.write_bytes(lzma.compress(buf.getvalue(), preset=9 | lzma.PRESET_EXTREME))
    print(DST, DST.stat().st_size, "bytes")
    from tests.bake_meshes import teaser_robot_raw
    V2, F2, UV2, Ft2 = teaser_robot_raw()
    assert np.array_equal(V2.view(np.uint32), V.view(np.uint32)) and np.array_equal(F2, F)
    assert np.array_equal(UV2.view(np.uint32), UV.view(np.uint32)) and np.array_equal(Ft2, Ft)
    print("round trip ok")


if __name__ == "__main__":
    main()
