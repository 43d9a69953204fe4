"""One timing + one checksum of the bench bake (bench.bench_uv_bake's mesh) for A/B runs of kernel variants selected by
environment variables: identical checksums across variants = bit-identical masks, indices and colours."""
import hashlib
import json
import sys
sys.path.insert(0, ".")
import os
import torch
import bench
if os.environ.get("UTX_LIB"):      # A/B against an alternative build of the library
    from pathlib import Path
    from unitex_b200 import _lib
    _lib._LIB_PATH = Path(os.environ["UTX_LIB"])

dev = torch.device("cuda", 0)
out = bench.bench_uv_bake(dev, return_tensors=True, mesh_name=sys.argv[1] if len(sys.argv) > 1 else "teaser_robot")
vis, m2, col, nn = out.pop("tensors")
h = hashlib.sha256()
for t in (vis, m2, col, nn):
    h.update(t.contiguous().cpu().numpy().tobytes())
print(json.dumps({"gpu_ms": round(out["gpu_ms_per_bake"], 3), "wall_ms": round(out["ms_per_bake"], 3), "mpix": round(out["value"], 1),
                  "sha": h.hexdigest()[:16], "visible": out["config"]["visible_texels"]}))
