"""CPU: host-side wire formats and grid re-ordering around the hot path (reference pipeline.py:239-244, 280-282;
io/link_pbr_to_mesh.py:9-31)."""
import json
import struct

import numpy as np

from tests.bake_meshes import two_spheres


def test_view_grid_reordering_roundtrip():
    from unitex_b200.export import control_grid_to_strip, strip_to_view_grid
    g = np.random.default_rng(0).integers(0, 255, (1024, 1536, 3), dtype=np.uint8)
    strip = control_grid_to_strip(g, g)
    assert strip.shape == (512, 3072, 3)
    tiles = g.reshape(2, 512, 3, 512, 3).transpose(0, 2, 1, 3, 4).reshape(6, 512, 512, 3)
    # strip order f,l,r,b,t,d = grid tiles [0,4,1,3,2,5]; tile 5 rotated by 180 degrees   
    assert np.array_equal(strip[:, 512:1024], tiles[4]) and np.array_equal(strip[:, 2048:2560], tiles[2])
    assert np.array_equal(strip[:, 2560:], tiles[5][::-1, ::-1])
    assert np.array_equal(strip_to_view_grid(strip), g)


def test_obj_and_glb_writers(tmp_path):
    from unitex_b200.bake import load_obj
    from unitex_b200.export import save_glb, save_obj
    v, f, uv, fuv = two_spheres(6, 8)
    uv01 = (uv + 1) / 2
    p = tmp_path / "m.obj"
    save_obj(str(p), v, f, uv01, fuv)
    V, F, UV, Ft = load_obj(str(p))
    assert np.allclose(V, v, atol=1e-6) and np.array_equal(F, f) and np.allclose(UV, uv01, atol=1e-6) and np.array_equal(Ft, fuv)
    tex = np.random.default_rng(1).integers(0, 255, (32, 32, 3), dtype=np.uint8)
    g = tmp_path / "m.glb"
    save_glb(str(g), v, f, uv01, fuv, tex)
    raw = g.read_bytes()
    magic, ver, total = struct.unpack("<III", raw[:12])
    assert magic == 0x46546C67 and ver == 2 and total == len(raw)
    jl, jt = struct.unpack("<II", raw[12:20])
    doc = json.loads(raw[20:20 + jl])
    assert jt == 0x4E4F534A and doc["accessors"][2]["count"] == f.size and doc["images"][0]["mimeType"] == "image/png"
    bl, bt = struct.unpack("<II", raw[20 + jl:28 + jl])
    assert bt == 0x004E4942 and 28 + jl + bl == len(raw)


def test_lens_blur_kernel_sums_to_one():
    from unitex_b200.bake import lens_blur_kernel_2d
    K = lens_blur_kernel_2d()
    assert K.shape == (7, 7) and abs(K.sum() - 1.0) < 1e-5 and np.allclose(K, K.T, atol=1e-7)


def test_preprocess_blank_mesh_normalises_bbox(tmp_path):
    """reference pipeline.py:170-179 -> uv_atlas.py:131-147: bbox centred, longest side = 2 * 0.95."""
    import types
    import pipeline as drop_in
    from tests.bake_meshes import two_spheres
    from unitex_b200 import bake as ub
    from unitex_b200.export import save_obj
    v, f, uv, fuv = two_spheres(6, 12)
    src = str(tmp_path / "in.obj")
    save_obj(src, v * 3.7 + np.array([5.0, -2.0, 1.0], np.float32), f, (uv + 1) / 2, fuv)
    drop_in.CustomRGBTextureFullPipeline.preprocess_blank_mesh(types.SimpleNamespace(), str(tmp_path), src)
    V, F, UV, Ft = ub.load_obj(str(tmp_path / "processed_mesh.obj"))
    V = np.asarray(V)
    lo, hi = V.min(0), V.max(0)
    assert np.allclose((lo + hi) / 2, 0.0, atol=1e-6) and abs((hi - lo).max() - 1.9) < 1e-6
    assert np.array_equal(np.asarray(F), f) and len(UV) == len(uv)


def test_glb_reader_roundtrip_and_node_transform(tmp_path):
    """load_glb inverts save_glb (positions, faces, UVs back in the OBJ convention) and applies node transforms."""
    import struct as st
    from unitex_b200 import bake as ub
    from unitex_b200.export import save_glb
    v, f, uv, fuv = two_spheres(6, 12)
    uv01 = (uv + 1) / 2
    p = str(tmp_path / "m.glb")
    save_glb(p, v, f, uv01, fuv, np.zeros((8, 8, 3), np.uint8))
    V, F, UV, Ft = ub.load_mesh(p)
    assert F.shape == f.shape and np.array_equal(F, Ft)
    assert np.allclose(V[F], v[f]) and np.allclose(UV[Ft], uv01[fuv], atol=1e-6)
    # same file with a +90 degree rotation about x on the node (as the reference's gamda_style GLBs carry): y -> z, z -> -y
    blob = open(p, "rb").read()
    jl = st.unpack_from("<I", blob, 12)[0]
    g = json.loads(blob[20:20 + jl])
    g["nodes"][0]["rotation"] = [0.7071067811865476, 0.0, 0.0, 0.7071067811865476]
    js = json.dumps(g, separators=(",", ":")).encode()
    js += b" " * ((-len(js)) % 4)
    rest = blob[20 + jl:]
    with open(p, "wb") as fh:
        fh.write(st.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + len(rest)) + st.pack("<II", len(js), 0x4E4F534A) + js + rest)
    V2, F2, _, _ = ub.load_glb(p)
    want = np.stack([v[:, 0], -v[:, 2], v[:, 1]], -1)
    assert np.allclose(V2[F2], want[f], atol=1e-6)


def test_obj_reader_bulk_and_fallback_paths(tmp_path):
    from unitex_b200 import bake as ub
    from unitex_b200.export import save_obj
    v, f, uv, fuv = two_spheres(4, 8)
    p = str(tmp_path / "tri.obj")
    save_obj(p, v, f, (uv + 1) / 2, fuv)
    V, F, UV, Ft = ub.load_obj(p)                                   # bulk path: all triangles, one corner format
    Vs, UVs, Fs, Fts = ub._load_obj_slow(p)
    assert np.array_equal(V, Vs) and np.array_equal(UV, UVs) and np.array_equal(F, Fs - 1) and np.array_equal(Ft, Fts - 1)
    assert np.allclose(V, v, atol=1e-6) and np.array_equal(F, f) and np.array_equal(Ft, fuv)
    q = str(tmp_path / "quad.obj")                                  # a quad, v/vt/vn corners, a negative index: line-by-line path
    open(q, "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                       "f 1/1/1 2/2/1 3/3/1 4/4/1\nf -4/1/1 -3/2/1 -2/3/1\n")
    V, F, UV, Ft = ub.load_obj(q)
    assert F.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]] and Ft.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2]]
    n = str(tmp_path / "nouv.obj")                                  # v//vn corners: no UVs
    open(n, "w").write("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n")
    V, F, UV, Ft = ub.load_obj(n)
    assert F.tolist() == [[0, 1, 2]] and len(UV) == 0


def test_per_triangle_atlas_is_a_valid_layout():
    """UV-less meshes (the reference unwraps them with open3d / UVAtlas, uv_atlas.py:149-175 [ext]): every triangle gets its own
    slot; slots do not overlap (checked with the oracle's rasteriser: every covered texel's barycentrics lie inside its own
    triangle and the covered area equals the sum of the triangles' areas), stay inside the unit square and keep their gutters."""
    import numpy as np
    from oracle import bake as ob
    from unitex_b200.uv_atlas import per_triangle_atlas
    n = 1999
    uv, ft = per_triangle_atlas(n, 512)
    assert uv.shape == (3 * n, 2) and ft.shape == (n, 3) and uv.min() > 0 and uv.max() < 1
    t = uv[ft] * 512.0                                                     # [F, 3, 2] texel coordinates
    e1, e2 = t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]
    area = 0.5 * (e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
    assert (area > 1.0).all()                                              # counter-clockwise, more than a texel each
    clip = np.concatenate([uv * 2 - 1, np.zeros((len(uv), 1), np.float32), np.ones((len(uv), 1), np.float32)], -1)[None].astype(np.float32)
    rast = ob.rasterize(clip, ft, 512, 512)
    tid = rast[0, ..., 3].astype(np.int64) - 1
    assert set(np.unique(tid[tid >= 0])) == set(range(n))                  # every face owns at least one texel
    # a texel centre inside two triangles would show up as covered area in excess of the sum of the areas (up to edge texels)
    assert (tid >= 0).sum() <= area.sum() + 2.0 * n
    # gutters: no two different faces on 4-adjacent texels
    a, b = tid[:, :-1], tid[:, 1:]
    assert not ((a >= 0) & (b >= 0) & (a != b)).any()
    a, b = tid[:-1], tid[1:]
    assert not ((a >= 0) & (b >= 0) & (a != b)).any()
    with np.testing.assert_raises(ValueError):
        per_triangle_atlas(400_000, 2048)                                  # 4-texel cells: refuse, the mesh has to be decimated first
