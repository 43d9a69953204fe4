"""The bake alone on the reference's teaser_robot mesh (staged in the git-ignored gpurun_in/), synthetic analytic view images:
for ncu launch lists / captures on real geometry."""
import os, sys, json
sys.path.insert(0, ".")
import numpy as np, torch
if os.environ.get("UTX_LIB"):
    from pathlib import Path
    from unitex_b200 import _lib
    _lib._LIB_PATH = Path(os.environ["UTX_LIB"])
from unitex_b200 import bake as ub
mesh = "gpurun_in/teaser_inputmesh.obj"
if not os.path.exists(mesh):
    print("skipped: fixture not staged"); sys.exit(0)
V, F, UV, Ft = ub.load_mesh(mesh)
V = np.asarray(V, np.float64); lo, hi = V.min(0), V.max(0); s = (hi - lo).max() / 1.9
V = (V / s - (lo + hi) / (2 * s)).astype(np.float32)
m = ub.BakeMesh(V, F, UV * 2 - 1, Ft)
r = ub.NVDiffRendererInverse(pbr_mesh=m)
c2ws, intr = ub.generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], ub.generate_intrinsics(1.0, 1.0, fov=False)
mats = torch.matmul(ub.intr_to_proj(intr, perspective=False), ub.c2w_to_w2c(c2ws)).cuda()
rast = ub.rasterize(ub.transform_points(m.vertices, mats), m.faces, (512, 512))
pos = ub.interpolate(m.vertices, rast, m.faces)
img = (0.5 + 0.4 * torch.sin(3.0 * pos + 0.3)) * (rast[..., 3:4] > 0)
kw = dict(H=512, W=512, H2D=2048, W2D=2048, perspective=False, ray_normal_angle_threhold=100.0, method="reproject", filt_gradient_points=False)
for _ in range(3):
    out = r.infer(m, c2ws, intr, img, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = r.infer(m, c2ws, intr, img, **kw)
e1.record(); torch.cuda.synchronize()
import hashlib
h = hashlib.sha256()
for t_ in (out[1], out[2], out[3], r.last_nn_index):
    h.update(t_.contiguous().cpu().numpy().tobytes())
print(json.dumps({"sha": h.hexdigest()[:16], "bake_ms": e0.elapsed_time(e1) / 5, "covered": int(out[2].sum()), "visible": int(out[1].any(dim=0).sum())}))
