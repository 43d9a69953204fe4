"""The bake alone on the reference's teaser_robot mesh (tests/golden/teaser_robot.npz.xz), synthetic analytic view images:
for ncu launch lists / captures on real geometry.  argv[1] = warm-up bakes (default 3), argv[2] = timed bakes (default 5); under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` run it as `... profile_bake_teaser.py 2 1`
so the CSV holds exactly one bake (add `--profile-from-start off`: cudaProfilerStart/Stop bracket the timed bakes)."""
import os, sys, json
sys.path.insert(0, ".")
import numpy as np, torch
if os.environ.get("UTX_LIB"):
    from pathlib import Path
    from unitex_b200 import _lib
    _lib._LIB_PATH = Path(os.environ["UTX_LIB"])
from unitex_b200 import bake as ub
from tests.bake_meshes import teaser_robot
V, F, UV2, Ft = teaser_robot()
m = ub.BakeMesh(V, F, UV2, Ft)
n_warm = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 5
r = ub.NVDiffRendererInverse(pbr_mesh=m)
c2ws, intr = ub.generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], ub.generate_intrinsics(1.0, 1.0, fov=False)
mats = torch.matmul(ub.intr_to_proj(intr, perspective=False), ub.c2w_to_w2c(c2ws)).cuda()
rast = ub.rasterize(ub.transform_points(m.vertices, mats), m.faces, (512, 512))
pos = ub.interpolate(m.vertices, rast, m.faces)
img = (0.5 + 0.4 * torch.sin(3.0 * pos + 0.3)) * (rast[..., 3:4] > 0)
kw = dict(H=512, W=512, H2D=2048, W2D=2048, perspective=False, ray_normal_angle_threhold=100.0, method="reproject", filt_gradient_points=False)
for _ in range(n_warm):
    out = r.infer(m, c2ws, intr, img, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()            # `ncu --profile-from-start off`: only the timed bakes are captured
e0.record()
for _ in range(n_time):
    out = r.infer(m, c2ws, intr, img, **kw)
e1.record(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
import hashlib
h = hashlib.sha256()
for t_ in (out[1], out[2], out[3], r.last_nn_index):
    h.update(t_.contiguous().cpu().numpy().tobytes())
print(json.dumps({"sha": h.hexdigest()[:16], "bake_ms": e0.elapsed_time(e1) / n_time, "covered": int(out[2].sum()), "visible": int(out[1].any(dim=0).sum())}))
