"""BASELINE config 4 on the GPU box: the reference's own test mesh (teaser_robot, from the committed lossless fixture
tests/golden/teaser_robot.npz.xz, written out as an OBJ) and a synthetic reference image through CustomRGBTextureFullPipeline
with the FULL-SIZE random-init FLUX (19+38 blocks) and the reference's call shapes (6 views, 512 x 3072 strip, 28 steps per
call, 2048^2 atlas).  Prints one JSON line of stage timings."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from PIL import Image

import warnings
warnings.simplefilter("ignore")
from tests.bake_meshes import teaser_robot_raw
from unitex_b200.export import save_obj
import tempfile
_tmp = tempfile.mkdtemp()            # (the 55 MB OBJ must not land in gpurun_out/: that directory is copied back, 64 MiB at most)
mesh, image = os.path.join(_tmp, "inputmesh.obj"), os.path.join(_tmp, "image.png")
save_obj(mesh, *[teaser_robot_raw()[i] for i in (0, 1, 2, 3)])
yy, xx = np.mgrid[0:1024, 0:1024]
Image.fromarray(np.stack([(xx // 4) % 256, (yy // 4) % 256, ((xx + yy) // 8) % 256], -1).astype(np.uint8)).save(image)
import pipeline as P

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

t0 = sync()
pipe = P.CustomRGBTextureFullPipeline(pretrain_models="random", super_resolutions=False, seed=63)
t1 = sync()
stages = {"build_pipeline_random_weights_s": t1 - t0}
save_dir = os.path.join(_tmp, "run")
cache = os.path.join(save_dir, "cache")
os.makedirs(cache, exist_ok=True)
for name, fn in (("preprocess_blank_mesh", lambda: pipe.preprocess_blank_mesh(cache, mesh)),
                 ("preprocess_reference_image", lambda: pipe.preprocess_reference_image(cache, image)),
                 ("render_geometry_images", lambda: pipe.render_geometry_images(cache, os.path.join(cache, "processed_mesh.obj"))),
                 ("infer_mv", lambda: pipe.infer_mv(cache, os.path.join(cache, "processed_image.png"), os.path.join(cache, "mv_normal.png"), os.path.join(cache, "mv_ccm.png"))),
                 ("reproject_and_query_field", lambda: pipe.reproject_and_query_field(os.path.join(cache, "wo_LTM"), os.path.join(cache, "processed_mesh.obj"), os.path.join(cache, "mv_rgb.png"), os.path.join(cache, "camera_info.pth")))):
    a = sync(); fn(); b = sync()
    stages[name + "_s"] = b - a
# the bake alone, mesh resident (metric 2 on the real mesh)
from unitex_b200 import bake as ub
r = ub.NVDiffRendererInverse(device="cuda").update_from_file(os.path.join(cache, "processed_mesh.obj"))
cam = torch.load(os.path.join(cache, "camera_info.pth"), weights_only=True, map_location="cuda")
img = torch.from_numpy(np.asarray(Image.open(os.path.join(cache, "mv_rgb.png")).convert("RGB")).astype(np.float32) / 255.0).cuda()
attrs = img.reshape(2, 512, 3, 512, 3).permute(0, 2, 1, 3, 4).reshape(6, 512, 512, 3)
a = sync(); r.pbr_mesh.optix; b = sync()
stages["lbvh_build_ms"] = (b - a) * 1e3
kw = dict(c2ws=cam["c2ws"].cpu(), intrinsics=cam["intrinsics"].cpu(), image_attrs=attrs, perspective=False, H=512, W=512, H2D=2048, W2D=2048,
          method="reproject", ray_normal_angle_threhold=100, filt_gradient_points=False)
for _ in range(2):
    out = r.infer(r.pbr_mesh, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = r.infer(r.pbr_mesh, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
_, vis, m2, col = out
V, F = r.pbr_mesh.vertices.shape[0], r.pbr_mesh.faces.shape[0]
res = {"mesh": {"V": V, "F": F}, "stages": stages, "bake_ms": ms, "bake_mpix_per_s": 2048 * 2048 / 1e6 / (ms * 1e-3),
       "covered_texels": int(m2.sum()), "visible_texels": int(vis.any(dim=0).sum()), "rays": int(6 * m2.sum()),
       "alpha_coverage": float((np.asarray(Image.open(os.path.join(cache, "mv_alpha.png"))) > 0).mean()),
       "color_finite": bool(torch.isfinite(col).all()), "glb_bytes": os.path.getsize(os.path.join(cache, "wo_LTM", "textured_mesh.glb"))}
print(json.dumps(res))
