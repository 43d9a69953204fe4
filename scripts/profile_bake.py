"""Runs the UV bake twice on the bench mesh so ncu can list / capture its kernels (numbers under ncu are not reported)."""
import sys
sys.path.insert(0, ".")
import torch
import bench

out = bench.bench_uv_bake(torch.device("cuda", 0))
print({k: out[k] for k in ("value", "gpu_ms_per_bake", "bvh_build_ms", "mrays_per_s")})
