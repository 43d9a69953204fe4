"""Times the 2-CTA GEMM kernel at the DiT shapes with CUDA events (not under a profiler)."""
import os, sys, json
import torch
sys.path.insert(0, ".")
if os.environ.get("UTX_LIB"):      # A/B against an alternative build of the library
    from pathlib import Path
    from unitex_b200 import _lib
    _lib._LIB_PATH = Path(os.environ["UTX_LIB"])
from unitex_b200 import ops

torch.manual_seed(0)
shapes = [(9728, 9216, 3072), (9216, 3072, 12288), (9728, 21504, 3072), (9728, 3072, 15360), (9216, 12288, 3072)]
res = {}
for M, N, K in shapes:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for impl in ("2",):
        os.environ["UTX_GEMM_IMPL"] = impl
        try:
            for _ in range(3):
                ops.gemm(A, W, b, out=C)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                ops.gemm(A, W, b, out=C)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 50
            res[f"{M}x{N}x{K}/impl{impl}"] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
        except Exception as e:
            res[f"{M}x{N}x{K}/impl{impl}"] = {"error": str(e)[:200]}
    del A, W, C
print(json.dumps(res, indent=1))
