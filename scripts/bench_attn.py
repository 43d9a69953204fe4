"""Times the attention kernel at the bench shape (S=9728, 24 heads) with CUDA events (not under a profiler)."""
import os, sys, json
import torch
sys.path.insert(0, ".")
from unitex_b200 import ops

S, H = int(sys.argv[1]) if len(sys.argv) > 1 else 9728, 24
torch.manual_seed(0)
qkv = torch.randn(S, 3 * H * 128, device="cuda").to(torch.bfloat16)
out = torch.empty(S, H * 128, device="cuda", dtype=torch.bfloat16)
flops = 4.0 * H * 128 * S * S
res = {}
for impl in ("2",):      # attention2_kernel (the two round-1 alternatives left the library)
    try:
        for _ in range(3):
            ops.attention(qkv, H, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention(qkv, H, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[impl] = {"ms": ms, "tflops": flops / ms / 1e9}
    except Exception as e:
        res[impl] = {"error": str(e)[:200]}
print(json.dumps(res))
