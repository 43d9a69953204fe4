"""Launches each hot kernel once at the bench shapes (S=9728, D=3072, 24 heads) so `ncu --set full -k regex:...` can
capture it without the 24 GB model around it.  Not a benchmark: numbers printed under ncu are never reported."""
import sys
import torch
sys.path.insert(0, ".")
from unitex_b200 import ops

torch.manual_seed(0)
S, D, H = 9728, 3072, 24
dev = "cuda"
bf = torch.bfloat16
x = torch.randn(S, D, device=dev).to(bf)
w_qkv = (torch.randn(3 * D, D, device=dev) * 0.02).to(bf)
b_qkv = torch.zeros(3 * D, device=dev, dtype=bf)
w_ff2 = (torch.randn(D, 4 * D, device=dev) * 0.02).to(bf)
hbuf = torch.randn(S, 4 * D, device=dev).to(bf)
gate = torch.randn(D, device=dev)
shift, scale = torch.randn(D, device=dev) * 0.1, torch.randn(D, device=dev) * 0.1
ids = torch.zeros(S, 3, device=dev)
ids[:, 1] = torch.arange(S, device=dev) % 96
ids[:, 2] = torch.arange(S, device=dev) // 96
cos, sin = ops.rope_table(ids)
rw = torch.ones(128, device=dev, dtype=bf)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    xn = ops.ln_modulate(x, shift, scale)
    qkv = ops.gemm(xn, w_qkv, b_qkv)
    ops.rmsnorm_rope_(qkv, H, rw, rw, cos, sin)
    o = ops.attention(qkv, H)
    res = x.clone()
    ops.gemm(hbuf, w_ff2, None, epi=ops.EPI_GATE_RES, gate=gate, res=res, out=res)
torch.cuda.synchronize()
print("ok")
