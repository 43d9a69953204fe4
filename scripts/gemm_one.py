"""One GEMM shape launched a few times (for ncu metric runs): python scripts/gemm_one.py M N K"""
import sys
import torch
sys.path.insert(0, ".")
from unitex_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(A, W, b, out=C)
torch.cuda.synchronize()
