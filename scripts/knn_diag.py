import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from oracle import bake as ob
from unitex_b200.bake import knn
g = torch.Generator().manual_seed(2)
src = torch.rand(6000, 3, generator=g)
dst = torch.rand(1500, 3, generator=g)
score, index = knn(src, dst, k=2)
torch.cuda.synchronize()
rd, ri = ob.nearest_k(src, dst, 2)
s = score.cpu().numpy(); r = rd.numpy()
bad = np.argwhere(s != r)
print("index equal", bool((index.cpu() == ri).all()), "score mismatches", len(bad), "of", s.size)
for (i, j) in bad[:6]:
    e = int(ri[i, j])
    d = src[e].numpy() - dst[i].numpy()
    f = np.float32
    d2 = f(f(f(d[0] * d[0]) + f(d[1] * d[1])) + f(d[2] * d[2]))
    d2b = f(f(d[0] * d[0]) + f(f(d[1] * d[1]) + f(d[2] * d[2])))
    d2fma = f(np.float64(d[0]) * d[0] + np.float64(d[1]) * d[1] + np.float64(d[2]) * d[2])
    h = lambda x: float(x).hex()
    print(i, j, "gpu", h(s[i, j]), "oracle", h(r[i, j]), "sqrt(d2)", h(np.sqrt(d2)), "sqrt(d2b)", h(np.sqrt(d2b)), "sqrt(d2 wide)", h(np.sqrt(d2fma)),
          "gpu^2", h(f(s[i, j]) * f(s[i, j])), "d2", h(d2))
