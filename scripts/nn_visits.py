"""Diagnosis: how many points does the exact 1-NN walk score per query on the bench bake?  Needs the debug build
(-DUTX_NN_DEBUG, scripts/_dbg/libunitex_dbg.so) and UTX_NN_COUNT=1."""
import os, sys, json
sys.path.insert(0, ".")
os.environ["UTX_NN_COUNT"] = "1"
from pathlib import Path
from unitex_b200 import _lib
_lib._LIB_PATH = Path("scripts/_dbg/libunitex_dbg.so")
import numpy as np, torch
import bench
out = bench.bench_uv_bake(torch.device("cuda", 0), return_tensors=True)
vis, m2, col, nn = out.pop("tensors")
c = nn.cpu().numpy().astype(np.int64)
covered = m2.cpu().numpy().reshape(-1) > 0
seen = vis.any(dim=0).cpu().numpy().reshape(-1)
q = covered & ~seen
v = c[q]
print(json.dumps({"queries": int(q.sum()), "sum": int(v.sum()), "mean": float(v.mean()), "p50": float(np.percentile(v, 50)), "p90": float(np.percentile(v, 90)),
                  "p99": float(np.percentile(v, 99)), "p999": float(np.percentile(v, 99.9)), "max": int(v.max()),
                  "share_top1pct": float(np.sort(v)[-len(v)//100:].sum() / v.sum())}))
H = 2048
yy, xx = np.nonzero(q.reshape(H, H))
big = v > np.percentile(v, 99)
print("top-1% queries bbox (rows, cols):", int(yy[big].min()), int(yy[big].max()), int(xx[big].min()), int(xx[big].max()), "count", int(big.sum()))
hist, edges = np.histogram(np.log2(np.maximum(v, 1)), bins=12)
print("log2 histogram", hist.tolist(), [round(float(e), 1) for e in edges])
