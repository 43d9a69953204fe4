"""CPU oracle for the UniTEX hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import anything from this package, and only as the
checker / the reported CPU baseline.  The product path (`unitex_b200/`,
`flux_piplines/`, `pipeline.py`) never imports it and fails loudly when the CUDA
library is missing.

PARITY UNPINNED: the reference (YixunLiang/UniTEX) ships no golden vectors or
known-answer tests for this path, and its arithmetic lives in third-party
packages that are absent from /root/reference and from this image (diffusers
[unpinned, >=0.32], peft==0.15.2, nvdiffrast@729261dc, slangtorch==1.3.7,
torch_kdtree@86961f7d).  Each function below restates the published algorithm
and cites the reference call site (file:line under /root/reference) it follows;
the analytic known-answer tests in tests/ (SDPA in fp32, closed-form RoPE,
closed-form sigma schedule, ray/triangle hits, constant-image pull-push ...) are
what pins it.
"""
