"""CPU oracle for the UniTEX hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import anything from this package, and only as the
checker / the reported CPU baseline.  The product path (`unitex_b200/`,
`flux_piplines/`, `pipeline.py`) never imports it and fails loudly when the CUDA
library is missing.

PARITY: PINNED FOR THE REFERENCE'S OWN CODE, UNPINNED FOR ITS ABSENT PACKAGES.
The reference (YixunLiang/UniTEX) ships no golden vectors or known-answer tests
for this path.  Its in-tree Python IS run here: tests/golden/ref_harness.py
imports the unmodified `NVDiffRendererInverse` (mv_to_pcd, uv_to_pcd, both bake
variants, infer), PBRMesh, PointCloud, knn, pull_push, lens_blur_torch, the
camera code, `PBRFluxPipeline.__call__` (+ prepare_latents_and_image_ids, pack /
unpack / ids, calculate_shift, retrieve_timesteps) and
`NativeFluxAttnProcessor2_0` from /root/reference and writes their outputs to
tests/golden/ref_*.npz; tests/test_reference_golden_cpu.py holds this oracle to
them BIT FOR BIT.  What stays unpinned is the arithmetic of the third-party
packages that are absent from /root/reference and from this image (diffusers
[unpinned, >=0.32] transformer / scheduler / VAE, peft==0.15.2,
nvdiffrast@729261dc, torch_kdtree@86961f7d; slangtorch==1.3.7 cannot compile the
in-tree .slang tracer, which is restated from its sources): in the harness those
are supplied BY this oracle, so for them the analytic known-answer tests in
tests/ (SDPA in fp32, closed-form RoPE and sigma schedule, ray/triangle hits,
raster fill rules, constant-image pull-push ...) are what pins it.  Each function
cites the reference call site (file:line under /root/reference) it follows.
"""
