#!/usr/bin/env python
"""bench.py -- multi-view denoise steps/s on the BASELINE.json config-2 workload.

One "step" = one FLUX.1-dev MM-DiT forward (19 double + 38 single blocks, 24 x 128 heads, LoRA merged) + the Euler
update over one texture_gen token grid: 1024x1024 canvas = 2x2 views of 512^2 -> 4096 noise + 4096 control + 1024
reference + 512 text tokens = S 9728, bf16, random-init weights, synthetic latents.

  python bench.py [--gpus N --steps K --warmup W]           our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference [...]                    the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

S_TXT, S_NOISE, S_CTRL, S_DUAL = 512, 4096, 4096, 1024
S_IMG = S_NOISE + S_CTRL + S_DUAL
S_TOT = S_TXT + S_IMG
WORKLOAD = "texture_gen 1024x1024 4-view (2x2 of 512^2): S=9728 = 512 txt + 4096 noise + 4096 control + 1024 reference"
METRIC, UNIT = "multi-view denoise steps/s", "steps/s"
LORA_RANK = 64


def bench_config(world: int) -> dict:
    """The workload description, IDENTICAL for the B200 arm and the reference arm (the driver compares the two dicts)."""
    return {"workload": WORKLOAD,
            "weights": f"random-init, FLUX.1-dev-shaped (19 double + 38 single blocks), rank-{LORA_RANK} texture_gen LoRA merged into "
                       "every target Linear (trainer.py:283-305: 12 per double block, q/k/v per single block) + x_embedder replacement",
            "tokens": S_TOT, "per_gpu_batch": 1, "parallelism": f"dp{world} (one independent grid per rank)",
            "schedule": "flow-match Euler, 28-step sigma grid, guidance 3.5",
            "l2": "inputs larger than L2: 23.8 GB of weights stream through every step"}


def _flops():
    """Algorithmic FLOPs of one step: total, linear (GEMM kernel) part, attention part."""
    D, M, L, Ls = 3072, 12288, 19, 38
    lin = (L * 2 * (4 * D * D + 2 * D * M) + Ls * 2 * (3 * D * D + D * M + (D + M) * D)) * S_TOT
    attn = (L + Ls) * 4 * D * S_TOT * S_TOT
    small = 2 * (S_IMG * 64 * D * 2) + 2 * D * D * (L * 12 + Ls * 3 + 2 + 6)
    return float(lin + attn + small), float(lin + 2 * S_IMG * 64 * D * 2), float(attn)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1386.9), d.get("bf16_tflops", 1649.4), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU reference arm
CPU_FRAC = 4   # the bounded CPU sample is exactly 1/CPU_FRAC of one block's work


def cpu_block_seconds(repeats: int = 1, device="cpu", frac: int = CPU_FRAC):
    """Times a 1/frac sample of ONE single-stream FLUX block at the full S=9728 with the oracle's eager op sequence
    (= diffusers' eager path restated, bf16 weights like the reference): LayerNorm-modulate, proj_mlp, GELU and proj_out
    on 1/frac of the token rows (rows are independent), q/k/v projection + RMSNorm + RoPE + SDPA for 1/frac of the heads
    over ALL 9728 keys (heads are independent).  All 57 blocks touch the same 113 246 208 params per token and do the same
    attention, so one step ~ 57 x frac x this.  device='cpu': the host-core baseline (frac 4);
    device=cuda, frac=1: the same eager sequence on the B200 through cuBLAS + torch SDPA ("the reference on B200" stand-in,
    SURVEY 8d) -- a reported baseline, never the product path."""
    import torch
    import torch.nn.functional as F
    from oracle import flux_dit as fd
    if str(device) == "cpu":
        # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would time one core)
        try:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except (AttributeError, RuntimeError):
            pass
    torch.manual_seed(0)
    cfg = fd.FluxConfig()
    D, H = cfg.inner_dim, cfg.num_attention_heads
    Hs, Sr = H // frac, S_TOT // frac
    bf = torch.bfloat16
    dev = torch.device(device)
    on_gpu = dev.type == "cuda"
    def rnd(*shape):
        return (torch.randn(*shape) * 0.02).to(bf).to(dev)
    w_mod, b_mod = rnd(3 * D, D), rnd(3 * D)
    w_qkv = [rnd(Hs * 128, D) for _ in range(3)]
    b_qkv = [rnd(Hs * 128) for _ in range(3)]
    w_mlp, b_mlp = rnd(4 * D, D), rnd(4 * D)
    w_out, b_out = rnd(D, 5 * D), rnd(D)
    rq, rk = torch.ones(128, dtype=bf, device=dev), torch.ones(128, dtype=bf, device=dev)
    x = torch.randn(1, S_TOT, D).to(bf).to(dev)
    temb = torch.randn(1, D).to(bf).to(dev)
    ids = torch.zeros(S_TOT, 3)
    ids[:, 1] = torch.arange(S_TOT) % 96
    ids[:, 2] = torch.arange(S_TOT) // 96
    cos, sin = (t.to(dev) for t in fd.rope_table(ids, cfg))
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            if on_gpu:
                torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            sh, sc, gate = F.linear(F.silu(temb), w_mod, b_mod).chunk(3, dim=1)
            nx = fd.layer_norm(x) * (1 + sc[:, None]) + sh[:, None]          # full rows: attention needs every key
            q, k, v = (F.linear(nx, w, b).view(1, S_TOT, Hs, 128).transpose(1, 2) for w, b in zip(w_qkv, b_qkv))
            q, k = fd.apply_rope(fd.rms_norm(q, rq), cos, sin), fd.apply_rope(fd.rms_norm(k, rk), cos, sin)
            ao = F.scaled_dot_product_attention(q.to(bf), k.to(bf), v).transpose(1, 2).reshape(1, S_TOT, Hs * 128)
            mlp = F.gelu(F.linear(nx[:, :Sr], w_mlp, b_mlp), approximate="tanh")
            cat = torch.cat([ao[:, :Sr].repeat(1, 1, frac), mlp], dim=2)
            y = x[:, :Sr] + gate.unsqueeze(1) * F.linear(cat, w_out, b_out)
            if on_gpu:
                torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            assert torch.isfinite(y.float()).all()
            best = dt if best is None else min(best, dt)
    return best * frac, torch.get_num_threads()


REF_SAMPLE = ("per step: ONE whole single-stream block (of 57 equal-cost blocks: every block touches the same 113 246 208 "
              "parameters per token and runs the same 24-head attention over all 9728 keys, SURVEY 8d) at the full S=9728 on the "
              "host cores -- all rows through LN/QKV/MLP/proj_out, all heads through RMSNorm+RoPE+SDPA -- eager oracle port of the "
              "diffusers CPU path, bf16 weights like the reference")


def run_reference(args, rank, world):
    """The reference's own implementation of the path (diffusers eager, CPU) cannot be installed here, so this arm
    times the oracle port of it.  One whole step is ~10 minutes of host time, so each timed step is a BOUNDED SAMPLE: one
    whole block, measured; the step time reported is that x 57 and the line says so in `extrapolated`."""
    if rank != 0:
        return
    times = []
    cores = 1
    for i in range(args.warmup + args.steps):
        dt, cores = cpu_block_seconds(1, frac=1)
        if i >= args.warmup:
            times.append(dt)
    block_s = sum(times) / len(times)
    step_s = 57.0 * block_s
    val = 1.0 / step_s
    out = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
           "data": "synthetic", "impl": "reference", "config": bench_config(world), "timing": "host wall clock",
           "extrapolated": {"factor": 57, "measured_unit": "one whole single-stream block at S=9728", "measured_ms": block_s * 1e3,
                            "note": "value = 1 / (57 x measured block time); the timed region of this run covers "
                                    f"{args.steps} blocks, not {args.steps} steps"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": REF_SAMPLE},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------- B200 arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from unitex_b200.flux import FluxConfig, FluxTransformer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    side = {}
    if world == 1 and not args.no_bake:      # before the engine: metric 2 is measured on an idle, un-throttled GPU
        side["uv_bake"] = bench_uv_bake(dev, return_tensors=True)
        bake_nn = side["uv_bake"].pop("tensors")[3].cpu().long()       # the GPU run's 1-NN table, handed to the CPU baseline of the same workload
        side["uv_bake_synthetic"] = bench_uv_bake(dev, mesh_name="two_spheres", reps=10)
        side["vae_decode"] = bench_vae_decode(dev)
    cfg = FluxConfig()
    eng = FluxTransformer(cfg, dev).random_init_(seed=0)
    # "LoRA merged" (BASELINE config 2): a random rank-64 texture_gen adapter over the reference's whole target list
    # (flux_piplines/texturing/trainer.py:283-305: the 12 Linears of every double block, to_q/k/v of every single block) plus
    # the x_embedder replacement, folded into the resident weights once before the loop: W' = W + s B A.  Merging adds no
    # work to the step, which is the point of merging.
    g = torch.Generator(device=dev).manual_seed(1)
    dbl_t = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj",
             "attn.to_add_out", "ff.net.0.proj", "ff.net.2", "ff_context.net.0.proj", "ff_context.net.2")
    names = [f"transformer_blocks.{i}.{t}" for i in range(cfg.num_layers) for t in dbl_t]
    names += [f"single_transformer_blocks.{i}.attn.{t}" for i in range(cfg.num_single_layers) for t in ("to_q", "to_k", "to_v")]
    adapter = {}
    for n in names:
        k, r0, r1 = eng.where[n]
        o, i_f = r1 - r0, eng.T["w_" + k].shape[1]
        adapter[n + ".lora_A.weight"] = torch.randn(LORA_RANK, i_f, device=dev, generator=g) * 0.02
        adapter[n + ".lora_B.weight"] = torch.randn(o, LORA_RANK, device=dev, generator=g) * 0.02
    adapter["x_embedder.weight"] = torch.randn(3072, 64, device=dev, generator=g) * 0.02
    adapter["x_embedder.bias"] = torch.randn(3072, device=dev, generator=g) * 0.02
    eng.merge_lora_(adapter, 1.0)
    n_lora_targets = len(names)
    del adapter
    torch.cuda.empty_cache()
    from unitex_b200.vae import AutoencoderKLB200
    from unitex_b200 import parallel as par
    vae = AutoencoderKLB200.from_random(seed=1, device=dev)

    ids = torch.zeros(S_TOT, 3)
    def grid(h, w, oy, ox):
        t = torch.zeros(h, w, 3)
        t[..., 1] = torch.arange(oy, oy + h)[:, None]
        t[..., 2] = torch.arange(ox, ox + w)[None, :]
        return t.reshape(-1, 3)
    ids[S_TXT:] = torch.cat([grid(64, 64, 0, 0), grid(64, 64, 64, 0), grid(32, 32, 64, 64)])   # pipeline.py:303-393
    eng.prepare(ids, None, None, s_txt=S_TXT)

    lat_host = torch.randn(S_IMG, 64, generator=torch.Generator().manual_seed(63 + rank)).to(torch.bfloat16).pin_memory()
    out_host = torch.empty(S_NOISE, 64, dtype=torch.bfloat16).pin_memory()
    lat = lat_host.to(dev, non_blocking=True)
    n_sched = 28
    import numpy as np
    s = np.linspace(1.0, 1.0 / n_sched, n_sched)
    mu = 1.15                                                          # calculate_shift(4096)
    sig = np.concatenate([np.exp(mu) / (np.exp(mu) + (1 / s - 1)), [0.0]]).astype(np.float32)

    def step(i):
        j = i % n_sched
        eng.denoise_(lat, S_NOISE, sig[j:j + 2], 3.5)

    def finish_asset():
        """What follows the loop for every asset (pipeline.py:688-692 + north_star): unpack the noise rows, VAE-decode them to
        the uint8 view tile [3,1024,1024] and -- the path's ONE collective -- all-gather the ranks' tiles so every rank holds
        every asset's tile before UV projection.  Real data: this rank's own latents through the real decode."""
        z = lat[:S_NOISE].view(1, 64, 64, 16, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(1, 16, 128, 128)
        img = vae.decode(z / vae.scaling_factor + vae.shift_factor)
        tile = ((img[0].float() / 2 + 0.5).clamp(0, 1) * 255.0 + 0.5).to(torch.uint8)
        return par.all_gather_grid_tiles([tile], world, (3, 1024, 1024), torch.uint8, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    tiles = finish_asset()                                             # untimed first pass (NCCL communicator set-up, VAE first launches)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    barrier()
    replays0 = eng.graph_replays()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    tiles = finish_asset()                                             # inside the timed region at every N (at N = 1: decode only)
    e2.record()
    barrier()
    ms = e0.elapsed_time(e2)
    ms_steps = e0.elapsed_time(e1)
    ms_finish = e1.elapsed_time(e2)
    graph_steps = eng.graph_replays() - replays0
    assert len(tiles) == world and all(t.shape == (3, 1024, 1024) for t in tiles)
    tile_checks = [int(t.sum().item()) for t in tiles]                 # every rank's tile arrived (non-trivial data)
    clocks = sampler.stop() if sampler else None

    # per-kernel breakdown for the roofline: the SAME K steps once more with the engine's per-launch CUDA events switched on
    # (eager launches: a step that runs as one graph launch cannot be bracketed kernel by kernel; the events themselves
    # cost a few microseconds per launch, which is why they are kept out of the region that produces `value`)
    eng.profile(True)
    eng.profile_read(reset=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.steps):
        step(i)
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    launches, cat_ms = eng.profile_read(reset=True)
    eng.profile(False)

    # end to end through the public call with HOST buffers: pinned H2D of the latents, one step, D2H of the result
    tile_host = torch.empty(world, 3, 1024, 1024, dtype=torch.uint8).pin_memory()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        lat.copy_(lat_host, non_blocking=True)
        step(i)
        out_host.copy_(lat[:S_NOISE], non_blocking=True)
    tiles = finish_asset()
    for r_, t_ in enumerate(tiles):
        tile_host[r_].copy_(t_, non_blocking=True)                     # the gathered tiles reach the host
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e, ms_steps, ms_finish], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_steps, ms_finish = t.tolist()
    sp = None
    if world > 1:          # side measurement, after every headline number is in hand; it must not be able to take the line down
        try:
            sp = bench_single_grid_sp(eng, dev, ids, sig, world, rank, args.steps)
        except Exception as e:
            sp = {"unavailable": repr(e)[:300]}
    if rank != 0:
        return
    total, f_gemm, f_attn = _flops()
    peak_sus, peak_burst, peak_src = _peaks()
    n_gemm = launches["gemm"] / args.steps
    gemm_tf = f_gemm * args.steps / (cat_ms["gemm"] * 1e-3) / 1e12 if cat_ms["gemm"] > 0 else 0.0
    attn_tf = f_attn * args.steps / (cat_ms["attn"] * 1e-3) / 1e12 if cat_ms["attn"] > 0 else 0.0
    step_ms = ms / args.steps
    out = {
        "metric": METRIC, "value": world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": bench_config(world), "timing": "CUDA events on the launching stream, max over ranks",
        "timed_region": {"what": f"{args.steps} denoise steps, then this rank's VAE decode to the uint8 view tile and the all-gather of "
                                 f"the {world} ranks' tiles (once per asset; the reference does it once per 28 steps, pipeline.py:688-692)",
                         "denoise_only_ms_per_step": ms_steps / args.steps, "decode_and_gather_ms": ms_finish,
                         "gathered_tile_checksums": tile_checks, "lora_targets_merged": n_lora_targets,
                         "steps_as_cuda_graph_launches": graph_steps,
                         "profiled_pass_ms_per_step": ms_profiled / args.steps,
                         "profiled_pass_note": "the roofline's per-kernel times come from a second pass over the same K steps with per-launch "
                                               "CUDA events on the launching stream (eager launches); `value` comes from the un-instrumented pass"},
        "roofline": {"bound": "tensor", "kernel": "gemm2_bf16_tn_kernel (tcgen05 cta_group::2; 1-CTA gemm_bf16_tn_kernel for narrow N)",
                     "achieved": gemm_tf, "peak": peak_sus, "unit": "TFLOP/s", "frac": gemm_tf / peak_sus,
                     "traffic": 5.20e8, "traffic_note": "dram read+write of ONE qkv-shaped launch (M 9728, N 9216, K 3072), warm L2 as inside the step (ncu --cache-control none, profiles/r01_gemm_groupm_sweep.txt: 335 MB read + 185 MB written); 1046 MB with ncu's cache flush (profiles/r01_gemm2_bf16_tn_final.ncu-rep); algorithmic 296 MB",
                     "peak_source": peak_src,
                     "launches_per_step": n_gemm, "flops_per_step": f_gemm, "ms_per_step": cat_ms["gemm"] / args.steps,
                     "attention": {"kernel": "attention2_kernel (tcgen05, P in TMEM)", "achieved": attn_tf, "frac": attn_tf / peak_sus,
                                   "flops_per_step": f_attn, "ms_per_step": cat_ms["attn"] / args.steps},
                     "elementwise_ms_per_step": cat_ms["elem"] / args.steps,
                     "whole_step": {"achieved": total / (ms_steps / args.steps * 1e-3) / 1e12,
                                    "frac": total / (ms_steps / args.steps * 1e-3) / 1e12 / peak_sus,
                                    "note": "DiT FLOPs of one step / denoise-only time per step (the decode + gather tail excluded)"}},
        "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": S_IMG * 64 * 2,
                "d2h_bytes_per_step": S_NOISE * 64 * 2, "d2h_bytes_once": world * 3 * 1024 * 1024,
                "note": "per step: latents from pinned host memory in, noise rows back out; after the loop: decode, gather, tiles to the host"},
        "gpu_launches": int(sum(launches.values())),
        "clocks": clocks,
    }
    out.update(side)
    if sp is not None:
        out["single_grid_sequence_parallel"] = sp
    if world == 1 and not args.no_bake:
        out["delight"] = bench_delight(eng, dev, sig)
        out["six_view"] = bench_six_view(eng, dev, sig)
        out["pipeline_call"] = bench_pipeline_call(eng, dev)
    if world == 1 and not args.no_cpu_baseline:
        del eng
        torch.cuda.empty_cache()
        cpu_block_seconds(2, device=dev, frac=1)                       # warm-up (cuBLAS heuristics, SDPA backend choice)
        dt_gpu, _ = cpu_block_seconds(5, device=dev, frac=1)
        out["gpu_eager_baseline"] = {"value": 1.0 / (57.0 * dt_gpu), "unit": UNIT, "kind": "port",
                                     "sample": "ONE full single-stream block at S=9728 through torch eager on this B200 (cuBLAS bf16 "
                                               "linears + torch SDPA, the oracle's op sequence); step = 57 x sample. The closest "
                                               "stand-in for 'the reference on B200' (SURVEY 8d); reported, not a target"}
        if "uv_bake" in out:
            try:
                out["uv_bake"]["cpu_baseline"] = bench_uv_bake_cpu(bake_nn)
            except Exception as e:      # the C oracle needs its prebuilt .so (or gcc) on the box; the main line must not depend on it
                out["uv_bake"]["cpu_baseline"] = {"unavailable": repr(e)[:200]}
        dt, cores = cpu_block_seconds(1, frac=1)
        out["cpu_baseline"] = {"value": 1.0 / (57.0 * dt), "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": REF_SAMPLE + "; step = 57 x the measured block", "measured_block_ms": dt * 1e3}
    print(json.dumps(out), flush=True)


def bench_single_grid_sp(eng, dev, ids, sig, world, rank, steps):
    """N > 1 side measurement: ONE texture_gen grid (the same S = 9728 workload) whose token sequence is split over all N ranks
    (sequence-parallel / Ulysses mode of the engine: include/unitex_b200.h utx_flux_set_sequence_parallel) -- the latency of a
    single asset's denoise step when the other GPUs would otherwise idle.  Same weights on every rank, rank 0's latents
    broadcast; CUDA events, max over ranks.  The headline `value` stays the batch-sharded throughput."""
    import torch
    import torch.distributed as dist
    from unitex_b200 import parallel as par
    if 24 % world or S_TOT % world:
        return {"unavailable": f"world size {world} does not divide 24 heads / {S_TOT} tokens"}
    res = {}
    for mode, direct in (("nccl_all_to_all", False), ("fused_peer_memory", True)):
        eng.set_sequence_parallel(par.tile_comm(dev), direct=direct)
        eng.prepare(ids, None, None, s_txt=S_TXT)
        lat = torch.randn(S_IMG, 64, generator=torch.Generator().manual_seed(63)).to(torch.bfloat16).to(dev)
        dist.broadcast(lat, 0)
        for i in range(2):
            eng.denoise_(lat, S_NOISE, sig[i:i + 2], 3.5)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            eng.denoise_(lat, S_NOISE, sig[i:i + 2], 3.5)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        chk = lat[:S_NOISE].float().sum().reshape(1)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        ms = t.item() / steps
        # where the step goes (rank 0's view): the engine's per-launch CUDA events by category; "other" = the exchanges
        # (all-to-alls + all-gather, or flag barriers + all-gather)
        eng.profile(True)
        eng.profile_read(reset=True)
        for i in range(2):
            eng.denoise_(lat, S_NOISE, sig[i:i + 2], 3.5)
        torch.cuda.synchronize()
        _, cat = eng.profile_read(reset=True)
        eng.profile(False)
        dist.barrier()
        res[mode] = {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "checksum": float(chk.item()),
                     "ranks_agree": bool(all(torch.equal(c, allc[0]) for c in allc)), "finite": bool(torch.isfinite(chk).all()),
                     "rank0_ms_per_step": {k: v / 2 for k, v in cat.items()}}
    eng.set_sequence_parallel(None)
    best = max(res.values(), key=lambda r: r["value"])
    return {"metric": "single-grid denoise steps/s (one asset over all GPUs)", "value": best["value"], "unit": UNIT,
            "ms_per_step": best["ms_per_step"], "n_gpus": world, "workload": WORKLOAD, "rows_per_rank": S_TOT // world,
            "heads_per_rank": 24 // world, "modes": res, "modes_agree": res["nccl_all_to_all"]["checksum"] == res["fused_peer_memory"]["checksum"],
            "modes_note": "nccl_all_to_all: 2 NCCL all-to-alls per block around the attention; fused_peer_memory: the QKV GEMM's and the "
                          "attention kernel's epilogues store into the owning rank's buffers over NVLink, 2 flag barriers per block; "
                          "both: 1 all-gather of v per step"}


def bench_delight(eng, dev, sig):
    """BASELINE config 3 side measurement: the delight call's shape, S = 512 txt + 4096 noise + 4096 control = 8704
    (no reference-image tokens), same engine and weights, CUDA events over 3 steps after 2 warm-up steps."""
    import torch
    s_img = S_NOISE + 4096
    ids = torch.zeros(S_TXT + s_img, 3)
    yy, xx = torch.meshgrid(torch.arange(64), torch.arange(64), indexing="ij")
    g = torch.stack([torch.zeros(64, 64), yy.float(), xx.float()], -1).reshape(-1, 3)
    g2 = g.clone()
    g2[:, 1] += 64
    ids[S_TXT:] = torch.cat([g, g2])
    eng.prepare(ids, None, None, s_txt=S_TXT)
    lat = torch.randn(s_img, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(5)).to(torch.bfloat16)
    for i in range(2):
        eng.denoise_(lat, S_NOISE, sig[i:i + 2], 3.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        eng.denoise_(lat, S_NOISE, sig[i:i + 2], 3.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    S = S_TXT + s_img
    flops = 57 * 2 * 113246208 * S + 57 * 4 * 3072 * S * S
    return {"metric": METRIC, "workload": "delight 1024x1024 4-view: S=8704 = 512 txt + 4096 noise + 4096 control", "value": 1e3 / ms,
            "unit": UNIT, "ms_per_step": ms, "tflops": flops / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(lat.float()).all())}


def bench_six_view(eng, dev, sig):
    """Side measurement at the shape UniTEX's own top-level pipeline runs (pipeline.py:246-259: 512 x 3072 strip of 6 views,
    control strip + 512^2 reference image): S = 512 txt + 6144 noise + 6144 control + 1024 reference = 13824 tokens
    (312.35 TFLOP per step).  Same engine and weights, CUDA events over 3 steps after 2 warm-up steps."""
    import torch
    s_noise, s_img = 6144, 6144 + 6144 + 1024

    def grid(h, w, oy, ox):
        t = torch.zeros(h, w, 3)
        t[..., 1] = torch.arange(oy, oy + h)[:, None]
        t[..., 2] = torch.arange(ox, ox + w)[None, :]
        return t.reshape(-1, 3)
    ids = torch.zeros(S_TXT + s_img, 3)
    ids[S_TXT:] = torch.cat([grid(32, 192, 0, 0), grid(32, 192, 32, 0), grid(32, 32, 32, 192)])   # flux pipeline :303-393, HL=64 WL=384
    eng.prepare(ids, None, None, s_txt=S_TXT)
    lat = torch.randn(s_img, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(6)).to(torch.bfloat16)
    for i in range(2):
        eng.denoise_(lat, s_noise, sig[i:i + 2], 3.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        eng.denoise_(lat, s_noise, sig[i:i + 2], 3.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    S = S_TXT + s_img
    flops = 57 * 2 * 113246208 * S + 57 * 4 * 3072 * S * S
    return {"metric": METRIC, "workload": "texture_gen 512x3072 6-view strip (the reference's own call): S=13824 = 512 txt + 6144 noise + 6144 control + 1024 reference",
            "value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "tflops": flops / (ms * 1e-3) / 1e12,
            "finite": bool(torch.isfinite(lat.float()).all())}


def bench_pipeline_call(eng, dev):
    """The whole public call of the sampler, as the reference makes it (pipeline.py:245-262): PIL control image 1024^2 + PIL
    reference image 512^2 in, VAE encode, 28 denoise steps at S = 9728, VAE decode, PIL image out.  Wall clock around the
    call with a device synchronize on both sides; random-init VAE weights."""
    import numpy as np
    import torch
    from PIL import Image
    from unitex_b200.flux_pipeline import PBRFluxPipeline
    from unitex_b200.vae import AutoencoderKLB200
    vae = AutoencoderKLB200.from_random(seed=1, device=dev)
    pipe = PBRFluxPipeline(eng, vae=vae)
    rng = np.random.default_rng(0)
    control = Image.fromarray(rng.integers(0, 255, (1024, 1024, 3), dtype=np.uint8))
    dual = Image.fromarray(rng.integers(0, 255, (512, 512, 3), dtype=np.uint8))
    kw = dict(prompt="", prompt_2="", dual_image=dual, control_image=control, height=1024, width=1024, n_rows=2, n_cols=2,
              num_inference_steps=28, guidance_scale=3.5, generator=torch.Generator().manual_seed(63))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    img = pipe(**kw).images[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = bool(np.isfinite(np.asarray(img, dtype=np.float32)).all()) and img.size == (1024, 1024)
    del pipe, vae
    torch.cuda.empty_cache()
    return {"what": "PBRFluxPipeline.__call__: PIL in -> VAE encode -> 28 steps at S=9728 -> VAE decode -> PIL out", "seconds": dt,
            "steps": 28, "steps_per_s_whole_call": 28 / dt, "ok": ok}


def bench_vae_decode(dev):
    """BASELINE config 3 side measurement: FLUX VAE decode of the 1024^2 canvas ([1,16,128,128] -> [1,3,1024,1024]),
    random-init weights, CUDA events.  Algorithmic FLOPs 10.47 T."""
    import torch
    from unitex_b200.vae import AutoencoderKLB200
    vae = AutoencoderKLB200.from_random(seed=1, device=dev)
    z = torch.randn(1, 16, 128, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(3)).to(torch.bfloat16)
    for _ in range(2):
        img = vae.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        img = vae.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    del vae
    torch.cuda.empty_cache()
    return {"metric": "VAE decode ms @1024^2", "ms": ms, "tflops": 10.47e12 / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(img.float()).all()),
            "note": "implicit-GEMM 3x3 convolutions (TMA boxes shifted by the tap) on the tcgen05 GEMM, NHWC bf16, deterministic GroupNorm+SiLU"}


# dram__bytes_read.sum + dram__bytes_write.sum summed over every kernel of ONE bake of the teaser_robot workload, from the ncu pass
# committed as profiles/r02_bake_launches_final.csv (scripts/profile_bake_teaser.py under `ncu --metrics dram__bytes_read.sum,
# dram__bytes_write.sum,gpu__time_duration.sum --profile-from-start off`)
BAKE_TRAFFIC_BYTES = 1.902e9   # 1553.0 MB read + 349.2 MB written (cold caches between kernels: ncu flushes them); the leaf grids of the
                               # ray stage trade 0.6 GB of extra list traffic for 2 ms (the tree walk of round 1: 1.30 GB)


def _two_sphere_mesh():
    import numpy as np

    def sphere(rows, cols, radius, center, rect):
        th = np.linspace(0.02, np.pi - 0.02, rows + 1)
        ph = np.linspace(0, 2 * np.pi, cols + 1)
        T, P = np.meshgrid(th, ph, indexing="ij")
        v = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], -1).reshape(-1, 3) * radius + np.asarray(center)
        uv = np.stack([rect[0] + rect[2] * (P / (2 * np.pi)), rect[1] + rect[3] * (T / np.pi)], -1).reshape(-1, 2)
        r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        a = (r * (cols + 1) + c).reshape(-1)
        b, d, e = a + 1, a + cols + 1, a + cols + 2
        f = np.concatenate([np.stack([a, b, d], -1), np.stack([b, e, d], -1)])
        return v.astype(np.float32), f.astype(np.int32), uv.astype(np.float32)

    v1, f1, uv1 = sphere(316, 632, 0.55, (0, 0, 0), (0.01, 0.01, 0.48, 0.98))
    v2, f2, uv2 = sphere(158, 316, 0.25, (0.62, 0.1, 0.05), (0.51, 0.01, 0.48, 0.98))
    f = np.concatenate([f1, f2 + len(v1)])
    return np.concatenate([v1, v2]), f, np.concatenate([uv1, uv2]) * 2 - 1, f.copy()


def bench_uv_bake(dev, return_tensors=False, mesh_name="teaser_robot", reps=20):
    """Second metric of BASELINE.json: UV-bake Mpix/s = atlas texels / time of NVDiffRendererInverse.infer(method='reproject')
    with the mesh resident on the GPU (SURVEY 8d metric 2).  Workload = BASELINE config 4's geometry: the reference's own
    test mesh test_cases/teaser_robot (V 269 026, F 499 981; lossless fixture tests/golden/teaser_robot.npz.xz), the 6 box views
    the reference bake is hard-wired to (renderer_inverse.py:171,256) at 512^2 with an analytic colour field, 2048^2 atlas.
    mesh_name='two_spheres': the synthetic mesh of round 1 (side key `uv_bake_synthetic`)."""
    import numpy as np
    import torch
    from unitex_b200 import bake as ub
    if mesh_name == "teaser_robot":
        from tests.bake_meshes import teaser_robot
        v, f, uv, fuv = teaser_robot()
        label = "teaser_robot (reference test_cases/teaser_robot/inputmesh.obj)"
    else:
        v, f, uv, fuv = _two_sphere_mesh()
        label = "synthetic 2-sphere mesh"
    V, F = len(v), len(f)
    H = W = 512
    H2 = W2 = 2048
    c2ws = ub.generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]]
    intr = ub.generate_intrinsics(1.0, 1.0, fov=False)
    mesh = ub.BakeMesh(v, f, uv, fuv, device=dev)
    r = ub.NVDiffRendererInverse(device=dev, pbr_mesh=mesh)
    ub.RayTracing(mesh.vertices, mesh.faces, device=dev)   # untimed: first launch of the build kernels (lazy module load, cub temp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mesh.optix                                     # LBVH build (once per mesh)
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    mats = torch.matmul(ub.intr_to_proj(intr, perspective=False), ub.c2w_to_w2c(c2ws)).to(dev)
    rast = ub.rasterize(ub.transform_points(mesh.vertices, mats), mesh.faces, (H, W))
    pos = ub.interpolate(mesh.vertices, rast, mesh.faces)
    img = (0.5 + 0.4 * torch.sin(3.0 * pos + 0.3)) * (rast[..., 3:4] > 0)
    kw = dict(H=H, W=W, H2D=H2, W2D=W2, perspective=False, ray_normal_angle_threhold=100.0, method="reproject",
              filt_gradient_points=False)
    for _ in range(3):
        r.infer(mesh, c2ws, intr, img, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        _, vis, m2, col = r.infer(mesh, c2ws, intr, img, **kw)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3 / reps
    gpu_ms = e0.elapsed_time(e1) / reps
    T = H2 * W2
    algo = V * 24 + len(uv) * 8 + F * 24 + (2 * F - 1) * 36 + 6 * H * W * 16 + T * 12 + T + 6 * T
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6589.6) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    covered = int(m2.sum())
    nn = r.last_nn_index
    extra = {"tensors": (vis, m2, col, nn)} if return_tensors else {}
    traffic = BAKE_TRAFFIC_BYTES if mesh_name == "teaser_robot" else None
    return {**extra, "metric": "UV-bake Mpix/s", "value": T / 1e6 / (wall_ms * 1e-3), "unit": "Mpix/s", "ms_per_bake": wall_ms,
            "gpu_ms_per_bake": gpu_ms, "bvh_build_ms": build_ms,
            "config": {"workload": f"{label}: V={V} F={F}, 6 box views 512^2, atlas 2048^2, method=reproject",
                       "covered_texels": covered, "rays": 6 * covered, "visible_texels": int(vis.any(dim=0).sum()),
                       "nn_queries": int((nn >= 0).sum())},
            "roofline": {"bound": "hbm", "achieved": algo / (gpu_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": algo / (gpu_ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": algo, "traffic": traffic,
                         "traffic_source": "profiles/r02_bake_launches_final.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum over every kernel of one bake)" if traffic else None,
                         "note": "compulsory bytes of SURVEY 8d; the bake is BVH-traversal (latency) bound, not streaming"},
            "mrays_per_s": 6 * covered / 1e6 / (gpu_ms * 1e-3)}


def bench_uv_bake_cpu(nn_index=None):
    """CPU baseline of metric 2: the oracle's restatement of NVDiffRendererInverse.infer (C rasteriser / LBVH tracer of
    oracle/bake_ref.c + the reference's torch tail) on the host cores.
    With `nn_index` (the 1-NN table of the GPU run, int tensor [2048^2]): the SAME workload as the GPU figure -- teaser_robot,
    6 views 512^2, atlas 2048^2 -- with the nearest-neighbour table supplied instead of searched (the oracle's exact search is
    brute force, 0.7 M x 1.4 M pairs; the reference uses a kd-tree there), so the CPU figure is optimistic by that stage.
    Without: a bounded small case (two-sphere mesh with 8 k faces, 6 views 128^2, atlas 256^2), everything computed."""
    import numpy as np
    import torch
    from oracle import bake as ob
    from tests.bake_meshes import analytic_color, teaser_robot, two_spheres
    from unitex_b200.bake import generate_box_views_c2ws, generate_intrinsics
    c2ws, intr = generate_box_views_c2ws(2.8)[[0, 1, 4, 2, 3, 5]], generate_intrinsics(1.0, 1.0, fov=False)
    mats = torch.matmul(ob.intr_to_proj_ortho(intr), ob.c2w_to_w2c(c2ws))
    if nn_index is not None:
        v, f, uv, fuv = teaser_robot()
        H, H2 = 512, 2048
        sample = ("oracle infer(method='reproject') on the same workload: teaser_robot, 6 views 512^2, atlas 2048^2; the 1-NN table of "
                  "the fill is supplied (not searched), everything else computed; measured, not extrapolated")
    else:
        v, f, uv, fuv = two_spheres(40, 80)
        H, H2 = 128, 256
        sample = "oracle infer(method='reproject') on a two-sphere mesh with 8000 faces, 6 views 128^2, atlas 256^2; measured, not extrapolated"
    rast = ob.rasterize(torch.matmul(torch.cat([torch.from_numpy(v), torch.ones(len(v), 1)], -1), mats.permute(0, 2, 1)).numpy(), f, H, H)
    img = torch.from_numpy(analytic_color(ob.interpolate(v, rast, f)) * (rast[..., 3:4] > 0)).float()
    t0 = time.perf_counter()
    ob.infer(v, f, uv, fuv, c2ws, intr, img, H, H, H2, H2, nn_index_given=nn_index)
    dt = time.perf_counter() - t0
    return {"value": H2 * H2 / 1e6 / dt, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port", "seconds": dt, "sample": sample}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bake", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
