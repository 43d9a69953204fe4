"""GPU parity of every sm_100a building block against fp32 references (called through the C ABI via ctypes)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _rel_err(out, ref):
    return ((out.float() - ref).norm() / ref.norm().clamp_min(1e-20)).item()


def _check_bf16(out, ref, tol_ulp=1.0, frac=0.999, name=""):
    """out (bf16) must equal the fp32 reference to within `tol_ulp` bf16 ulps for >= frac of the elements
    and 4 ulps everywhere (fp32 accumulation order differs from the reference's)."""
    out, ref = out.float(), ref.float()
    ulp = torch.maximum(ref.abs(), torch.full_like(ref, 1e-3)) * 2.0 ** -7
    err = (out - ref).abs() / ulp
    assert torch.isfinite(out).all(), name
    assert (err <= tol_ulp).float().mean().item() >= frac, f"{name}: {(err <= tol_ulp).float().mean().item()}"
    assert err.max().item() <= 4.0 * tol_ulp + 4.0, f"{name}: max err {err.max().item()} ulp"


@pytest.fixture(params=["1", "2"])
def gemm_impl(request, monkeypatch):
    """1 = one CTA per 128x256 tile; 2 = cta_group::2 pairs on 256x256 tiles (used when N % 256 == 0)."""
    monkeypatch.setenv("UTX_GEMM_IMPL", request.param)
    return request.param


GEMM_SHAPES = [(128, 256, 64), (256, 512, 256), (1000, 768, 1280), (512, 64, 256), (130, 200, 128),
               (384, 128, 3072), (2304, 3072, 3072), (640, 1792, 256)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm(lib, gemm_impl, M, N, K, epi):
    from unitex_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    A = _bf(torch.randn(M, K, device="cuda", generator=g))
    W = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = _bf(torch.randn(N, device="cuda", generator=g))
    ref = A.float() @ W.float().T + bias.float()
    if epi == 0:
        out = ops.gemm(A, W, bias)
    elif epi == 1:
        out = ops.gemm(A, W, bias, epi=ops.EPI_BIAS_GELU)
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    else:
        gate = torch.randn(N, device="cuda", generator=g)
        res = _bf(torch.randn(M, N, device="cuda", generator=g))
        out = res.clone()
        ops.gemm(A, W, bias, epi=ops.EPI_GATE_RES, gate=gate, res=out, out=out)    # in place like the engine
        ref = res.float() + gate * ref
    torch.cuda.synchronize()
    # tanh.approx in the GELU epilogue: a few 1e-4 absolute
    _check_bf16(out, ref, tol_ulp=1.0 if epi != 1 else 1.5, name=f"gemm {M}x{N}x{K} epi{epi}")


def test_gemm_strided_views_and_no_bias(lib, gemm_impl):
    from unitex_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    big = _bf(torch.randn(300, 1280, device="cuda", generator=g))
    A = big[:, 256:1280]                       # lda = 1280, K = 1024
    W = _bf(torch.randn(256, 1024, device="cuda", generator=g) / 32)
    outbig = torch.zeros(300, 768, device="cuda", dtype=torch.bfloat16)
    out = outbig[:, 256:512]
    ops.gemm(A, W, None, out=out)
    torch.cuda.synchronize()
    _check_bf16(out, A.float() @ W.float().T, name="strided")
    assert outbig[:, :256].abs().max() == 0 and outbig[:, 512:].abs().max() == 0


def test_gemm_grouped_two_streams(lib, gemm_impl):
    from unitex_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    K, N, M0, M1 = 256, 768, 128, 1152
    X = _bf(torch.randn(M0 + M1, K, device="cuda", generator=g))
    W0, W1 = (_bf(torch.randn(N, K, device="cuda", generator=g) / 16) for _ in range(2))
    b0, b1 = (_bf(torch.randn(N, device="cuda", generator=g)) for _ in range(2))
    C = torch.empty(M0 + M1, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_grouped2(X[:M0], W0, b0, C[:M0], X[M0:], W1, b1, C[M0:])
    torch.cuda.synchronize()
    _check_bf16(C[:M0], X[:M0].float() @ W0.float().T + b0.float(), name="grouped txt")
    _check_bf16(C[M0:], X[M0:].float() @ W1.float().T + b1.float(), name="grouped img")


@pytest.mark.parametrize("S,H,qscale", [(128, 2, 1.0), (256, 2, 1.0), (1280, 2, 1.0), (1000, 3, 1.0), (640, 2, 6.0),
                                        (2432, 4, 1.0), (300, 1, 1.0)])
def test_attention(lib, S, H, qscale):
    from unitex_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(S + H)
    qkv = torch.randn(S, 3 * H * 128, device="cuda", generator=g)
    qkv[:, : H * 128] *= qscale          # large logits exercise the lazy-rescale path
    qkv = _bf(qkv)
    out = ops.attention(qkv, H)
    torch.cuda.synchronize()
    q, k, v = (t.float().view(S, H, 128).transpose(0, 1) for t in qkv.split(H * 128, dim=1))
    ref = torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1).reshape(S, -1)
    assert torch.isfinite(out.float()).all()
    rel = _rel_err(out, ref)
    assert rel < 1.5e-2, f"attention rel err {rel}"
    assert (out.float() - ref).abs().max().item() < 0.05 * ref.abs().max().item() + 2e-2


@pytest.mark.parametrize("S,slope", [(640, 0.02), (1000, 0.05), (384, -0.03)])
def test_attention_growing_logits(lib, S, slope):
    """Logits that climb (or fall) steadily along the key axis: every key tile outgrows the running reference, in both of
    its halves -- the lazy-reference fast path must fall back to the exact maximum and rescale O / l each time."""
    from unitex_b200 import ops
    H = 2
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.ones(S, H * 128, device="cuda") + 0.1 * torch.randn(S, H * 128, device="cuda", generator=g)
    ramp = (torch.arange(S, device="cuda", dtype=torch.float32) * slope)[:, None]
    k = ramp * torch.ones(S, H * 128, device="cuda") + 0.1 * torch.randn(S, H * 128, device="cuda", generator=g)
    v = torch.randn(S, H * 128, device="cuda", generator=g)
    qkv = _bf(torch.cat([q, k, v], dim=1))
    out = ops.attention(qkv, H)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().view(S, H, 128).transpose(0, 1) for t in qkv.split(H * 128, dim=1))
    ref = torch.nn.functional.scaled_dot_product_attention(qf[None], kf[None], vf[None])[0].transpose(0, 1).reshape(S, -1)
    assert torch.isfinite(out.float()).all()
    rel = _rel_err(out, ref)
    assert rel < 1.5e-2, f"attention rel err {rel}"


def test_attention_matches_explicit_bf16_p(lib):
    """Tighter: emulate the kernel's one deliberate rounding (P -> bf16 before PV) in fp32 torch."""
    from unitex_b200 import ops
    S, H = 384, 2
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = _bf(torch.randn(S, 3 * H * 128, device="cuda", generator=g))
    out = ops.attention(qkv, H)
    torch.cuda.synchronize()
    q, k, v = (t.float().view(S, H, 128).transpose(0, 1) for t in qkv.split(H * 128, dim=1))
    s = q @ k.transpose(-1, -2) / math.sqrt(128)
    p = torch.exp(s - s.amax(-1, keepdim=True))
    ref = (p.to(torch.bfloat16).float() @ v) / p.sum(-1, keepdim=True)
    ref = ref.transpose(0, 1).reshape(S, -1)
    assert _rel_err(out, ref) < 6e-3


@pytest.mark.parametrize("rows,D,rows0", [(64, 256, 0), (1000, 3072, 300), (7, 512, 7), (640, 1024, 128)])
def test_ln_modulate(lib, rows, D, rows0):
    from unitex_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = _bf(torch.randn(rows, D, device="cuda", generator=g) * 3 + 0.5)
    sh0, sc0, sh1, sc1 = (torch.randn(D, device="cuda", generator=g) * 0.3 for _ in range(4))
    out = ops.ln_modulate(x, sh1, sc1, rows0=rows0, shift0=sh0, scale0=sc0)
    torch.cuda.synchronize()
    ln = torch.nn.functional.layer_norm(x.float(), (D,), eps=1e-6)
    ref = torch.cat([ln[:rows0] * (1 + sc0) + sh0, ln[rows0:] * (1 + sc1) + sh1])
    _check_bf16(out, ref, name="ln_modulate")


def test_rope_table_and_rmsnorm_rope(lib):
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from unitex_b200 import ops
    cfg = fd.FluxConfig.tiny(heads=4)
    H, S_txt = 4, 64
    ids = torch.cat([torch.zeros(S_txt, 3), fs.build_ids(32, 48, (32, 48), (16, 16))]).cuda()
    S = ids.shape[0]
    cos, sin = ops.rope_table(ids)
    rc, rs = fd.rope_table(ids, cfg)
    torch.cuda.synchronize()
    assert (cos - rc).abs().max().item() < 2e-7 and (sin - rs).abs().max().item() < 2e-7
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = _bf(torch.randn(S, 3 * H * 128, device="cuda", generator=g) * 2)
    w = [_bf(1 + 0.2 * torch.randn(128, device="cuda", generator=g)) for _ in range(4)]   # wq_txt, wk_txt, wq_img, wk_img
    got = ops.rmsnorm_rope_(qkv.clone(), H, w[2], w[3], cos, sin, rows0=S_txt, wq0=w[0], wk0=w[1])
    torch.cuda.synchronize()
    x = qkv.float().view(S, 3, H, 128)

    def ref_one(t, w_txt, w_img):
        t = t.transpose(0, 1)[None]                       # [1,H,S,128]
        n = torch.cat([fd.rms_norm(t[:, :, :S_txt], w_txt.float()), fd.rms_norm(t[:, :, S_txt:], w_img.float())], 2)
        return fd.apply_rope(n, rc, rs)[0].transpose(0, 1)

    _check_bf16(got.view(S, 3, H, 128)[:, 0], ref_one(x[:, 0], w[0], w[2]), name="q")
    _check_bf16(got.view(S, 3, H, 128)[:, 1], ref_one(x[:, 1], w[1], w[3]), name="k")
    assert torch.equal(got.view(S, 3, H, 128)[:, 2], qkv.view(S, 3, H, 128)[:, 2])       # v untouched


def test_gemm_qkv_fused_norm_rope(lib, gemm_impl):
    """QKV GEMM with RMSNorm + RoPE in the epilogue == fp32 reference of Linear -> RMSNorm -> RoPE (v untouched)."""
    from oracle import flux_dit as fd
    from oracle import flux_sampler as fs
    from unitex_b200 import ops
    H, K, off = 4, 512, 64
    cfg = fd.FluxConfig.tiny(heads=H)
    ids = torch.cat([torch.zeros(off, 3), fs.build_ids(32, 48, (32, 48), None)]).cuda()
    M = ids.shape[0] - off
    cos, sin = ops.rope_table(ids)
    g = torch.Generator(device="cuda").manual_seed(4)
    A = _bf(torch.randn(M, K, device="cuda", generator=g))
    W = _bf(torch.randn(3 * H * 128, K, device="cuda", generator=g) / math.sqrt(K))
    b = _bf(torch.randn(3 * H * 128, device="cuda", generator=g) * 0.1)
    wq, wk = (_bf(1 + 0.2 * torch.randn(128, device="cuda", generator=g)) for _ in range(2))
    out = ops.gemm_qkv(A, W, b, H, wq, wk, cos, sin, row_offset=off)
    torch.cuda.synchronize()
    lin = (A.float() @ W.float().T + b.float()).view(M, 3, H, 128)
    rc, rs = cos[off:], sin[off:]
    ref_q = fd.apply_rope(fd.rms_norm(lin[:, 0].transpose(0, 1)[None], wq.float()), rc, rs)[0].transpose(0, 1)
    ref_k = fd.apply_rope(fd.rms_norm(lin[:, 1].transpose(0, 1)[None], wk.float()), rc, rs)[0].transpose(0, 1)
    o = out.view(M, 3, H, 128)
    _check_bf16(o[:, 0], ref_q, name="q")
    _check_bf16(o[:, 1], ref_k, name="k")
    _check_bf16(o[:, 2], lin[:, 2], name="v")


def test_gemv_euler_lora(lib):
    from unitex_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    W = _bf(torch.randn(1000, 3072, device="cuda", generator=g) / 55)
    b = _bf(torch.randn(1000, device="cuda", generator=g))
    x = torch.randn(3072, device="cuda", generator=g)
    y = ops.gemv(W, b, x, silu_in=True)
    ref = W.float() @ torch.nn.functional.silu(x) + b.float()
    torch.cuda.synchronize()
    assert torch.allclose(y, ref, rtol=1e-4, atol=1e-4)
    y2 = ops.gemv(W, None, x, silu_in=False, out=y.clone(), accumulate=True)
    torch.cuda.synchronize()
    assert torch.allclose(y2, ref + W.float() @ x, rtol=1e-4, atol=2e-4)

    lat = _bf(torch.randn(96, 64, device="cuda", generator=g))
    v = _bf(torch.randn(96, 64, device="cuda", generator=g))
    for ds in (-0.03125, -0.0371094):
        want = lat.clone()
        dt = torch.tensor(ds, dtype=torch.float32, device="cuda")              # 0-dim fp32 sigma difference, as in diffusers' step
        want[:64] = (lat[:64].float() + dt * v[:64]).to(torch.bfloat16)        # dt * v is formed in bf16 (torch promotion), then fp32 add
        got = ops.euler_update_(lat.clone(), v, 64, ds)
        torch.cuda.synchronize()
        assert torch.equal(got, want)                  # mul -> bf16 RN -> fp32 add -> bf16 RN: bit-exact

    Wl = _bf(torch.randn(3 * 256, 320, device="cuda", generator=g) * 0.05)
    A = torch.randn(16, 320, device="cuda", generator=g) * 0.1
    B = torch.randn(256, 16, device="cuda", generator=g) * 0.1
    Wm = Wl.clone()
    ops.lora_merge_(Wm[256:512], A, B, 0.7)            # merge into the k slice of a stacked qkv weight
    torch.cuda.synchronize()
    want = Wl.clone()
    want[256:512] = (Wl[256:512].float() + 0.7 * (B @ A)).to(torch.bfloat16)
    assert (Wm.float() - want.float()).abs().max().item() <= 2.0 ** -8 * want.abs().max().item()
    assert torch.equal(Wm[:256], Wl[:256]) and torch.equal(Wm[512:], Wl[512:])
