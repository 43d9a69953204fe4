"""Tensor-level wrappers over the C ABI building blocks (device pointers + current stream in, nothing else).

Used by the parity tests and by the host-side pipeline code.  Every function requires CUDA tensors and raises
`UtxError` on failure; there is no eager-PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_GATE_RES = 0, 1, 2


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype):
        raise _lib.UtxError(f"{name}: expected a CUDA {dtype} tensor, got {t.device} {t.dtype}")
    if t.stride(-1) != 1:
        raise _lib.UtxError(f"{name}: innermost dimension must be contiguous")


def gemm(A, W, bias=None, epi=EPI_BIAS, gate=None, res=None, out=None):
    """out[M,N] = epi(A[M,K] @ W[N,K]^T + bias)   (bf16 in/out, fp32 accumulate; gate fp32 [N])."""
    _req(A, torch.bfloat16, "A"); _req(W, torch.bfloat16, "W")
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=torch.bfloat16)
    L = _lib.load()
    _lib.check(L.utx_gemm_bf16(_p(A), A.stride(0), _p(W), W.stride(0), _p(bias), _p(out), out.stride(0), M, N, K, epi,
                               _p(gate), _p(res), 0 if res is None else res.stride(0), _stream()), "utx_gemm_bf16")
    return out


def gemm_qkv(A, W, bias, heads, wq, wk, cos, sin, row_offset=0, out=None):
    """QKV projection with per-head RMSNorm + RoPE fused into the epilogue for q and k."""
    _req(A, torch.bfloat16, "A"); _req(W, torch.bfloat16, "W")
    M, K = A.shape
    if out is None:
        out = torch.empty(M, 3 * heads * 128, device=A.device, dtype=torch.bfloat16)
    L = _lib.load()
    _lib.check(L.utx_gemm_bf16_qkv(_p(A), A.stride(0), _p(W), _p(bias), _p(out), out.stride(0), M, heads, K, _p(wq), _p(wk),
                                   _p(cos), _p(sin), row_offset, _stream()), "utx_gemm_bf16_qkv")
    return out


def gemm_grouped2(A0, W0, b0, C0, A1, W1, b1, C1, epi=EPI_BIAS, gate0=None, gate1=None):
    L = _lib.load()
    N, K = W0.shape
    _lib.check(L.utx_gemm_bf16_grouped2(_p(A0), A0.stride(0), _p(W0), _p(b0), _p(C0), C0.stride(0), A0.shape[0],
                                        _p(A1), A1.stride(0), _p(W1), _p(b1), _p(C1), C1.stride(0), A1.shape[0],
                                        N, K, epi, _p(gate0), _p(gate1), _stream()), "utx_gemm_bf16_grouped2")
    return C0, C1


def attention(qkv, H, out=None):
    """qkv [S, 3*H*128] bf16 (q|k|v) -> out [S, H*128]."""
    _req(qkv, torch.bfloat16, "qkv")
    S = qkv.shape[0]
    if out is None:
        out = torch.empty(S, H * 128, device=qkv.device, dtype=torch.bfloat16)
    L = _lib.load()
    _lib.check(L.utx_attention_bf16(_p(qkv), qkv.stride(0), _p(out), out.stride(0), S, H, _stream()), "utx_attention_bf16")
    return out


def ln_modulate(x, shift, scale, rows0=0, shift0=None, scale0=None, out=None):
    _req(x, torch.bfloat16, "x")
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    L = _lib.load()
    _lib.check(L.utx_ln_modulate(_p(x), x.stride(0), _p(out), out.stride(0), rows, D, rows0,
                                 _p(shift0 if shift0 is not None else shift), _p(scale0 if scale0 is not None else scale),
                                 _p(shift), _p(scale), _stream()), "utx_ln_modulate")
    return out


def rmsnorm_rope_(qkv, H, wq, wk, cos, sin, rows0=0, wq0=None, wk0=None):
    _req(qkv, torch.bfloat16, "qkv")
    L = _lib.load()
    _lib.check(L.utx_rmsnorm_rope(_p(qkv), qkv.stride(0), qkv.shape[0], H, rows0, _p(wq0 if wq0 is not None else wq),
                                  _p(wk0 if wk0 is not None else wk), _p(wq), _p(wk), _p(cos), _p(sin), _stream()),
               "utx_rmsnorm_rope")
    return qkv


def gemv(W, b, x, silu_in=False, out=None, accumulate=False):
    _req(W, torch.bfloat16, "W"); _req(x, torch.float32, "x")
    N, K = W.shape
    if out is None:
        out = torch.empty(N, device=W.device, dtype=torch.float32)
    L = _lib.load()
    _lib.check(L.utx_gemv_bf16(_p(W), _p(b), _p(x), _p(out), N, K, int(silu_in), int(accumulate), _stream()), "utx_gemv_bf16")
    return out


def rope_table(ids):
    _req(ids, torch.float32, "ids")
    S = ids.shape[0]
    cos = torch.empty(S, 128, device=ids.device, dtype=torch.float32)
    sin = torch.empty_like(cos)
    L = _lib.load()
    _lib.check(L.utx_rope_table(_p(ids.contiguous()), S, _p(cos), _p(sin), _stream()), "utx_rope_table")
    return cos, sin


def euler_update_(latents, v, rows, dsigma):
    L = _lib.load()
    _lib.check(L.utx_euler_update(_p(latents), _p(v), rows, latents.shape[-1], float(dsigma), _stream()), "utx_euler_update")
    return latents


def lora_merge_(W, A, B, scale):
    """W (bf16 [out,in], may be a row slice of a stacked weight) += scale * B @ A  (A, B fp32)."""
    _req(W, torch.bfloat16, "W")
    A = A.to(device=W.device, dtype=torch.float32).contiguous()
    B = B.to(device=W.device, dtype=torch.float32).contiguous()
    L = _lib.load()
    _lib.check(L.utx_lora_merge(_p(W), W.stride(0), _p(A), _p(B), W.shape[0], W.shape[1], A.shape[0], float(scale),
                                _stream()), "utx_lora_merge")
    return W
