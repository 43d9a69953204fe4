"""ctypes binding of libunitex_b200.so (the C ABI declared in include/unitex_b200.h).

There is no fallback: if the library is missing or a call fails this raises.  PyTorch is only used by the callers for
device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libunitex_b200.so"
_lib = None

vp, i32, f32, lng, fp = C.c_void_p, C.c_int, C.c_float, C.c_long, C.POINTER(C.c_float)


class FluxConfigC(C.Structure):
    _fields_ = [(n, i32) for n in ("in_channels", "num_layers", "num_single_layers", "num_heads", "head_dim",
                                   "joint_attention_dim", "pooled_projection_dim", "guidance_embeds", "mlp_ratio")]


_DBL = ("w_qkv_img", "b_qkv_img", "w_qkv_txt", "b_qkv_txt", "rms_q_img", "rms_k_img", "rms_q_txt", "rms_k_txt",
        "w_out_img", "b_out_img", "w_out_txt", "b_out_txt", "w_ff1_img", "b_ff1_img", "w_ff2_img", "b_ff2_img",
        "w_ff1_txt", "b_ff1_txt", "w_ff2_txt", "b_ff2_txt")
_SGL = ("w_qkvmlp", "b_qkvmlp", "rms_q", "rms_k", "w_out", "b_out")


class DoubleBlockC(C.Structure):
    _fields_ = [(n, vp) for n in _DBL]


class SingleBlockC(C.Structure):
    _fields_ = [(n, vp) for n in _SGL]


class FluxWeightsC(C.Structure):
    _fields_ = ([(n, vp) for n in ("w_x_embed", "b_x_embed", "w_ctx_embed", "b_ctx_embed", "w_t1", "b_t1", "w_t2", "b_t2",
                                   "w_g1", "b_g1", "w_g2", "b_g2", "w_p1", "b_p1", "w_p2", "b_p2", "w_mod", "b_mod")]
                + [("double_blocks", C.POINTER(DoubleBlockC)), ("single_blocks", C.POINTER(SingleBlockC)),
                   ("w_proj_out", vp), ("b_proj_out", vp)])


class VaeConfigC(C.Structure):
    _fields_ = [("in_channels", i32), ("latent_channels", i32), ("num_blocks", i32), ("block_out_channels", i32 * 8),
                ("layers_per_block", i32), ("norm_num_groups", i32)]


_VRES = ("gn1_w", "gn1_b", "conv1_w", "conv1_b", "gn2_w", "gn2_b", "conv2_w", "conv2_b", "short_w", "short_b")
_VATT = ("gn_w", "gn_b", "wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo")


class VaeResnetC(C.Structure):
    _fields_ = [(n, vp) for n in _VRES] + [("cin", i32), ("cout", i32)]


class VaeAttnC(C.Structure):
    _fields_ = [(n, vp) for n in _VATT]


class VaeMidC(C.Structure):
    _fields_ = [("res0", VaeResnetC), ("attn", VaeAttnC), ("res1", VaeResnetC)]


class VaeWeightsC(C.Structure):
    _fields_ = [("enc_conv_in_w", vp), ("enc_conv_in_b", vp), ("enc_res", C.POINTER(VaeResnetC)), ("enc_down_w", C.POINTER(vp)),
                ("enc_down_b", C.POINTER(vp)), ("enc_mid", VaeMidC), ("enc_norm_w", vp), ("enc_norm_b", vp),
                ("enc_conv_out_w", vp), ("enc_conv_out_b", vp), ("dec_conv_in_w", vp), ("dec_conv_in_b", vp), ("dec_mid", VaeMidC),
                ("dec_res", C.POINTER(VaeResnetC)), ("dec_up_w", C.POINTER(vp)), ("dec_up_b", C.POINTER(vp)),
                ("dec_norm_w", vp), ("dec_norm_b", vp), ("dec_conv_out_w", vp), ("dec_conv_out_b", vp)]


_SIGS = {
    "utx_last_error": (C.c_char_p, []),
    "utx_version": (i32, []),
    "utx_flux_create": (i32, [C.POINTER(FluxConfigC), C.POINTER(vp)]),
    "utx_flux_destroy": (None, [vp]),
    "utx_flux_set_weights": (i32, [vp, C.POINTER(FluxWeightsC)]),
    "utx_flux_workspace_bytes": (C.c_size_t, [vp, i32, i32]),
    "utx_flux_prepare": (i32, [vp, vp, C.c_size_t, vp, vp, vp, i32, i32, vp]),
    "utx_flux_forward": (i32, [vp, vp, f32, f32, vp, vp]),
    "utx_flux_denoise": (i32, [vp, vp, i32, fp, i32, f32, vp]),
    "utx_flux_graph_replays": (C.c_long, [vp]),
    "utx_flux_set_sequence_parallel": (i32, [vp, vp]),
    "utx_flux_sp_region_bytes": (C.c_size_t, [vp, i32, i32]),
    "utx_flux_set_sp_peers": (i32, [vp, C.POINTER(vp), C.c_size_t]),
    "utx_peer_alloc": (i32, [C.POINTER(vp), C.c_size_t]),
    "utx_peer_free": (None, [vp]),
    "utx_peer_export": (i32, [vp, vp]),
    "utx_peer_import": (i32, [vp, C.POINTER(vp)]),
    "utx_peer_close": (None, [vp]),
    "utx_flux_profile": (i32, [vp, i32]),
    "utx_flux_profile_read": (i32, [vp, C.POINTER(C.c_long), fp, i32]),
    "utx_lora_merge": (i32, [vp, lng, vp, vp, i32, i32, i32, f32, vp]),
    "utx_gemm_bf16": (i32, [vp, lng, vp, lng, vp, vp, lng, i32, i32, i32, i32, vp, vp, lng, vp]),
    "utx_gemm_bf16_qkv": (i32, [vp, lng, vp, vp, vp, lng, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
    "utx_gemm_bf16_grouped2": (i32, [vp, lng, vp, vp, vp, lng, i32, vp, lng, vp, vp, vp, lng, i32, i32, i32, i32, vp, vp, vp]),
    "utx_attention_bf16": (i32, [vp, lng, vp, lng, i32, i32, vp]),
    "utx_ln_modulate": (i32, [vp, lng, vp, lng, i32, i32, i32, vp, vp, vp, vp, vp]),
    "utx_rmsnorm_rope": (i32, [vp, lng, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "utx_gemv_bf16": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "utx_rope_table": (i32, [vp, i32, vp, vp, vp]),
    "utx_euler_update": (i32, [vp, vp, i32, i32, f32, vp]),
    "utx_rasterize_workspace_bytes": (C.c_size_t, [i32, i32, i32, i32]),
    "utx_rasterize": (i32, [vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp]),
    "utx_interpolate": (i32, [vp, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp]),
    "utx_transform_points": (i32, [vp, i32, vp, i32, vp, vp]),
    "utx_bvh_nodes_bytes": (C.c_size_t, [i32]),
    "utx_bvh_workspace_bytes": (C.c_size_t, [i32]),
    "utx_bvh_build": (i32, [vp, i32, vp, i32, vp, vp, C.c_size_t, vp]),
    "utx_bvh_export": (i32, [vp, i32, vp, vp, vp]),
    "utx_bvh_intersect": (i32, [vp, vp, vp, i32, vp, vp, C.c_longlong, vp, vp, vp, vp, vp]),
    "utx_conv3x3_nhwc": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, C.c_long, vp, vp, C.c_long, vp]),
    "utx_im2col3x3": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "utx_upsample2x_nhwc": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "utx_groupnorm_workspace_bytes": (C.c_size_t, [i32, i32, i32, i32]),
    "utx_groupnorm_nhwc": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "utx_gemm_bf16_f32out": (i32, [vp, lng, vp, lng, vp, vp, lng, i32, i32, i32, f32, vp]),
    "utx_softmax_rows": (i32, [vp, lng, vp, lng, i32, i32, vp]),
    "utx_transpose_bf16": (i32, [vp, lng, vp, lng, i32, i32, vp]),
    "utx_vae_create": (i32, [C.POINTER(VaeConfigC), C.POINTER(vp)]),
    "utx_vae_destroy": (None, [vp]),
    "utx_vae_set_weights": (i32, [vp, C.POINTER(VaeWeightsC)]),
    "utx_vae_workspace_bytes": (C.c_size_t, [vp, i32, i32, i32, i32]),
    "utx_vae_decode": (i32, [vp, vp, i32, i32, i32, vp, vp, C.c_size_t, vp]),
    "utx_vae_encode": (i32, [vp, vp, i32, i32, i32, vp, vp, C.c_size_t, vp]),
    "utx_vae_launches": (C.c_long, [vp, i32]),
    "utx_comm_unique_id": (i32, [vp]),
    "utx_comm_init": (i32, [C.POINTER(vp), vp, i32, i32]),
    "utx_comm_destroy": (None, [vp]),
    "utx_allgather_tiles": (i32, [vp, vp, vp, C.c_size_t, vp]),
    "utx_comm_alltoall": (i32, [vp, vp, vp, C.c_size_t, vp]),
    "utx_mv_visibility_filter": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, C.c_float, C.c_float, vp, vp]),
    "utx_mvpaint_blend": (i32, [vp, vp, C.c_longlong, i32, vp, vp, vp, vp, vp]),
    "utx_knn1": (i32, [vp, i32, vp, C.c_longlong, vp, vp, vp, vp, C.c_size_t, vp]),
    "utx_knn": (i32, [vp, i32, vp, C.c_longlong, i32, vp, vp, vp, vp, C.c_size_t, vp]),
    "utx_uv_bake_workspace_bytes": (C.c_size_t, [i32, i32]),
    "utx_uv_bake_layout": (i32, [i32, i32] + [C.POINTER(C.c_size_t)] * 4),
    "utx_uv_bake_visibility": (i32, [vp, i32, vp, i32, vp, vp, i32, i32, i32, fp, fp, i32, C.POINTER(C.c_int32), vp, i32, i32, f32,
                                     vp, vp, vp, C.c_size_t, vp]),
    "utx_uv_bake_views_workspace_bytes": (C.c_size_t, [i32, i32, i32]),
    "utx_uv_bake_views_knn": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, i32, i32, vp, C.c_size_t, vp, C.c_size_t, vp]),
    "utx_uv_bake_fill": (i32, [vp, i32, i32, i32, vp, vp, C.c_size_t, vp]),
    "utx_uv_bake_finish": (i32, [vp, i32, i32, i32, vp, f32, vp, vp, C.c_size_t, vp]),
    "utx_uv_bake": (i32, [vp, i32, vp, i32, vp, vp, i32, i32, i32, fp, fp, i32, C.POINTER(C.c_int32), vp, i32, i32, f32, vp, f32,
                          fp, f32, vp, vp, vp, vp, vp, C.c_size_t, vp]),
}


class UtxError(RuntimeError):
    pass


def lib_path() -> Path:
    return _LIB_PATH


def exported_symbols():
    return sorted(_SIGS)


def load():
    """Load the shared library (built by unitex_b200.build); raises if absent -- no CPU/torch fallback exists."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise UtxError(f"{_LIB_PATH} not found: run `python -m unitex_b200.build` (nvcc, sm_100a). "
                           "unitex_b200 has no fallback path.")
        lib = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise UtxError(f"{what} failed (rc={rc}): {load().utx_last_error().decode()}")
