"""CPU: the C-ABI library builds, loads without a driver, exports every symbol include/unitex_b200.h declares,
and fails loudly (error code + message) instead of falling back."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT / "include" / "unitex_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(utx_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/unitex_b200.h but not exported"
    from unitex_b200 import _lib
    assert set(_lib.exported_symbols()) == set(names)


def test_error_reporting_no_fallback(lib):
    assert lib.utx_version() >= 100
    rc = lib.utx_gemm_bf16(None, 0, None, 0, None, None, 0, 128, 128, 64, 0, None, None, 0, None)
    assert rc != 0 and b"null pointer" in lib.utx_last_error()
    from unitex_b200 import _lib
    with pytest.raises(_lib.UtxError):
        _lib.check(rc, "utx_gemm_bf16")


def test_ops_reject_cpu_tensors(lib):
    from unitex_b200 import _lib, ops
    with pytest.raises(_lib.UtxError):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16))


def test_flux_handle_and_packing_on_cpu(lib):
    """The host-side packing (state dict -> stacked qkv / modulation rows) is pure bookkeeping: check it on CPU."""
    from oracle import flux_dit as fd
    from unitex_b200.flux import FluxConfig, FluxTransformer
    ocfg = fd.FluxConfig.tiny(2, 2)
    P = fd.init_params(ocfg, 0, norm_weight_std=0.1)
    cfg = FluxConfig(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256,
                     pooled_projection_dim=64)
    eng = FluxTransformer(cfg, device="cpu").load_state_dict(P)
    D = cfg.inner_dim
    bf = lambda t: t.to(torch.bfloat16)
    assert torch.equal(eng.T["w_d1.qkv_img"][D:2 * D], bf(P["transformer_blocks.1.attn.to_k.weight"]))
    assert torch.equal(eng.T["b_d0.qkv_txt"][2 * D:], bf(P["transformer_blocks.0.attn.add_v_proj.bias"]))
    assert torch.equal(eng.T["w_s1.qkvmlp"][3 * D:], bf(P["single_transformer_blocks.1.proj_mlp.weight"]))
    assert torch.equal(eng.T["w_s0.out"], bf(P["single_transformer_blocks.0.proj_out.weight"]))
    assert torch.equal(eng.T["d1.rms_k_txt"], bf(P["transformer_blocks.1.attn.norm_added_k.weight"]))
    mod = eng.T["w_mod"]
    assert mod.shape == (cfg.n_mod_rows, D)
    assert torch.equal(mod[6 * D:12 * D], bf(P["transformer_blocks.0.norm1_context.linear.weight"]))
    assert torch.equal(mod[24 * D:27 * D], bf(P["single_transformer_blocks.0.norm.linear.weight"]))
    assert torch.equal(mod[-2 * D:], bf(P["norm_out.linear.weight"]))
    assert eng.lib.utx_flux_workspace_bytes(eng._handle, 128, 192) > (128 + 192) * D * 2 * 10
    # calling the engine without prepare() is an error, not a silent no-op
    assert eng.lib.utx_flux_forward(eng._handle, 1, 0.5, 3.5, 1, None) != 0


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: a cgo / JNI / FFI consumer must be able to include the header as C99 (no C++ constructs, no
    torch or CUDA types in the signatures)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    hdr = ROOT / "include" / "unitex_b200.h"
    src = tmp_path / "use.c"
    src.write_text('#include "unitex_b200.h"\nint main(void) { return utx_version() < 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-I{hdr.parent}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    txt = hdr.read_text()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S).lower() and "cudaStream_t" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def test_python_bindings_match_header():
    """Every prototype of the header and its ctypes signature in unitex_b200/_lib.py take the same number of arguments of the same class
    (pointer / int / 64-bit int / size_t / float) -- a binding that drifts from the header corrupts the call silently."""
    from unitex_b200 import _lib
    txt = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "unitex_b200.h").read_text(), flags=re.S)
    protos = dict(re.findall(r"\b(utx_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S))
    assert set(protos) == set(_lib._SIGS)
    for name, args in protos.items():
        args = " ".join(args.split())
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(_lib._SIGS[name][1]), f"{name}: header takes {n} arguments, _lib.py binds {len(_lib._SIGS[name][1])}"
        if n == 0:
            continue
        for i, (decl, ct) in enumerate(zip(args.split(","), _lib._SIGS[name][1])):
            assert _c_class(decl) == _ctypes_class(ct), f"{name} argument {i}: header `{decl.strip()}` vs binding {ct}"


def _c_class(decl):
    decl = decl.strip()
    if "*" in decl:
        return "ptr"
    words = decl.replace("const", " ").replace("unsigned", " ").split()
    words = words[:-1] if len(words) > 1 else words          # drop the parameter name
    t = " ".join(words)
    return {"int": "i32", "int32_t": "i32", "float": "f32", "double": "f64", "long": "i64", "long long": "i64", "size_t": "size"}[t]   # LP64: long = long long


def _ctypes_class(ct):
    if ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, "contents") or issubclass(ct, ctypes._Pointer):
        return "ptr"
    assert ctypes.sizeof(ctypes.c_long) == 8
    return {ctypes.c_int: "i32", ctypes.c_float: "f32", ctypes.c_double: "f64", ctypes.c_long: "i64", ctypes.c_longlong: "i64",
            ctypes.c_size_t: "size"}[ct]


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors of the header's structs (weights / config tables handed across the ABI by pointer) have the same
    field names, order, offsets and sizes as the C declarations -- measured by compiling the header with gcc."""
    import shutil
    import subprocess
    from unitex_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = {"utx_flux_config": _lib.FluxConfigC, "utx_double_block": _lib.DoubleBlockC, "utx_single_block": _lib.SingleBlockC,
             "utx_flux_weights": _lib.FluxWeightsC, "utx_vae_config": _lib.VaeConfigC, "utx_vae_resnet": _lib.VaeResnetC,
             "utx_vae_attn": _lib.VaeAttnC, "utx_vae_mid": _lib.VaeMidC, "utx_vae_weights": _lib.VaeWeightsC}
    txt = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "unitex_b200.h").read_text(), flags=re.S)
    declared = set(re.findall(r"typedef struct (utx_[a-z_]+) \{", txt))
    assert declared == set(pairs), declared ^ set(pairs)           # a new struct in the header needs a mirror and an entry here
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "unitex_b200.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    r = subprocess.run(["gcc", "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr                             # a field the mirror names but the header lacks fails here
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(ct), cname
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), txt, flags=re.S).group(1)
        c_fields = [f for d in body.split(";") for f in re.findall(r"\**\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[\d+\])?\s*(?:,|$)", d.strip().split(" ", 1)[-1] if d.strip() else "")]
        assert [n for n, _ in ct._fields_] == [f for f in c_fields if f not in ("const", "void", "float", "int")], cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"
