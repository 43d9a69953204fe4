"""Shared by tests/golden/make_reference_golden.py (which drives the REFERENCE's top-level pipeline.py glue with these) and by
the tests that drive this repo's drop-in with the same fakes: seeded inputs and a stand-in for the FLUX pipeline object."""
import types

import numpy as np


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def glue_inputs():
    """Seeded inputs of the glue vectors (regenerated, not stored: 4.7 MB each)."""
    rng = np.random.default_rng(11)
    normal = rng.integers(0, 256, (1024, 1536, 3), dtype=np.uint8)
    ccm = rng.integers(0, 256, (1024, 1536, 3), dtype=np.uint8)
    ref = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    return normal, ccm, ref


class FakeFlux:
    """Stands in for the FLUX pipeline object inside infer_mv: records every call, returns a fixed function of the control image."""
    def __init__(self):
        self.calls, self.adapters, self._num_inference_steps = [], [], 28

    def set_adapters(self, adapter_names=None, adapter_weights=None):
        self.adapters.append((list(adapter_names), [float(w) for w in adapter_weights]))

    def __call__(self, **kw):
        from PIL import Image
        c = np.array(kw["control_image"])
        self.calls.append({k: v for k, v in kw.items() if k not in ("generator",)})
        out = 255 - c if len(self.calls) == 1 else np.roll(c, 7, axis=1)
        return types.SimpleNamespace(images=[Image.fromarray(out)])


def reference_rgba():
    """A 1024^2 RGBA reference image with an off-centre elliptical matte (soft edge) over a colour pattern."""
    yy, xx = np.mgrid[0:1024, 0:1024].astype(np.float32)
    rgb = np.stack([127 + 100 * np.sin(xx / 37.0), 127 + 100 * np.cos(yy / 29.0), 127 + 90 * np.sin((xx + yy) / 53.0)], -1)
    d = np.sqrt(((xx - 430.0) / 260.0) ** 2 + ((yy - 560.0) / 330.0) ** 2)
    alpha = np.clip((1.0 - d) * 12.0, 0.0, 1.0) * 255.0
    return np.concatenate([np.clip(rgb, 0, 255), alpha[..., None]], -1).astype(np.uint8)
