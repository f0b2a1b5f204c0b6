"""Compatibility shims that let the *reference* modules under /root/reference import in this
container (transformers 5.5.0 instead of the pinned 4.57.3; no librosa/soundfile/sox/onnxruntime).

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py to pin
the oracle against the reference's own module forwards.  Never imported
by the product path; /root/reference does not exist on the GPU box.
"""
import os
import sys
import types
import importlib.machinery

REFERENCE_ROOT = os.environ.get("QWEN3TTS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "qwen_tts"))


_installed = False


def install():
    """Idempotently install the three probe-only shims, then make `qwen_tts` importable."""
    global _installed
    if _installed:
        return
    import torch
    import transformers  # noqa: F401
    # resolve lazies before stub modules appear (transformers probes librosa.__spec__)
    from transformers import AutoConfig, AutoModel, AutoProcessor, AutoFeatureExtractor  # noqa: F401
    import transformers.utils.generic as G

    _orig = G.check_model_inputs

    def _check_model_inputs(func=None, **kw):
        if func is not None:
            return _orig(func, **kw)
        return lambda f: _orig(f, **kw)

    G.check_model_inputs = _check_model_inputs

    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS

    def _default_rope(config, device=None, seq_len=None, **kw):
        base = getattr(config, "rope_theta", None) or 10000.0
        dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
        inv = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / dim))
        return inv, 1.0

    ROPE_INIT_FUNCTIONS.setdefault("default", _default_rope)

    def _stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    lib = _stub("librosa")
    lib.filters = _stub("librosa.filters", mel=lambda **k: None)
    _stub("soundfile")
    _stub("sox")
    _stub("onnxruntime")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
