"""Import shims that let the UNMODIFIED reference (read-only, /root/reference) run in this container on CPU.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture generation, run in the build container
where /root/reference exists) and by the optional live-reference tests.  Nothing here is imported by the product
package, by bench.py's GPU arm, or on the GPU box (where /root/reference does not exist).

Shims: an `omegaconf` stub (imported but unused at runtime, mingtok/modeling_mingtok.py:2,
mingtok/utils/processor.py:2), `transformers.utils.import_utils.is_torch_fx_available` (modeling_bailing_moe.py:61),
and stubs for `funasr` / `whisper` (modeling_bailingmm.py:22, modeling_utils.py:17).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MING_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mingtok"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    try:
        import omegaconf  # noqa: F401
    except ImportError:
        _stub("omegaconf", MISSING="???", OmegaConf=type("OmegaConf", (), {}))
    import transformers.utils.import_utils as iu

    if not hasattr(iu, "is_torch_fx_available"):
        iu.is_torch_fx_available = lambda: False
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mu = os.path.join(REFERENCE_ROOT, "mingunivision")
    if mu not in sys.path:
        sys.path.insert(0, mu)


def build_reference_mingtok(config: dict, state_dict=None, fa_enable=False):
    """Builds the reference MingTok (mingtok/modeling_mingtok.py:97) on CPU in fp32 with `config` (the dict layout of
    mingtok/config/config_mingtok.json) and loads `state_dict` (strict).  fa_enable=False selects the eager attention
    classes, which are the only ones whose KV-cache path is correct on CPU."""
    install()
    import copy

    import torch
    from mingtok.modeling_mingtok import MingTok, MingTokConfig

    cfg = copy.deepcopy(config)
    for k in ("low_level_encoder", "semantic_decoder", "pixel_decoder"):
        cfg[k] = dict(cfg[k], fa_enable=fa_enable)
    orig_init = MingTok._init_weights if hasattr(MingTok, "_init_weights") else None
    MingTok._init_weights = lambda self, module: None  # weights come from the state_dict; skip the ~95 s HF init
    try:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            model = MingTok(MingTokConfig(**cfg))
    finally:
        if orig_init is not None:
            MingTok._init_weights = orig_init
    model = model.float().eval()
    if state_dict is not None:
        model.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=True)
    return model
