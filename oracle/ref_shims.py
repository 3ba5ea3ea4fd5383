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


# ---------------------------------------------------------------------------------------------------------------
# Bailing-MoE LLM (mingunivision/modeling_bailing_moe.py) under transformers 5.x
# ---------------------------------------------------------------------------------------------------------------
def _legacy_cache_class():
    """A minimal stand-in for the transformers-4.x DynamicCache API the reference uses (from_legacy_cache,
    get_usable_length, key_cache / value_cache lists, seen_tokens, to_legacy_cache, get_max_length; called at
    modeling_bailing_moe.py:778, 1439-1440, 1527, 1896-1901, 1993-1997)."""
    import torch

    class LegacyDynamicCache:
        def __init__(self):
            self.key_cache, self.value_cache = [], []
            self.seen_tokens = 0

        @classmethod
        def from_legacy_cache(cls, past=None):
            c = cls()
            if past is not None:
                for k, v in past:
                    c.key_cache.append(k)
                    c.value_cache.append(v)
                if c.key_cache:
                    c.seen_tokens = c.key_cache[0].shape[-2]
            return c

        def to_legacy_cache(self):
            return tuple((k, v) for k, v in zip(self.key_cache, self.value_cache))

        def __len__(self):
            return len(self.key_cache)

        def get_seq_length(self, layer_idx=0):
            return 0 if len(self.key_cache) <= layer_idx else self.key_cache[layer_idx].shape[-2]

        def get_max_length(self):
            return None

        def get_usable_length(self, new_seq_length, layer_idx=0):
            return self.get_seq_length(layer_idx)

        def update(self, k, v, layer_idx, cache_kwargs=None):
            if layer_idx == 0:
                self.seen_tokens += k.shape[-2]
            if len(self.key_cache) <= layer_idx:
                self.key_cache.append(k)
                self.value_cache.append(v)
            else:
                self.key_cache[layer_idx] = torch.cat([self.key_cache[layer_idx], k], dim=-2)
                self.value_cache[layer_idx] = torch.cat([self.value_cache[layer_idx], v], dim=-2)
            return self.key_cache[layer_idx], self.value_cache[layer_idx]

    return LegacyDynamicCache


def build_reference_llm(llm_cfg: dict, vishead_cfg: dict, state_dict=None):
    """Builds the reference BailingMoeForCausalLM (eager attention, fp32, CPU) + vis_head + diffloss."""
    install()
    os.environ["XFORMERS_DISABLED"] = "1"
    import contextlib
    import io

    import torch
    import transformers.cache_utils as cu

    import modeling_bailing_moe as M
    from configuration_bailing_moe import BailingMoeConfig

    Legacy = _legacy_cache_class()
    M.DynamicCache = Legacy
    M.Cache = (Legacy, cu.Cache)
    cfg = BailingMoeConfig(**llm_cfg)
    cfg._attn_implementation = "eager"
    cfg.rope_scaling = None  # transformers 5 rewrites it to {'rope_type': 'default', ...}; the path uses the 1-D legacy rotary
    with contextlib.redirect_stdout(io.StringIO()):
        orig = M.BailingMoePreTrainedModel._init_weights
        M.BailingMoePreTrainedModel._init_weights = lambda self, module: None
        try:
            model = M.BailingMoeForCausalLM(cfg)
        finally:
            M.BailingMoePreTrainedModel._init_weights = orig
        model.config.rope_scaling = None
        for layer in model.model.layers:  # transformers 5 rewrites rope_scaling; force the 1-D legacy rotary
            layer.attention.config.rope_scaling = None
            layer.attention._init_rope()
        model.setup_vishead_diffloss(**vishead_cfg)
    model = model.float().eval()
    if state_dict is not None:
        model.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=True)
    return model, Legacy
