from .modeling_mingtok import MingTok, MingTokConfig, MingTokKVCache  # noqa: F401
