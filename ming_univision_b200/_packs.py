"""Invalidation of the kernel-ready ("packed") weight copies and of everything captured on top of them.

The product modules keep bf16 re-layouts of their parameters (stacked expert slabs, folded-LayerNorm packs, the RF
head's stacked adaLN matrix, ...) and CUDA graphs whose kernel arguments point INTO those copies.  They must be
dropped whenever the parameters change under them:

  * `.to()` / `.cuda()` / `.bfloat16()`            -> every owner overrides `_apply`
  * `load_state_dict` on the owner OR ANY ANCESTOR  -> nn.Module recurses through `_load_from_state_dict`, so an
    overridden `load_state_dict` of a sub-module is never reached; a load-state-dict POST HOOK is
    (`register_load_state_dict_post_hook` runs for every module of the recursion).

`watch(module, reset)` registers that hook.  Every reset also bumps a process-wide epoch: graph workspaces remember
the epoch they were captured under and are re-captured when it moved (a graph of a PARENT module holds pointers into
the packs of its CHILDREN, which may be reloaded on their own).
"""
from __future__ import annotations

_EPOCH = [0]


def epoch() -> int:
    return _EPOCH[0]


def bump() -> int:
    _EPOCH[0] += 1
    return _EPOCH[0]


def watch(module, reset) -> None:
    """Calls `reset()` (and bumps the epoch) after every state-dict load that reaches `module`."""

    def _hook(mod, incompatible_keys):
        reset()
        bump()

    module.register_load_state_dict_post_hook(_hook)
