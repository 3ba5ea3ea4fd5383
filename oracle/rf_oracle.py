"""CPU oracle for the rectified-flow SwiGLU head: fp32 functional restatement of
mingunivision/diff_loss_rf_swiglu.py (RectifiedFlowLoss.sample :103-181, SimpleMLPAdaLN.forward :363-385,
TimestepEmbedder :188-239, ResBlock :242-272, FinalLayer :275-292, SwiGLUFFN :14-34).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (see oracle/mingtok_oracle.py for the rules).  Pinned against the unmodified
reference module by tests/golden/make_golden_rf.py -> tests/golden/rf_*.npz (tests/test_oracle_golden.py).
`sd` uses the reference's state_dict keys below `diffloss.` (net.time_embed.mlp.0.weight, net.res_blocks.{i}...).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def timestep_embedding(t, dim=256, max_period=10000):
    """TimestepEmbedder.timestep_embedding — diff_loss_rf_swiglu.py:216-234 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def modulate(x, shift, scale):
    """diff_loss_rf_swiglu.py:184-185."""
    return x * (1 + scale) + shift


def net_forward(sd, x, t, c, prefix="net"):
    """SimpleMLPAdaLN.forward — diff_loss_rf_swiglu.py:363-385.  x [N, C], t [N] in (0, 1], c [N, Z] -> [N, C]."""
    x = _lin(sd, f"{prefix}.input_proj", x)
    temb = _lin(sd, f"{prefix}.time_embed.mlp.2", F.silu(_lin(sd, f"{prefix}.time_embed.mlp.0",
                                                               timestep_embedding(t * 1000))))
    y = temb + _lin(sd, f"{prefix}.cond_embed", c)
    i = 0
    while f"{prefix}.res_blocks.{i}.in_ln.weight" in sd:
        p = f"{prefix}.res_blocks.{i}"
        shift, scale, gate = _lin(sd, p + ".adaLN_modulation.1", F.silu(y)).chunk(3, dim=-1)
        h = F.layer_norm(x, (x.shape[-1],), sd[p + ".in_ln.weight"], sd[p + ".in_ln.bias"], 1e-6)
        h = modulate(h, shift, scale)
        x1, x2 = _lin(sd, p + ".mlp.w12", h).chunk(2, dim=-1)
        h = _lin(sd, p + ".mlp.w3", F.silu(x1) * x2)
        x = x + gate * h
        i += 1
    p = f"{prefix}.final_layer"
    shift, scale = _lin(sd, p + ".adaLN_modulation.1", F.silu(y)).chunk(2, dim=-1)
    x = modulate(F.layer_norm(x, (x.shape[-1],), None, None, 1e-6), shift, scale)
    return _lin(sd, p + ".linear", x)


def sample(sd, z, noise, num_sampling_steps=16, temperature=1.0, text_cfg=1.0, image_cfg=1.0):
    """RectifiedFlowLoss.sample — diff_loss_rf_swiglu.py:103-181 (cfg_renorm_type / time_shifting_factor are always
    None on the path, modeling_bailing_moe.py:1859-1860).  `noise` replaces the torch.randn draw: shape [1, C] when
    text_cfg != 1 (shared across rows, :117-119) else [B, C].  t runs 1 -> 1/steps, x += v / steps (:135-136, :177)."""
    B = z.shape[0]
    if text_cfg != 1.0:
        x = torch.cat([noise] * B, dim=0) * temperature
    else:
        x = noise * temperature
    steps = num_sampling_steps
    time_steps = torch.linspace(1.0, 0.0, steps + 1)[:-1]
    for t in time_steps:
        t_batch = torch.ones(B) * t
        if B == 3:
            half = x[: B // 3]
            v_all = net_forward(sd, torch.cat([half, half, half], dim=0), t_batch, z)
            v_c, v_u, v_tu = torch.split(v_all, B // 3, dim=0)
            v = v_u + image_cfg * (v_tu - v_u) + text_cfg * (v_c - v_tu)
            v = torch.cat([v, v, v], dim=0)
        elif B == 2:
            half = x[: B // 2]
            v_all = net_forward(sd, torch.cat([half, half], dim=0), t_batch, z)
            v_c, v_u = torch.split(v_all, B // 2, dim=0)
            v = v_u + text_cfg * (v_c - v_u)
            v = torch.cat([v, v], dim=0)
        else:
            v = net_forward(sd, x, t_batch, z)
        x = x + v * (1.0 / steps)
    return x
