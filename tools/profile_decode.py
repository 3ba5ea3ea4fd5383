"""ncu helper: a few eager text-decode steps of a depth-reduced full-width Bailing-MoE after a 1552-token prefill, bracketed
by cudaProfilerStart/Stop:  ncu --profile-from-start off --metrics gpu__time_duration.sum python tools/profile_decode.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200 import ops, synthetic  # noqa: E402
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeForCausalLM  # noqa: E402

dev = torch.device("cuda:0")
layers, S = int(os.environ.get("PF_LAYERS", "4")), int(os.environ.get("PF_TOKENS", "1552"))
cfg = BailingMoeConfig(**dict(synthetic.LLM_CONFIG, num_hidden_layers=layers))
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    llm = BailingMoeForCausalLM(cfg)
torch.set_default_dtype(torch.float32)
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    for name, p in llm.named_parameters():
        if p.dim() >= 2:
            p.copy_(torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * (0.5 if name.endswith("gate.weight") else p.shape[-1] ** -0.5))
        elif "norm" in name:
            p.fill_(1.0)
        else:
            p.zero_()
emb = (torch.randn((1, S, cfg.hidden_size), generator=g, device=dev) * 0.5).to(torch.bfloat16)
pos = torch.arange(S, device=dev, dtype=torch.int32).unsqueeze(0)
cache = llm.new_cache(max_len=S + 64, max_batch=1)
hidden = llm.model.forward_tokens(emb, pos, cache, key_mask=None)
llm.use_cuda_graph = False
llm.greedy_decode(hidden[:, -1], cache, 2)  # warm-up
torch.cuda.synchronize()
torch.cuda.profiler.start()
llm.greedy_decode(hidden[:, -1], cache, 3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
