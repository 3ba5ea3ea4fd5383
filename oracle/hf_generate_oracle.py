"""TEST INFRASTRUCTURE ONLY — the greedy decoding loop of Hugging Face `GenerationMixin`, restated.

The reference reaches its text / image generation loop only through `transformers.GenerationMixin.generate`
(mingunivision/modeling_bailingmm.py:253-268 calls `self.model.generate(...)`; requirements.txt:23 pins
`transformers==4.52.4`).  That dependency is absent here — this image has transformers 5.5, whose loop no longer matches
the reference's `prepare_inputs_for_generation` — so the published algorithm of 4.52.4's `generation/utils.py`
(`generate` -> `_prepare_model_inputs` -> `_get_initial_cache_position` -> `_sample` with `do_sample=False` ->
`_update_model_kwargs_for_generation`) is restated below and used to drive the LIVE reference's own functions
(`BailingMoeForCausalLM.prepare_inputs_for_generation` :1966-2066, `.forward` :1677-1843, and around them
`MingUniVisionForConditionalGeneration.generate` :206-301) when the golden trace of the multi-round state is generated
(tests/golden/make_golden_generate.py).

PARITY UNPINNED for the loop itself: it is written from the published algorithm, not checked against a 4.52.4 install.
What it anchors is everything the reference's own code does around it — which tokens are fed, with which positions and
masks, how the cache and the three masks evolve over rounds.

Restated steps (batch 1, greedy):
  1. both `input_ids` and `inputs_embeds` are given: `inputs_embeds` feeds the first step, `input_ids` is the running
     sequence (`_prepare_model_inputs`, decoder-only branch);
  2. `cache_position = arange(inputs_embeds.shape[1])[past_length:]`, `past_length` = what the passed cache already holds
     (`_get_initial_cache_position`) — EMPTY in later rounds, where the cache is longer than the round's prompt;
  3. loop: `model_inputs = prepare_inputs_for_generation(input_ids, **model_kwargs)`; `outputs = model(**model_inputs,
     return_dict=True)`; next token = argmax of the last position's logits (or the next scripted token: the tests force
     the token stream, random tiny weights sit on argmax near-ties); `input_ids = cat(input_ids, next)`;
     `model_kwargs`: `past_key_values <- outputs.past_key_values`, `attention_mask <- cat(attention_mask, 1)`,
     `cache_position <- cache_position[-1:] + 1` (`_update_model_kwargs_for_generation`);
  4. stop after the eos token was appended or after `max_new_tokens` new tokens.
"""
from __future__ import annotations

import types

import torch


def greedy_generate(model, input_ids, max_new_tokens: int, eos_token_id, forced_tokens=None, trace=None, **model_kwargs):
    """Drives `model.prepare_inputs_for_generation` / `model.__call__` as GenerationMixin's greedy search does.
    `forced_tokens`: list popped from the front instead of the argmax while it lasts.  `trace`: list receiving one dict
    per forward call (what was fed).  Returns an object with `.sequences` and `.past_key_values` (return_dict_in_generate)."""
    model_kwargs = {k: v for k, v in model_kwargs.items() if k != "return_dict_in_generate"}
    embeds = model_kwargs.get("inputs_embeds")
    cache = model_kwargs.get("past_key_values")
    if cache is None:
        raise ValueError("pass the (possibly empty) cache object: the reference's forward needs its legacy cache class")
    n_first = embeds.shape[1] if embeds is not None else input_ids.shape[1]
    past_length = cache.get_seq_length()
    model_kwargs["cache_position"] = torch.arange(n_first)[past_length:]
    eos = set(eos_token_id if isinstance(eos_token_id, (list, tuple)) else [eos_token_id])
    forced = list(forced_tokens or [])
    new = 0
    while new < max_new_tokens:
        cache_len = model_kwargs["past_key_values"].get_seq_length()
        inputs = model.prepare_inputs_for_generation(input_ids, **model_kwargs)
        if trace is not None:
            fed = inputs["inputs_embeds"].shape[1] if "inputs_embeds" in inputs else inputs["input_ids"].shape[1]
            first = None if "inputs_embeds" in inputs else int(inputs["input_ids"][0, 0])
            trace.append({"cache_len": cache_len, "fed": fed, "used_embeds": "inputs_embeds" in inputs, "first_id": first,
                          "position_ids": inputs["position_ids"][0].tolist(),
                          "attention_mask_len": int(inputs["attention_mask"].shape[1]),
                          "attention_mask_sum": int(inputs["attention_mask"].sum())})
        outputs = model(**inputs, return_dict=True)
        model_kwargs["past_key_values"] = outputs.past_key_values
        am = model_kwargs["attention_mask"]
        model_kwargs["attention_mask"] = torch.cat((am, am.new_ones((am.shape[0], 1))), dim=-1)
        model_kwargs["cache_position"] = model_kwargs["cache_position"][-1:] + 1
        nxt = forced.pop(0) if forced else int(outputs.logits[0, -1].argmax())
        input_ids = torch.cat((input_ids, torch.tensor([[nxt]], dtype=input_ids.dtype)), dim=1)
        new += 1
        if nxt in eos:
            break
    return types.SimpleNamespace(sequences=input_ids, past_key_values=model_kwargs["past_key_values"])
