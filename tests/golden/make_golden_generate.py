"""Golden trace of the multi-round `generate` state from the LIVE reference, run in the build
container:

    python tests/golden/make_golden_generate.py   ->   tests/golden/generate_trace.json

What runs: the reference's own `MingUniVisionForConditionalGeneration.generate` (modeling_bailingmm.py:206-301: mask
concatenation across rounds, PAST_MODE KEEP / DROP padding, state hand-over) around its own
`BailingMoeForCausalLM.prepare_inputs_for_generation` / `.forward` / `.generate_image` (tiny 2-layer model, tiny RF head,
tiny MingTok, real arithmetic on the CPU).  The one piece that is NOT the reference's is the decoding loop between them:
the reference calls `transformers==4.52.4`'s `GenerationMixin.generate`, absent from this image, so
`oracle/hf_generate_oracle.greedy_generate` (a restatement of that published loop; its header says "parity unpinned")
is bound in its place.  The token stream is SCRIPTED (random tiny weights never emit `<image>` and sit on argmax
near-ties): text tokens, then the `<image>` start token — which makes the reference's forward run generate_image —, more
text, eos; a second and a third round continue behind the saved context.

Recorded per PAST_MODE and round: the returned sequence, the cache length, the three saved masks, every forward call
(tokens fed, embeddings or ids, position ids, mask length) and what generate_image received (cache length, the cond /
uncond / text-uncond masks).  tests/test_host_logic_cpu.py::test_multi_round_state_matches_the_reference_trace replays
the same script through this package's `generate` (compute stubbed) and compares."""
import contextlib
import io
import json
import os
import sys
import types
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from ming_univision_b200 import synthetic  # noqa: E402
from oracle import hf_generate_oracle, ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
EOS = 7


def script_rounds(cfg):
    """(prompt ids, uncond mask, text-uncond mask, scripted new tokens, max_new_tokens) per round."""
    img = cfg["image_start_token"]
    return [
        (list(range(10, 16)), [1, 1, 0, 0, 0, 0], [1, 1, 1, 0, 0, 1], [21, 22, 23, img, 31, 32, EOS], 32),
        (list(range(40, 44)), [0, 0, 0, 0], [0, 1, 1, 0], [img, EOS], 8),
        (list(range(50, 53)), [1, 0, 0], [1, 0, 1], [61, 62, 63], 3),          # stops on max_new_tokens, no eos
    ]


def build_reference_wrapper():
    ref_shims.install()
    for name, attrs in (("whisper", {}), ("whisper.model", {"AudioEncoder": object}), ("funasr", {}), ("funasr.models", {}),
                        ("funasr.models.sanm", {}), ("funasr.models.sanm.encoder", {"SANMEncoder": object})):
        ref_shims._stub(name, **attrs)
    import modeling_bailingmm as RM

    cfg, vh, tok_cfg = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG, synthetic.MINGTOK_TINY_CONFIG
    F_dim = tok_cfg["semantic_decoder"]["embed_dim"]
    sd = synthetic.llm_state_dict(cfg, vh, feature_dim=F_dim, seed=0)
    llm, Legacy = ref_shims.build_reference_llm(cfg, vh, None)
    llm.load_state_dict({k: v for k, v in sd.items() if not k.startswith("linear_proj.")}, strict=False)
    mingtok = ref_shims.build_reference_mingtok(tok_cfg, synthetic.mingtok_state_dict(tok_cfg, 0), fa_enable=False)
    D = cfg["hidden_size"]
    lin = torch.nn.Sequential(torch.nn.Linear(F_dim, D), torch.nn.GELU(), torch.nn.Linear(D, D)).eval()
    lin.load_state_dict({k[len("linear_proj."):]: v for k, v in sd.items() if k.startswith("linear_proj.")})
    W = object.__new__(RM.MingUniVisionForConditionalGeneration)  # (its __init__ loads ./models/MingTok-Vision from disk)
    torch.nn.Module.__init__(W)
    W.config = types.SimpleNamespace(llm_config=llm.config)
    W.vision, W.model, W.linear_proj = mingtok, llm, lin
    W.tokenizer = None
    W.past_key_values = W.past_attention_mask = None
    W.past_text_uncond_attention_mask = W.past_uncond_attention_mask = None
    return W, llm, Legacy, cfg


def run_mode(mode: str) -> list:
    os.environ["PAST_MODE"] = mode
    W, llm, Legacy, cfg = build_reference_wrapper()
    rounds = []
    state = {}

    def bound_generate(input_ids=None, **kw):  # stands where GenerationMixin.generate is called (:253-268)
        if kw.get("past_key_values") is None:
            kw["past_key_values"] = Legacy()
        kw.pop("rope_deltas", None)  # (None on this path; with it the reference would build 3-D position ids)
        return hf_generate_oracle.greedy_generate(llm, input_ids, state["max_new"], EOS, forced_tokens=state["script"],
                                                  trace=state["calls"], rope_deltas=None, **kw)

    llm.generate = bound_generate
    real_gen_image = llm.generate_image

    def spy_generate_image(**kw):
        state["images"].append({"cache_len": kw["past_key_values"].get_seq_length(),
                                "attention_mask": kw["attention_mask"][0].tolist(),
                                "uncond": kw["uncond_attention_mask"][0].tolist(),
                                "text_uncond": kw["text_uncond_attention_mask"][0].tolist()})
        return real_gen_image(**kw)

    llm.generate_image = spy_generate_image
    import modeling_bailing_moe as M
    import torchvision.transforms as T

    # the reference's tensor_to_pil moves its constants to CUDA (:84-90); same arithmetic on the CPU for this run
    M.tensor_to_pil = lambda x: T.ToPILImage()((x * 0.5 + 0.5)[0].clamp(0, 1))
    cwd = os.getcwd()
    os.chdir("/tmp")  # (the reference's forward saves every generated image as `{prefix}.png` in the working directory)
    try:
        for ids, un, tun, script, max_new in script_rounds(cfg):
            state.update(script=list(script), max_new=max_new, calls=[], images=[])
            t = lambda v: torch.tensor([v], dtype=torch.long)  # noqa: E731
            torch.manual_seed(11)
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                seq = W.generate(input_ids=t(ids), attention_mask=torch.ones((1, len(ids)), dtype=torch.long),
                                 uncond_attention_mask=t(un), text_uncond_attention_mask=t(tun), use_cache=True,
                                 output_image_prefix="/tmp/_golden_generate")
            rounds.append({"prompt": ids, "uncond": un, "text_uncond": tun, "script": script, "max_new_tokens": max_new,
                           "sequence": seq[0].tolist(), "cache_len": W.past_key_values.get_seq_length(),
                           "past_attention_mask": W.past_attention_mask[0].tolist(),
                           "past_uncond_attention_mask": W.past_uncond_attention_mask[0].tolist(),
                           "past_text_uncond_attention_mask": W.past_text_uncond_attention_mask[0].tolist(),
                           "forward_calls": state["calls"], "generate_image_calls": state["images"],
                           "cache_rows": int(W.past_key_values.key_cache[0].shape[0])})
    finally:
        os.chdir(cwd)
    return rounds


def main():
    torch.set_num_threads(8)
    out = {"eos": EOS, "n_image_tokens": synthetic.LLM_TINY_CONFIG["num_image_tokens_for_gen"],
           "image_start_token": synthetic.LLM_TINY_CONFIG["image_start_token"],
           "hf_loop": "oracle/hf_generate_oracle.py (restated transformers==4.52.4 greedy loop; parity unpinned)"}
    for mode in ("DROP", "KEEP"):
        out[mode] = run_mode(mode)
        for i, r in enumerate(out[mode]):
            print(mode, "round", i, "sequence", len(r["sequence"]), "cache", r["cache_len"], "forward calls",
                  [(c["cache_len"], c["fed"], c["used_embeds"]) for c in r["forward_calls"]], "images",
                  [g["cache_len"] for g in r["generate_image_calls"]])
    with open(os.path.join(OUT, "generate_trace.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
