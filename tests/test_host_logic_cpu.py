"""Host-side logic of the multi-round `generate` (modeling_bailingmm.py:206-301) on CPU: the CUDA-backed pieces
(embedding lookup, LLM forward, greedy decoder, generate_image) are replaced by shape-faithful stand-ins, so what is
checked here is the bookkeeping the reference does in Python — mask concatenation across rounds, the PAST_MODE KEEP / DROP
padding rules, the cache-length contract ("the last emitted token is not part of the cached context"), the `<image>`
hand-off to generate_image and the returned sequences."""
import os

import pytest
import torch

from ming_univision_b200 import synthetic
from ming_univision_b200.mingtok import MingTokConfig
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig
from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration


@pytest.fixture()
def stubbed():
    cfg, vh, tok = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG, synthetic.MINGTOK_TINY_CONFIG
    m = MingUniVisionForConditionalGeneration(BailingMoeConfig(**cfg), MingTokConfig(**tok), vh)
    llm = m.model
    D, n_tok = cfg["hidden_size"], llm.config.num_image_tokens_for_gen
    log = {"prefill": [], "image": []}
    script = {"tokens": []}

    llm.model.embed = lambda ids: torch.zeros(tuple(ids.shape) + (D,))

    def forward_tokens(emb, pos, cache, key_mask=None, image_mask=None, t_dev=None):
        B, S, _ = emb.shape
        assert int(pos[0, 0]) == cache.seq_len, "the round's prompt must be appended right behind the cached context"
        log["prefill"].append((cache.seq_len, S))
        cache.seq_len += S
        return torch.zeros((B, S, D))

    def greedy_decode(last, cache, max_new_tokens, stop_ids=()):
        out = []
        while script["tokens"] and len(out) < max_new_tokens:
            t = script["tokens"].pop(0)
            out.append(t)
            cache.seq_len += 1
            if t in stop_ids:
                break
        if out:
            cache.seq_len -= 1
        return out

    def generate_image(input_embeds, past_key_values, attention_mask, uncond_attention_mask,
                       text_uncond_attention_mask, **kw):
        assert attention_mask.shape[1] == past_key_values.seq_len + 1
        log["image"].append((past_key_values.seq_len, uncond_attention_mask.clone(), text_uncond_attention_mask.clone()))
        past_key_values.seq_len += n_tok + 1
        return torch.zeros((1, 3, 8, 8)), torch.zeros((2, 1, D)), None

    llm.model.forward_tokens = forward_tokens
    llm.greedy_decode = greedy_decode
    llm.generate_image = generate_image
    return m, llm, log, script, n_tok


def _ref_masks(mode, am, un, tun, cache_len):
    """modeling_bailingmm.py:272-299 restated."""
    pad1 = torch.ones((1, cache_len - am.shape[1]), dtype=am.dtype)
    pad0 = torch.zeros_like(pad1)
    if mode == "KEEP":
        return torch.cat((am, pad1), 1), torch.cat((tun, pad1), 1), torch.cat((un, pad0), 1)
    return torch.cat((am, pad1), 1), torch.cat((am, pad1), 1), torch.cat((am, pad0), 1)


@pytest.mark.parametrize("mode", ["DROP", "KEEP"])
def test_multi_round_mask_bookkeeping(stubbed, mode, monkeypatch):
    m, llm, log, script, n_tok = stubbed
    monkeypatch.setenv("PAST_MODE", mode)
    eos, img_tok = 7, llm.config.image_start_token
    # round 1: 6 prompt tokens, the model answers 3 text tokens, asks for an image, then 2 more tokens and EOS
    ids1 = torch.arange(10, 16).view(1, -1)
    un1 = torch.tensor([[1, 1, 0, 0, 0, 0]], dtype=torch.int32)
    tun1 = torch.tensor([[1, 1, 1, 0, 0, 1]], dtype=torch.int32)
    script["tokens"] = [21, 22, 23, img_tok, 31, 32, eos]
    seq1 = m.generate(ids1, uncond_attention_mask=un1, text_uncond_attention_mask=tun1, max_new_tokens=32,
                      eos_token_id=eos)
    assert seq1[0].tolist() == list(range(10, 16)) + [21, 22, 23, img_tok, 31, 32, eos]
    assert log["prefill"] == [(0, 6)]
    # the image step starts right behind prompt + 3 fed tokens; it sees the round's uncond masks unchanged
    assert log["image"][0][0] == 6 + 3 and torch.equal(log["image"][0][1], un1) and torch.equal(log["image"][0][2], tun1)
    L1 = 6 + 3 + (n_tok + 1) + 3 - 1  # prompt, 3 tokens, <image> + visual tokens, 31 32 eos minus the unfed last one
    assert m.past_key_values.seq_len == L1 and len(m.generated_images) == 1
    am1 = torch.ones((1, 6), dtype=torch.int32)
    exp = _ref_masks(mode, am1, un1, tun1, L1)
    assert torch.equal(m.past_attention_mask, exp[0])
    assert torch.equal(m.past_text_uncond_attention_mask, exp[1])
    assert torch.equal(m.past_uncond_attention_mask, exp[2])
    # round 2: 4 new prompt tokens appended behind the cached context; masks concatenate with the stored ones
    ids2 = torch.arange(40, 44).view(1, -1)
    un2 = torch.zeros((1, 4), dtype=torch.int32)
    tun2 = torch.tensor([[0, 1, 1, 0]], dtype=torch.int32)
    script["tokens"] = [img_tok, eos]
    seq2 = m.generate(ids2, uncond_attention_mask=un2, text_uncond_attention_mask=tun2, max_new_tokens=8, eos_token_id=eos)
    assert seq2[0].tolist() == [40, 41, 42, 43, img_tok, eos]
    assert log["prefill"][1] == (L1, 4)
    assert torch.equal(log["image"][1][1], torch.cat((exp[2], un2), 1))      # accumulated uncond mask reaches generate_image
    assert torch.equal(log["image"][1][2], torch.cat((exp[1], tun2), 1))
    L2 = L1 + 4 + (n_tok + 1) + 1 - 1
    assert m.past_key_values.seq_len == L2
    exp2 = _ref_masks(mode, torch.cat((exp[0], torch.ones((1, 4), dtype=torch.int32)), 1), torch.cat((exp[2], un2), 1),
                      torch.cat((exp[1], tun2), 1), L2)
    assert torch.equal(m.past_attention_mask, exp2[0])
    assert torch.equal(m.past_text_uncond_attention_mask, exp2[1])
    assert torch.equal(m.past_uncond_attention_mask, exp2[2])
    m.reset_inner_state()
    assert m.past_key_values is None and m.past_uncond_attention_mask is None


def test_generate_rejects_misuse(stubbed):
    m, llm, log, script, n_tok = stubbed
    with pytest.raises(ValueError):
        m.generate(torch.zeros((2, 4), dtype=torch.long))                       # batch 1 only (:1865)
    with pytest.raises(NotImplementedError):
        m.generate(torch.zeros((1, 4), dtype=torch.long), attention_mask=torch.tensor([[0, 1, 1, 1]]))  # left padding
    m.reset_inner_state()
    script["tokens"] = [5]
    m.generate(torch.zeros((1, 4), dtype=torch.long), max_new_tokens=1, eos_token_id=99)
    with pytest.raises(ValueError):  # a later round whose masks do not line up with the cached context
        m.past_attention_mask = m.past_attention_mask[:, :-1]
        m.generate(torch.zeros((1, 3), dtype=torch.long), max_new_tokens=1)


def test_committed_bench_line_keeps_the_contract():
    """The newest committed bench line (profiles/r02*_bench.json, else r01e_bench.json — lines bench.py printed on a B200)
    must carry every key of the measurement contract (metric / value / e2e / roofline / cpu_baseline / clocks /
    gpu_launches) with consistent numbers — a guard against editing bench.py out of the contract."""
    import glob
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = sorted(glob.glob(os.path.join(root, "profiles", "r02*_bench.json"))) or \
        [os.path.join(root, "profiles", "r01e_bench.json")]
    with open(cands[-1]) as f:
        d = json.load(f)
    with open(os.path.join(root, "BASELINE.json")) as f:
        base = json.load(f)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"].split(" (")[0] in base["metric"] and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and base["published"] == {}          # no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    tokens = d["config"]["tokens_per_step"]
    assert abs(d["value"] - tokens / d["ms_per_step"] * 1e3) / d["value"] < 1e-6      # value == units / time
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb


def test_reference_arm_prints_the_same_metric():
    """The live `bench.py --impl reference` (CPU, one bounded sample step: the round with a 2-layer true-width LLM and
    2 AR steps, ~1 min here): same metric / unit / workload as the GPU arm, its own keys, steps = the steps it timed."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["steps"] == 1 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["extrapolated"] is True
    assert line["config"]["workload"] == bench.WORKLOAD and line["config"]["tokens_per_step"] == bench.TOKENS_PER_STEP
    full = line["config"]["extrapolated_full_round_ms"] / 1e3
    assert abs(line["value"] - bench.TOKENS_PER_STEP / full) / line["value"] < 1e-9


def test_state_dict_load_through_the_parent_drops_every_pack():
    """nn.Module.load_state_dict recurses through `_load_from_state_dict`, so a sub-module's own
    `load_state_dict` override is never reached when the PARENT wrapper loads a checkpoint.  Every pack owner therefore
    registers a load-state-dict post hook (ming_univision_b200/_packs.py); captured-graph workspaces are keyed on the
    pack epoch the hooks bump."""
    from ming_univision_b200 import _packs, synthetic
    from ming_univision_b200.mingtok import MingTokConfig
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig
    from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration

    m = MingUniVisionForConditionalGeneration(BailingMoeConfig(**synthetic.LLM_TINY_CONFIG),
                                              MingTokConfig(**synthetic.MINGTOK_TINY_CONFIG),
                                              synthetic.VISHEAD_TINY_CONFIG)
    llm = m.model
    moe = llm.model.layers[1].mlp

    def poison():
        llm._pk = llm.model._pk = moe._pk = m.linear_proj._pk = "stale"
        llm.diffloss._packed = m.vision._packed = "stale"
        llm._gen_ws, llm._txt_ws, llm.diffloss._graphs = {"k": 1}, {"k": 1}, {"k": 1}

    def clean():
        return (llm._pk is None and llm.model._pk is None and moe._pk is None and m.linear_proj._pk is None
                and llm.diffloss._packed is None and m.vision._packed is None
                and llm._gen_ws == {} and llm._txt_ws == {} and llm.diffloss._graphs == {})

    poison()
    e0 = _packs.epoch()
    m.load_state_dict(m.state_dict(), strict=True)          # through the top-level wrapper
    assert clean() and _packs.epoch() > e0
    poison()
    e1 = _packs.epoch()
    llm.diffloss.load_state_dict(llm.diffloss.state_dict())  # a child on its own: its packs go, the epoch moves
    assert llm.diffloss._packed is None and llm.diffloss._graphs == {} and _packs.epoch() > e1
    poison()
    m.float()                                                # _apply path (.to / .float / .bfloat16)
    assert clean()


def test_flat_module_names_of_the_reference_resolve_to_the_package():
    """The reference's scripts import `modeling_bailingmm`, `mingunivisioninfer`, `mingtok.modeling_mingtok`, ... as
    top-level modules (mingunivisioninfer.py:1-7, modeling_bailingmm.py:25-28); install_flat_modules() makes exactly those
    import lines work against this package (run in a fresh interpreter so no reference checkout is involved)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import ming_univision_b200 as pkg\n"
        "names = pkg.install_flat_modules()\n"
        "from modeling_bailingmm import MingUniVisionForConditionalGeneration\n"
        "from modeling_bailing_moe import BailingMoeForCausalLM, BailingMoeSparseMoeBlock, BailingMoeConfig\n"
        "from diff_loss_rf_swiglu import RectifiedFlowLoss\n"
        "from mingtok.modeling_mingtok import MingTok, MingTokConfig\n"
        "from mingtok.utils.processor import CenterCropProcessor\n"
        "from mingunivisioninfer import MingUniVisionInfer\n"
        "from image_processing_bailingmm import BailingMMImageProcessor, smart_resize\n"
        "from processing_bailingmm import BailingMMProcessor, MingTokUndProcessor, MingTokCenterCropProcessor\n"
        "assert MingTok.__module__.startswith('ming_univision_b200.') and len(names) == 10, names\n"
        "print('flat ok')\n" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "flat ok" in r.stdout, r.stdout + r.stderr


def test_facade_round_trip_with_the_processor(stubbed, monkeypatch, tmp_path):
    """MingUniVisionInfer.generate (mingunivisioninfer.py:82-117) end to end on the host: this package's processor (chat
    template -> image fetch -> transform -> placeholder expansion -> ids + CFG masks) feeding the wrapper's `generate`
    with the CUDA-backed pieces stubbed.  A text-to-image round, then an in-context EDIT round with an input image: the
    masks the processor built reach generate_image behind the saved context, the image features land on the
    `<imagePatch>` positions, and the decoded answer is the text after the prompt."""
    import numpy as np
    import torchvision.transforms as T
    from PIL import Image

    from ming_univision_b200.mingunivisioninfer import MingUniVisionInfer
    from ming_univision_b200.processing_bailingmm import BailingMMProcessor
    from test_processing_cpu import byte_tokenizer

    m, llm, log, script, n_tok = stubbed
    tok = byte_tokenizer()
    ids_of = lambda s: tok.convert_tokens_to_ids(s)  # noqa: E731
    cfg = llm.config
    cfg.image_patch_token, cfg.image_start_token, cfg.pad_token_id = ids_of("<imagePatch>"), ids_of("<image>"), ids_of("<|endoftext|>")
    half = [0.5, 0.5, 0.5]
    tf = T.Compose([T.Resize(64), T.CenterCrop(64), T.ToTensor(), T.Normalize(half, half)])
    proc = BailingMMProcessor(tokenizer=tok, vis_processor=tf, gen_processor=tf)
    D = cfg.hidden_size
    seen = {}

    def extract_image_feature(pixel_values, grid_thw=None):
        seen["pixels"] = (tuple(pixel_values.shape), pixel_values.dtype)
        return torch.ones((int(grid_thw.prod()), D))

    scattered = {}
    real_wrap = m.prompt_wrap_vision

    def prompt_wrap_vision(input_ids, emb, feats, image_token_id=None):
        out, mask = real_wrap(input_ids, emb, feats, image_token_id)
        scattered["mask"] = mask.clone()
        return out, mask

    m.extract_image_feature = extract_image_feature
    m.prompt_wrap_vision = prompt_wrap_vision
    monkeypatch.delenv("PAST_MODE", raising=False)
    agent = MingUniVisionInfer("unused", model=m, processor=proc, tokenizer=tok)
    eos, img_tok = cfg.pad_token_id, cfg.image_start_token
    # ---- round 1: text -> image
    script["tokens"] = [img_tok] + tok.encode("done", add_special_tokens=False) + [eos]
    msg1 = [{"role": "HUMAN", "content": [{"type": "text", "text": "Generate a corgi."}]}]
    answer = agent.generate(msg1, max_new_tokens=32, output_image_prefix=str(tmp_path / "gen"))
    assert (tmp_path / "gen.png").is_file()                       # the generated image is saved as `{prefix}.png` (:1788-1796)
    assert answer == "done"                                        # special tokens (<image>, <|endoftext|>) are skipped
    text1 = proc.apply_chat_template(msg1)
    enc1 = proc(text=[text1])
    S1 = enc1["input_ids"].shape[1]
    assert log["prefill"] == [(0, S1)]
    pos, un, tun = log["image"][0]
    assert pos == S1
    assert torch.equal(un.long(), enc1["uncond_attention_mask"]) and torch.equal(tun.long(), enc1["text_uncond_attention_mask"])
    body = len(tok.encode("<role>HUMAN</role>", add_special_tokens=False))
    tail = len(tok.encode("<role>ASSISTANT</role>", add_special_tokens=False))
    assert un[0].tolist() == [1] * body + [0] * (S1 - body - tail) + [1] * tail   # the prompt's body is hidden from the uncond row
    L1 = m.past_key_values.seq_len
    assert L1 == S1 + (n_tok + 1) + len(tok.encode("done", add_special_tokens=False)) + 1 - 1
    # ---- round 2: edit with an input image, behind the saved context
    img = Image.fromarray(np.random.default_rng(0).integers(0, 256, (80, 120, 3), dtype=np.uint8))
    msg2 = [{"role": "HUMAN", "content": [{"type": "image", "image": img}, {"type": "text", "text": "add a hat"}]}]
    script["tokens"] = [img_tok, eos]
    agent.generate(msg2, max_new_tokens=8, for_edit=True, output_image_prefix=str(tmp_path / "edit"))
    assert seen["pixels"] == ((1, 3, 64, 64), torch.bfloat16)      # the facade hands bf16 pixels to the model (:104-105)
    n_patch = (64 // m.vision.patch_size) ** 2
    side = 64 // m.vision.patch_size
    text2 = proc._expand_image_tokens([proc.apply_chat_template(msg2)], torch.tensor([[1, side, side]]))[0]
    enc2 = proc.tokenize([text2])
    S2 = enc2["input_ids"].shape[1]
    assert int(scattered["mask"].sum()) == n_patch == text2.count("<imagePatch>")
    assert log["prefill"][1] == (L1, S2)
    pos2, un2, tun2 = log["image"][1]
    assert pos2 == L1 + S2 and un2.shape[1] == L1 + S2
    # DROP mode: the saved context is visible to the text-uncond row and (except the padded tail) to the uncond row; this
    # round's masks follow behind it
    assert torch.equal(un2[:, L1:].long(), enc2["uncond_attention_mask"])
    assert torch.equal(tun2[:, L1:].long(), enc2["text_uncond_attention_mask"])
    assert int(tun2[0, L1:].sum()) >= n_patch + 2                  # the edit's input image stays visible without its text
    agent.reset_inner_state()
    assert m.past_key_values is None


@pytest.mark.parametrize("mode", ["DROP", "KEEP"])
def test_multi_round_state_matches_the_reference_trace(stubbed, mode, monkeypatch):
    """SURVEY.md §8a row a23 against the LIVE reference: tests/golden/generate_trace.json holds what the reference's own
    `MingUniVisionForConditionalGeneration.generate` + `prepare_inputs_for_generation` + `forward` + `generate_image` did
    over three rounds of a scripted token stream (text, `<image>`, text, eos; an image right after the prompt; a round
    cut by max_new_tokens), under both PAST_MODEs — driven by the restated HF 4.52.4 greedy loop
    (oracle/hf_generate_oracle.py, the one part that is not the reference's: parity unpinned for it).  The same script
    through this package's `generate` (compute stubbed) must give the same sequences, cache lengths, saved masks, the
    same prefill extents, and hand generate_image the same cache length and masks."""
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generate_trace.json")) as f:
        gold = json.load(f)
    m, llm, log, script, n_tok = stubbed
    assert n_tok == gold["n_image_tokens"] and llm.config.image_start_token == gold["image_start_token"]
    monkeypatch.setenv("PAST_MODE", mode)
    t = lambda v: torch.tensor([v], dtype=torch.int32)  # noqa: E731
    for i, r in enumerate(gold[mode]):
        script["tokens"] = list(r["script"])
        n_pre, n_img = len(log["prefill"]), len(log["image"])
        seq = m.generate(torch.tensor([r["prompt"]]), uncond_attention_mask=t(r["uncond"]),
                         text_uncond_attention_mask=t(r["text_uncond"]), max_new_tokens=r["max_new_tokens"],
                         eos_token_id=gold["eos"])
        assert seq[0].tolist() == r["sequence"], (mode, i)
        assert m.past_key_values.seq_len == r["cache_len"], (mode, i)
        assert m.past_attention_mask[0].tolist() == r["past_attention_mask"]
        assert m.past_uncond_attention_mask[0].tolist() == r["past_uncond_attention_mask"]
        assert m.past_text_uncond_attention_mask[0].tolist() == r["past_text_uncond_attention_mask"]
        # the round's prompt: ONE forward call of the reference behind the cached context, consecutive positions
        first = r["forward_calls"][0]
        assert log["prefill"][n_pre:] == [(first["cache_len"], first["fed"])]
        assert first["position_ids"] == list(range(first["cache_len"], first["cache_len"] + first["fed"]))
        # every later call of the reference feeds ONE token at position == cache length (what greedy_decode does),
        # except the `<image>` start token, which runs generate_image (+ n_tok + 1 cache positions)
        for c in r["forward_calls"][1:]:
            assert c["fed"] == 1 and c["position_ids"] == [c["cache_len"]] and c["attention_mask_len"] == c["cache_len"] + 1
        images = r["generate_image_calls"]
        assert len(log["image"]) - n_img == len(images)
        for (pos, un, tun), g in zip(log["image"][n_img:], images):
            assert pos == g["cache_len"] and g["attention_mask"] == [1] * (g["cache_len"] + 1)
            assert un[0].tolist() == g["uncond"] and tun[0].tolist() == g["text_uncond"]
        assert r["cache_rows"] == 1  # the CFG rows are trimmed again after every image (:1954-1962)
    m.reset_inner_state()


def test_kv_cache_row_bookkeeping_and_growth():
    """BailingKVCache (the static stand-in for the reference's DynamicCache): CFG-row replication / trim
    (modeling_bailing_moe.py:1891-1902, :1954-1962), their batched forms for G requests generated together, and growth by
    re-allocation (the reference's cache grows with every torch.cat) — pure tensor bookkeeping, checked on CPU tensors."""
    from ming_univision_b200.modeling_bailing_moe import BailingKVCache

    cfg = BailingMoeConfig(**synthetic.LLM_TINY_CONFIG)
    L, Hkv, hd = cfg.num_hidden_layers, cfg.num_key_value_heads, cfg.head_dim
    c = BailingKVCache(cfg, max_batch=6, max_len=8, device="cpu")
    assert len(c.k) == L and tuple(c.k[0].shape) == (6, Hkv, 8, hd) and c.get_seq_length() == 0
    g = torch.Generator().manual_seed(0)
    T = 5
    ref_k = [torch.randn((2, Hkv, T, hd), generator=g).to(torch.bfloat16) for _ in range(L)]
    ref_v = [torch.randn((2, Hkv, T, hd), generator=g).to(torch.bfloat16) for _ in range(L)]
    for li in range(L):  # two requests prefilled together: rows 0 and 1
        c.k[li][:2, :, :T] = ref_k[li]
        c.v[li][:2, :, :T] = ref_v[li]
    c.seq_len, c.batch = T, 2
    # G = 2 requests x B = 3 CFG rows: rows g*3 + b start as copies of request g's cond row
    c.expand_groups(2, 3)
    assert c.batch == 6
    for li in range(L):
        for gi in range(2):
            for b in range(3):
                assert torch.equal(c.k[li][gi * 3 + b, :, :T], ref_k[li][gi]) and torch.equal(c.v[li][gi * 3 + b, :, :T], ref_v[li][gi])
    # the generation appends to every row; trimming keeps each request's cond row (row g*3 -> row g)
    for li in range(L):
        c.k[li][:, :, T] = torch.arange(6, dtype=torch.bfloat16).view(6, 1, 1)
    c.seq_len = T + 1
    c.trim_groups(2, 3)
    assert c.batch == 2
    for li in range(L):
        assert torch.equal(c.k[li][0, :, :T], ref_k[li][0]) and torch.equal(c.k[li][1, :, :T], ref_k[li][1])
        assert float(c.k[li][0, 0, T, 0]) == 0.0 and float(c.k[li][1, 0, T, 0]) == 3.0   # rows 0 and 3 of the six
    with pytest.raises(ValueError):
        c.expand_groups(3, 3)
    # single request: repeat_rows / trim_rows
    c.batch = 1
    c.repeat_rows(3)
    assert c.batch == 3 and all(torch.equal(c.k[li][2, :, :T + 1], c.k[li][0, :, :T + 1]) for li in range(L))
    c.trim_rows()
    assert c.batch == 1
    with pytest.raises(ValueError):
        c.repeat_rows(7)
    # growth keeps the content, at least doubles, and changes max_len (the graph workspaces are keyed on it)
    before = [k[:, :, :c.seq_len].clone() for k in c.k]
    c.grow(9)
    assert c.max_len == 16 and all(tuple(k.shape) == (6, Hkv, 16, hd) for k in c.k + c.v)
    assert all(torch.equal(k[:, :, :c.seq_len], b) for k, b in zip(c.k, before))
    c.grow(100)
    assert c.max_len == 100 and all(torch.equal(k[:, :, :c.seq_len], b) for k, b in zip(c.k, before))
