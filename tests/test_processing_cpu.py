"""`ming_univision_b200.processing_bailingmm.BailingMMProcessor` (chat template, image fetching, placeholder expansion,
tokenisation, CFG masks) — host-side string / integer work whose parity bar is EXACT.

Pinned against the LIVE reference where it exists (the build container): the reference's `BailingMMProcessor` methods are
called unbound on an uninitialised instance that only carries the tokenizer (its `ProcessorMixin.__init__` and its
`BailingTokenizer` target transformers 4.52 and do not construct under the transformers 5 of this image), with
  * a synthetic byte-level tokenizer built here with the `tokenizers` library (every string tokenises, special tokens of
    the chat / image markup are single ids), and
  * the reference's real vocabulary (`mingunivision/tokenizer.json`, data) when the checkout is present.
Hand-computed expectations cover the same rules where the reference is absent."""
import base64
import importlib.util
import io
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

from ming_univision_b200 import processing_bailingmm as P
from oracle import ref_shims

REF_DIR = os.path.join(ref_shims.REFERENCE_ROOT, "mingunivision")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF_DIR, "processing_bailingmm.py")),
                               reason="live reference checkout not present")
SPECIALS = ["<role>", "</role>", "<image>", "<imagePatch>", "</image>", "<|endoftext|>", "<|startoftext|>"]


def byte_tokenizer():
    """Byte-level BPE without merges: one id per byte, plus the markup tokens as single ids."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast

    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    tok = Tokenizer(models.BPE(vocab={c: i for i, c in enumerate(alphabet)}, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
    tok.decoder = decoders.ByteLevel()
    tok.add_special_tokens(SPECIALS)
    return PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<|endoftext|>", pad_token="<|endoftext|>",
                                   bos_token="<|startoftext|>", clean_up_tokenization_spaces=False)


@pytest.fixture(scope="module")
def tok():
    return byte_tokenizer()


@pytest.fixture(scope="module")
def ref_module():
    ref_shims.install()
    added = REF_DIR not in sys.path
    if added:
        sys.path.insert(0, REF_DIR)  # (the reference module imports its sibling `bailingmm_utils` flat)
    try:
        spec = importlib.util.spec_from_file_location("_reference_processing_bailingmm",
                                                      os.path.join(REF_DIR, "processing_bailingmm.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if added:
            sys.path.remove(REF_DIR)
    return mod


def _ref_stub(ref_module, tokenizer):
    stub = object.__new__(ref_module.BailingMMProcessor)  # no ProcessorMixin.__init__ (transformers 4.52 API)
    stub.tokenizer = tokenizer
    return stub


def _cpu_processor(tokenizer, **kw):
    """Our processor with CPU stand-ins for the two device transforms (plain torchvision stacks)."""
    import torchvision.transforms as T
    from torchvision.transforms import InterpolationMode

    half = [0.5, 0.5, 0.5]
    und = T.Compose([T.Resize((64, 64), interpolation=InterpolationMode.BICUBIC), T.ToTensor(), T.Normalize(half, half)])
    gen = T.Compose([T.Resize(32, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(32), T.ToTensor(),
                     T.Normalize(half, half)])
    return P.BailingMMProcessor(tokenizer=tokenizer, vis_processor=und, gen_processor=gen, **kw)


def _photo(h, w, seed=0):
    rng = np.random.default_rng(seed)
    return Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))


def conversations():
    img, img2 = _photo(40, 56), _photo(48, 32, 1)
    return [
        [{"role": "HUMAN", "content": [{"type": "text", "text": "Generate a corgi on a beach."}]}],
        [{"role": "HUMAN", "content": [{"type": "image", "image": img}, {"type": "text", "text": "Describe 这张图片 in detail"}]}],
        [{"role": "HUMAN", "content": [{"type": "image", "image": img}, {"type": "text", "text": "make it <image> night"}]}],
        [{"role": "HUMAN", "content": [{"type": "image", "image": [img, img2]}, {"type": "text", "text": "compare"}]},
         {"role": "ASSISTANT", "content": [{"type": "text", "text": "The first is wider."}]},
         {"role": "HUMAN", "content": [{"type": "text", "text": "Now edit the second: add a hat"}]}],
        [{"role": "HUMAN", "content": [{"type": "text", "text": "a"}]},
         {"role": "ASSISTANT", "content": [{"type": "text", "text": "b"}]}],
    ]


def test_chat_template_rules(tok):
    p = _cpu_processor(tok)
    c = conversations()
    assert p.apply_chat_template(c[0]) == "<role>HUMAN</role>Generate a corgi on a beach.<role>ASSISTANT</role>"
    assert p.apply_chat_template(c[0], add_generation_prompt=False, tokenize=False, use_system=True) == \
        "<role>HUMAN</role>Generate a corgi on a beach."
    assert p.apply_chat_template(c[1]).startswith("<role>HUMAN</role><IMAGE>Describe")
    # an `<image>` already written in the message's text stands for one of its images: no placeholder is added
    assert "<IMAGE>" not in p.apply_chat_template(c[2])
    assert p.apply_chat_template(c[3]) == ("<role>HUMAN</role><IMAGE>\n<IMAGE>compare<role>ASSISTANT</role>The first is "
                                           "wider.<|endoftext|><role>HUMAN</role>Now edit the second: add a hat"
                                           "<role>ASSISTANT</role>")
    assert p.apply_chat_template(c[0], system_template="SYS") .startswith("SYSGenerate")
    with pytest.raises(AssertionError):
        p.apply_chat_template([{"role": "SYSTEM", "content": []}])
    with pytest.raises(NotImplementedError):
        p.apply_chat_template([{"role": "HUMAN", "content": [{"type": "video", "video": "x.mp4"}]}])


def test_cfg_masks_by_hand():
    U, A, IMG = [1, 2, 3], [1, 9, 3], {50, 51, 52}
    #       0  1  2 | 3   4   5   6   7  8 | 9  10 11
    seq = [1, 2, 3, 50, 51, 51, 52, 20, 21, 1, 9, 3]
    un, tun = P.cfg_masks(seq, U, A, IMG)
    assert un == [1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1]          # the whole last HUMAN body is hidden
    assert tun == [1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1]         # only its text is hidden, the image tokens stay
    # an OPEN last turn (no ASSISTANT tag behind it): uncond stays all ones, text_uncond runs to the end
    un, tun = P.cfg_masks(seq[:9], U, A, IMG)
    assert un == [1] * 9 and tun == [1, 1, 1, 1, 1, 1, 1, 0, 0]
    # only the LAST HUMAN turn counts; earlier turns stay visible
    two = seq + [30, 31] + [1, 2, 3, 40, 1, 9, 3]
    un, tun = P.cfg_masks(two, U, A, IMG)
    assert un[:17] == [1] * 17 and un[17:] == [0, 1, 1, 1] and tun == un
    assert P.cfg_masks([5, 6, 7], U, A, IMG) == ([1, 1, 1], [1, 1, 1])
    assert P.cfg_masks([], U, A, IMG) == ([], [])


def test_fetch_size_and_fetch_image(tmp_path):
    assert P.fetch_size(40, 56) == (56, 84)                      # grown to the 4 * 28 * 28 minimum
    assert P.fetch_size(1000, 1000) == (896, 896)                # shrunk under the 1024 * 28 * 28 budget
    assert P.fetch_size(300, 500) == (308, 504)
    assert P.fetch_size(10, 300) == (28, 308)
    with pytest.raises(ValueError):
        P.fetch_size(10, 3000)
    img = _photo(300, 500)
    path = str(tmp_path / "a.png")
    img.save(path)
    buf = io.BytesIO()
    img.save(buf, format="PNG")
    uri = "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode()
    outs = [P.fetch_image({"image": src}) for src in (img, path, "file://" + path, uri)]
    assert all(o.size == (504, 308) and o.mode == "RGB" for o in outs)
    assert all(np.array_equal(np.asarray(o), np.asarray(outs[0])) for o in outs)
    assert P.fetch_image({"image": img, "resized_height": 100, "resized_width": 200}).size == (196, 112)
    imgs, vids, auds = P.process_vision_info(conversations()[3])
    assert len(imgs) == 2 and vids is None and auds is None
    assert P.process_vision_info(conversations()[0]) == (None, None, None)
    with pytest.raises(NotImplementedError):
        P.process_vision_info([{"role": "HUMAN", "content": [{"type": "video", "video": "x.mp4"}]}])


def test_call_builds_ids_masks_and_pixels(tok):
    p = _cpu_processor(tok)
    conv = conversations()[1]
    text = p.apply_chat_template(conv)
    images, _, _ = p.process_vision_info(conv)
    for for_edit, side in ((False, 64), (True, 32)):
        out = p(text=[text], images=images, return_tensors="pt", image_patch_size=16, for_edit=for_edit)
        n = (side // 16) ** 2
        assert tuple(out["pixel_values"].shape) == (1, 3, side, side)
        assert out["image_grid_thw"].tolist() == [[1, side // 16, side // 16]]
        ids = out["input_ids"][0].tolist()
        patch = tok.convert_tokens_to_ids("<imagePatch>")
        assert ids.count(patch) == n and out["attention_mask"].shape == out["input_ids"].shape
        un, tun = out["uncond_attention_mask"][0].tolist(), out["text_uncond_attention_mask"][0].tolist()
        user = tok.encode(P.USER_PREFIX, add_special_tokens=False)
        body0 = len(user)
        closing = len(ids) - len(tok.encode(P.ASSISTANT_PREFIX, add_special_tokens=False))
        assert un == [1] * body0 + [0] * (closing - body0) + [1] * (len(ids) - closing)
        img_span = n + 2  # <image> + patches + </image>
        assert tun[body0:body0 + img_span] == [1] * img_span and set(tun[body0 + img_span:closing]) == {0}
    out = p(text=text.replace("<IMAGE>", ""))  # a bare string, no images
    assert "pixel_values" not in out and out["input_ids"].shape[0] == 1
    with pytest.raises(ValueError):
        p(text=123)
    with pytest.raises(NotImplementedError):
        p(text="x", videos=[object()])
    assert p.batch_decode(out["input_ids"])[0] == text.replace("<IMAGE>", "")
    assert p.gen_terminator == [tok.convert_tokens_to_ids("<|endoftext|>")]


def _compare_with_reference(ref_module, tokenizer):
    ref = _ref_stub(ref_module, tokenizer)
    R = ref_module.BailingMMProcessor
    ours = _cpu_processor(tokenizer)
    grids = {1: torch.tensor([[1, 2, 2]]), 2: torch.tensor([[1, 2, 2], [1, 1, 3]])}
    for conv in conversations():
        for agp in (True, False):
            a, b = ours.apply_chat_template(conv, add_generation_prompt=agp), R.apply_chat_template(ref, conv, add_generation_prompt=agp)
            assert a == b
        text = ours.apply_chat_template(conv)
        n = text.count("<IMAGE>")
        if n:
            e_ours, e_ref = ours._expand_image_tokens([text], grids[n]), R._expand_image_tokens(ref, [text], grids[n])
            assert e_ours == e_ref
            text = e_ours[0]
        t_ours, t_ref = ours.tokenize([text]), R.tokenize(ref, [text])
        assert set(t_ours) == set(t_ref)
        for k in t_ref:
            assert t_ours[k].dtype == t_ref[k].dtype and torch.equal(t_ours[k], t_ref[k]), k
    # a multi-round context as the wrapper sees it after an edit round: two samples of different content in one call is
    # not supported by either side without padding; the single-sample path is what the facade uses
    text = ours.apply_chat_template(conversations()[3], add_generation_prompt=False)
    t_ours, t_ref = ours.tokenize(text), R.tokenize(ref, text)
    for k in t_ref:
        assert torch.equal(t_ours[k], t_ref[k]), k


@needs_ref
def test_matches_live_reference_with_synthetic_tokenizer(ref_module, tok):
    _compare_with_reference(ref_module, tok)


@needs_ref
def test_matches_live_reference_with_real_vocabulary(ref_module):
    real = P.load_tokenizer(REF_DIR)
    assert len(real) == 126368 and real.convert_tokens_to_ids("<image>") == 126347
    _compare_with_reference(ref_module, real)


@needs_ref
def test_fetch_matches_live_reference(ref_module):
    import bailingmm_utils as U  # imported flat by the reference module above

    for h, w in ((40, 56), (300, 500), (1000, 1000), (10, 300), (28, 28), (2000, 900), (57, 3001)):
        assert P.fetch_size(h, w) == U.smart_resize(h, w, factor=28, min_pixels=U.MIN_PIXELS, max_pixels=U.MAX_PIXELS)
    for seed, (h, w) in enumerate(((300, 500), (40, 56), (900, 1400))):
        img = _photo(h, w, seed)
        a, b = P.fetch_image({"image": img}), U.fetch_image({"image": img})
        assert a.size == b.size and np.array_equal(np.asarray(a), np.asarray(b))
    conv = conversations()[3]
    ours, ref = P.process_vision_info(conv), U.process_vision_info(conv)
    assert ref[1] is None and ref[2] is None and len(ours[0]) == len(ref[0])
    assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(ours[0], ref[0]))


def test_facade_loads_the_in_package_processor(tmp_path, tok):
    """MingUniVisionInfer finds `tokenizer.json` in the checkpoint directory (else in `reference_dir`) and builds this
    package's processor around it — no reference code is imported."""
    from ming_univision_b200.mingunivisioninfer import MingUniVisionInfer

    tok.backend_tokenizer.save(str(tmp_path / "tokenizer.json"))
    (tmp_path / "tokenizer_config.json").write_text(
        '{"eos_token": "<|endoftext|>", "pad_token": "<|endoftext|>", "bos_token": "<|startoftext|>"}')
    t, p = MingUniVisionInfer._load_processor(str(tmp_path), "/nonexistent")
    assert isinstance(p, P.BailingMMProcessor) and p.tokenizer is t
    assert t.encode("<role>HUMAN</role>hi", add_special_tokens=False) == tok.encode("<role>HUMAN</role>hi",
                                                                                     add_special_tokens=False)
    t2, p2 = MingUniVisionInfer._load_processor("/nonexistent-checkpoint", str(tmp_path))
    assert isinstance(p2, P.BailingMMProcessor)
    with pytest.raises(RuntimeError):
        MingUniVisionInfer._load_processor("/nonexistent-a", "/nonexistent-b")


@needs_ref
def test_from_pretrained_on_the_reference_data_files():
    """The processor built from the reference directory's DATA files (tokenizer.json, tokenizer_config.json,
    preprocessor_config.json): vocabulary, terminator, the video image processor's pixel budget and patch geometry, and
    the two image transforms' configuration (processing_bailingmm.py:175-176)."""
    p = P.BailingMMProcessor.from_pretrained(REF_DIR)
    assert len(p.tokenizer) == 126368 and p.gen_terminator == [126081]
    ip = p.image_processor
    assert (ip.min_pixels, ip.max_pixels, ip.patch_size, ip.temporal_patch_size, ip.merge_size) == (78400, 802816, 14, 2, 2)
    assert (p.vis_processor.image_size, p.gen_processor.image_size) == (1024, 512)
    assert p.vis_processor.mean == (0.5, 0.5, 0.5) and p.gen_processor.std == (0.5, 0.5, 0.5)
    # the ids the model's config hard-wires (mingunivision/config.json: image_patch_token / image_start_token) are the
    # tokenizer's ids of the markup tokens
    assert p.tokenizer.convert_tokens_to_ids(["<imagePatch>", "<image>", "</image>"]) == [126346, 126347, 126348]
