"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mingb200.h
declares, the ctypes table covers all of them, and the product path refuses to run without a GPU (no fallback)."""
import ctypes
import os

import pytest
import torch

from ming_univision_b200 import _lib, synthetic


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mingb200.h but not exported by libmingb200.so"
    undeclared = set(_lib.SIGNATURES) - set(names)
    assert not undeclared, f"ctypes table binds symbols missing from the header: {undeclared}"
    unbound = set(names) - set(_lib.SIGNATURES) - {"mb_last_error"}
    assert not unbound, f"header symbols without a ctypes signature: {unbound}"
    assert lib.mb_abi_version() >= 1
    assert isinstance(lib.mb_last_error(), bytes)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.mb_device_ok() == 0
    rc = lib.mb_gemm_bf16(None, 8, None, 8, None, None, 8, 1, 8, 8, 0, None, 0, 0, 0, 0, None)
    assert rc == -3  # MB_ERR_ARCH
    assert b"sm_100" in lib.mb_last_error()
    with pytest.raises(RuntimeError):
        _lib.require_device()
    from ming_univision_b200.mingtok import MingTok, MingTokConfig

    m = MingTok(MingTokConfig(**synthetic.MINGTOK_TINY_CONFIG))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward(torch.zeros(1, 3, 128, 128))


def test_sass_has_blackwell_paths():
    """The shipped cubin takes the hardware paths DESIGN.md claims: the GEMM and the flash attention issue tcgen05.mma
    (UTCHMMA) with TMEM loads (LDTM), TMA tensor loads and mbarriers; no kernel of the GEMM family falls back to mma.sync."""
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    bodies = {}
    for part in sass.split("Function : ")[1:]:
        name, body = part.split("\n", 1)
        bodies[name.strip()] = body
    gemm = [b for n, b in bodies.items() if "gemm_bf16_kernel" in n]
    attn = [b for n, b in bodies.items() if "attn_tc_kernel" in n]
    assert len(gemm) >= 8 and len(attn) == 2
    for b in gemm + attn:
        assert "UTCHMMA" in b and "LDTM" in b and "UTMALDG" in b and "SYNCS" in b
        assert " HMMA." not in b
    assert all("UTMASTG" in b for b in gemm)  # staged TMA-store epilogue
    assert any("UTMALDG.2D.2CTA" in b for b in gemm)  # the cta_group::2 instances


def test_state_dict_schema_matches_reference_keys():
    """Same keys and shapes as the reference's MingTok.state_dict() and HF save/load round trip."""
    import tempfile

    from ming_univision_b200.mingtok import MingTok, MingTokConfig

    cfg = synthetic.MINGTOK_TINY_CONFIG
    m = MingTok(MingTokConfig(**cfg))
    shapes = synthetic.mingtok_param_shapes(cfg)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    m.load_state_dict(synthetic.mingtok_state_dict(cfg, 3), strict=True)
    with tempfile.TemporaryDirectory() as d:
        m.save_pretrained(d)
        m2 = MingTok.from_pretrained(d)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert (m.latent_dim, m.feature_dim, m.patch_size) == (32, 128, 32)


def test_reference_keys_live():
    """When the reference tree is present (build container), compare the key schema against the real thing."""
    from oracle import ref_shims

    if not ref_shims.reference_available():
        pytest.skip("reference tree not present")
    cfg = synthetic.MINGTOK_TINY_CONFIG
    ref = ref_shims.build_reference_mingtok(cfg)
    shapes = synthetic.mingtok_param_shapes(cfg)
    ref_sd = ref.state_dict()
    assert set(ref_sd) == set(shapes)
    assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)


def test_rf_state_dict_schema_live():
    """RectifiedFlowLoss keys/shapes == the reference module's (and == the synthetic factory's)."""
    from ming_univision_b200.diff_loss_rf_swiglu import RectifiedFlowLoss
    from oracle import ref_shims

    cfg = synthetic.RF_TINY_CONFIG
    m = RectifiedFlowLoss(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                          str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    shapes = synthetic.rf_param_shapes(cfg)
    sd = m.state_dict()
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(torch.zeros(2, cfg["z_channels"]), text_cfg=3.0)
    if not ref_shims.reference_available():
        return
    ref_shims.install()
    import contextlib
    import io

    os.environ["XFORMERS_DISABLED"] = "1"
    from diff_loss_rf_swiglu import RectifiedFlowLoss as RefRF

    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefRF(cfg["target_channels"], cfg["z_channels"], cfg["depth"], cfg["width"],
                    str(cfg["num_sampling_steps"]), mlp_mult=cfg["mlp_mult"])
    rsd = ref.state_dict()
    assert set(rsd) == set(shapes) and all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)


def test_llm_state_dict_schema_live():
    """BailingMoeForCausalLM (+ vis_head + diffloss) keys / shapes == the reference module's == the synthetic factory's."""
    from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig, BailingMoeForCausalLM
    from oracle import ref_shims

    cfg, vh = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG
    m = BailingMoeForCausalLM(BailingMoeConfig(**cfg))
    m.setup_vishead_diffloss(**vh)
    shapes = synthetic.llm_param_shapes(cfg, vh)
    sd = m.state_dict()
    assert set(sd) == set(shapes), (sorted(set(sd) ^ set(shapes))[:6])
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    if not ref_shims.reference_available():
        return
    ref, _ = ref_shims.build_reference_llm(cfg, vh)
    rsd = ref.state_dict()
    # the reference also stores the rotary inv_freq buffers; ours are derived from rope_theta and dropped on load
    extra = {k for k in rsd if k.endswith("rotary_emb.inv_freq")}
    assert set(rsd) - extra == set(shapes), (sorted((set(rsd) - extra) ^ set(shapes))[:6])
    assert all(tuple(rsd[k].shape) == tuple(shapes[k]) for k in shapes)
    res = m.load_state_dict(rsd, strict=True)  # a reference checkpoint loads unchanged
    assert not res.missing_keys and not res.unexpected_keys


def test_load_checkpoint_hf_layout(tmp_path, monkeypatch):
    """mingunivisioninfer.load_checkpoint: config.json + *.safetensors in the reference's HF layout (LLM + vis_head +
    diffloss + linear_proj shards at the top level, MingTok under models/MingTok-Vision) -> the wrapper module, with the
    reference-only buffers (rotary inv_freq) ignored.  Synthetic tiny checkpoint (no pretrained weights exist offline)."""
    import json

    from safetensors.torch import save_file

    from ming_univision_b200.mingunivisioninfer import load_checkpoint

    cfg, vh, tok = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG, synthetic.MINGTOK_TINY_CONFIG
    llm_sd = synthetic.llm_state_dict(cfg, vh, tok["semantic_decoder"]["embed_dim"], 0)
    top = {(k if k.startswith("linear_proj.") else "model." + k): v.contiguous() for k, v in llm_sd.items()}
    top["model.model.layers.0.attention.rotary_emb.inv_freq"] = torch.ones(4)
    keys = sorted(top)
    save_file({k: top[k] for k in keys[: len(keys) // 2]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: top[k] for k in keys[len(keys) // 2:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    (tmp_path / "models" / "MingTok-Vision").mkdir(parents=True)
    save_file({k: v.contiguous() for k, v in synthetic.mingtok_state_dict(tok, 0).items()},
              str(tmp_path / "models" / "MingTok-Vision" / "model.safetensors"))
    with open(tmp_path / "models" / "MingTok-Vision" / "config.json", "w") as f:
        json.dump(tok, f)
    with open(tmp_path / "config.json", "w") as f:
        json.dump({"llm_config": dict(cfg, architectures=["BailingMoeForCausalLM"], torch_dtype="bfloat16"),
                   "vishead_diffloss_config": vh}, f)
    m = load_checkpoint(str(tmp_path), device="cpu")
    sd = m.state_dict()
    assert sd["model.lm_head.weight"].dtype == torch.bfloat16
    # the call the reference's facade makes (mingunivisioninfer.py:72-78) lands in the same loader
    from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration as Wrapper

    m2 = Wrapper.from_pretrained(str(tmp_path), torch_dtype=torch.bfloat16, attn_implementation="flash_attention_2",
                                 trust_remote_code=True, device_map="cpu")
    assert all(torch.equal(v, m2.state_dict()[k]) for k, v in sd.items())
    with pytest.raises(NotImplementedError):
        Wrapper.from_pretrained(str(tmp_path), quantization_config=object(), device_map="cpu")
    # a hub id instead of a directory goes through huggingface_hub's snapshot (stubbed: there is no network here)
    import huggingface_hub

    asked = []
    monkeypatch.setattr(huggingface_hub, "snapshot_download", lambda name, **kw: asked.append(name) or str(tmp_path))
    m3 = Wrapper.from_pretrained("inclusionAI/Ming-UniVision-16B-A3B", device_map="cpu")
    assert asked == ["inclusionAI/Ming-UniVision-16B-A3B"] and torch.equal(m3.state_dict()["model.lm_head.weight"],
                                                                           sd["model.lm_head.weight"])
    for k, v in top.items():
        if k.endswith("inv_freq"):
            continue
        assert torch.equal(sd[k].float(), v.to(torch.bfloat16).float()), k
    assert torch.equal(sd["vision.sem_to_pix.weight"].float(),
                       synthetic.mingtok_state_dict(tok, 0)["sem_to_pix.weight"].to(torch.bfloat16).float())


def test_load_checkpoint_expert_parallel_shard(tmp_path, monkeypatch):
    """load_checkpoint(ep_rank, ep_size): a rank allocates and READS only its own routed experts (direct
    safetensors -> EP-sharded slabs); everything else is replicated; the staging chunks are flushed as they fill."""
    import json

    from safetensors.torch import save_file

    from ming_univision_b200 import mingunivisioninfer as MI

    cfg, vh, tok = synthetic.LLM_TINY_CONFIG, synthetic.VISHEAD_TINY_CONFIG, synthetic.MINGTOK_TINY_CONFIG
    llm_sd = synthetic.llm_state_dict(cfg, vh, tok["semantic_decoder"]["embed_dim"], 0)
    top = {(k if k.startswith("linear_proj.") else "model." + k): v.contiguous() for k, v in llm_sd.items()}
    save_file(top, str(tmp_path / "model.safetensors"))
    (tmp_path / "models" / "MingTok-Vision").mkdir(parents=True)
    save_file({k: v.contiguous() for k, v in synthetic.mingtok_state_dict(tok, 0).items()},
              str(tmp_path / "models" / "MingTok-Vision" / "model.safetensors"))
    with open(tmp_path / "config.json", "w") as f:
        json.dump({"llm_config": cfg, "vishead_diffloss_config": vh, "mingtok_config": tok}, f)
    monkeypatch.setattr(MI, "_LOAD_CHUNK_BYTES", 64 << 10)  # many flushes on the tiny checkpoint
    E = cfg["num_experts"]
    world = 2
    seen = set()
    for rank in range(world):
        m = MI.load_checkpoint(str(tmp_path), device="cpu", ep_rank=rank, ep_size=world)
        sd = m.state_dict()
        local = range(rank * E // world, (rank + 1) * E // world)
        for k, v in top.items():
            if ".mlp.experts." in k:
                e = int(k.split(".mlp.experts.")[1].split(".")[0])
                if e not in local:
                    assert sd[k].is_meta, k            # other ranks' experts: never allocated, never read
                    continue
                seen.add(k)
            assert not sd[k].is_meta and torch.equal(sd[k].float(), v.to(torch.bfloat16).float()), k
    assert seen == {k for k in top if ".mlp.experts." in k}  # the ranks together hold every routed expert exactly once
    with pytest.raises(ValueError):
        MI.load_checkpoint(str(tmp_path), device="cpu", ep_rank=2, ep_size=2)


def test_full_size_tree_from_the_reference_config(tmp_path):
    """`build_from_config` on the reference's REAL `mingunivision/config.json` + `mingtok/config/config_mingtok.json`
    (data files of the checkout, when present; otherwise this package's own full-size configs), on the META device: the
    complete 16B-A3B parameter tree without a byte of storage.  Parameter counts are SURVEY.md §8d's constants; with
    expert parallelism a rank's tree has storage for its 64 / 8 routed experts only."""
    import json
    import shutil

    from ming_univision_b200.mingunivisioninfer import build_from_config
    from oracle import ref_shims

    ref_cfg = os.path.join(ref_shims.REFERENCE_ROOT, "mingunivision", "config.json")
    ref_tok = os.path.join(ref_shims.REFERENCE_ROOT, "mingtok", "config", "config_mingtok.json")
    (tmp_path / "models" / "MingTok-Vision").mkdir(parents=True)
    if os.path.isfile(ref_cfg) and os.path.isfile(ref_tok):
        shutil.copy(ref_cfg, tmp_path / "config.json")           # (it carries no vishead_diffloss_config: defaults apply)
        shutil.copy(ref_tok, tmp_path / "models" / "MingTok-Vision" / "config.json")
    else:
        with open(tmp_path / "config.json", "w") as f:
            json.dump({"llm_config": synthetic.LLM_CONFIG, "vishead_diffloss_config": synthetic.VISHEAD_CONFIG}, f)
        with open(tmp_path / "models" / "MingTok-Vision" / "config.json", "w") as f:
            json.dump(synthetic.MINGTOK_CONFIG, f)
    with pytest.warns(UserWarning) if os.path.isfile(ref_cfg) else _nullcontext():
        m = build_from_config(str(tmp_path), device="meta")
    n = lambda mod: sum(p.numel() for p in mod.parameters())  # noqa: E731
    assert n(m.model.model) + n(m.model.lm_head) == 16_809_314_304        # 16.81 B (2.50 B active per token)
    assert n(m.model.diffloss) == 1_284_876_320 and n(m.model.vis_head) == 6_300_672
    assert n(m.vision) == 697_719_584 and n(m.linear_proj) == 6_295_552
    assert all(p.is_meta and p.dtype == torch.bfloat16 for p in m.parameters())
    cfg = m.model.config
    assert (cfg.num_experts, cfg.num_experts_per_tok, cfg.num_hidden_layers, cfg.vocab_size) == (64, 6, 28, 126464)
    rs = cfg.rope_scaling  # the checkpoint's "3D" entry is mapped to the 1-D legacy rotary (transformers 5 spells None "default")
    assert rs is None or rs.get("rope_type", rs.get("type")) == "default"
    assert cfg.image_patch_token == 126346 and cfg.image_start_token == 126347
    m8 = build_from_config(str(tmp_path), device="meta", ep_rank=3, ep_size=8)
    keys = [k for k in m8.state_dict() if ".mlp.experts." in k]
    assert len(keys) == 28 * 64 * 3                                       # the key schema is unchanged under sharding
    assert m8.model.model.layers[0].mlp.experts[24].gate_proj.weight.shape == (1408, 2048)


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
