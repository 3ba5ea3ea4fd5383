"""CPU checks of the image pre- / post-processing path:

1. the numpy oracle (oracle/preprocess_oracle.py) is pinned against Pillow and torchvision THEMSELVES (live, bit-exact)
   and against the committed fixtures they produced (tests/golden/preprocess.npz);
2. the device code — the per-thread bodies of ming_univision_b200/csrc/preprocess_core.h, shared verbatim by the CUDA
   kernels — is compiled with g++ and run over the kernels' grids on the CPU (tests/native/preprocess_emu.cpp): every
   index computation, coefficient and rounding step must reproduce the oracle exactly;
3. the host logic (torchvision's size / crop rules, workspace sizing through the real C ABI, error behaviour).
The launches themselves are checked on the GPU box (tests/test_preprocess_gpu.py).
"""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HALF = (0.5, 0.5, 0.5)
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def synthetic_photo(rng, h, w):
    from PIL import Image

    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = np.asarray(Image.fromarray(base).resize((w, h), Image.BILINEAR)).astype(np.int32)
    return np.clip(img + rng.integers(-24, 25, (h, w, 3)), 0, 255).astype(np.uint8)


def torchvision_pipeline(img, size, crop, mean, std):
    """The reference's transform stack, verbatim composition (mingtok/utils/processor.py:17-27)."""
    import torchvision.transforms as T
    from PIL import Image
    from torchvision.transforms import InterpolationMode

    tf = [T.Resize(size=size, interpolation=InterpolationMode.BICUBIC)]
    if crop is not None:
        tf.append(T.CenterCrop(crop))
    tf += [T.ToTensor(), T.Normalize(mean, std)]
    return T.Compose(tf)(Image.fromarray(img)).numpy()


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(ROOT, "tests", "golden", "preprocess.npz"))
    return g, json.loads(str(g["cases"]))


# ---------------------------------------------------------------------------------------------------------------------
# 1. oracle pinned
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,oh,ow", [(64, 64, 32, 32), (100, 75, 37, 51), (33, 47, 80, 90), (300, 200, 512, 341),
                                       (517, 389, 256, 256), (64, 64, 64, 32), (64, 64, 128, 64), (7, 9, 3, 4),
                                       (1, 1, 5, 5), (5, 5, 1, 1), (1000, 60, 17, 60), (640, 480, 20, 15),
                                       # more than 100 times taller than wide: Pillow >= 11 goes height-first when the
                                       # height shrinks (PIL/Image.py, Image.resize) — and only then
                                       (306, 3, 95, 203), (221, 2, 153, 278), (301, 3, 300, 7), (300, 3, 95, 203),
                                       (306, 3, 400, 203), (3, 306, 203, 95)])
def test_oracle_resize_matches_pillow(h, w, oh, ow):
    from PIL import Image

    img = synthetic_photo(np.random.default_rng(h * 1000 + w), h, w)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(po.resize_bicubic_u8(img, oh, ow), ref)


def test_oracle_resize_extreme_pixels():
    """Saturated checkerboards drive the cubic's overshoot into both clips (and the int32 accumulator to its widest)."""
    from PIL import Image

    yy, xx = np.mgrid[0:96, 0:128]
    img = np.repeat((((yy // 3 + xx // 2) % 2) * 255).astype(np.uint8)[:, :, None], 3, axis=2)
    img[:, :, 1] = 255 - img[:, :, 1]
    for oh, ow in ((40, 50), (200, 300), (96, 31)):
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(po.resize_bicubic_u8(img, oh, ow), ref)


@pytest.mark.parametrize("h,w,size,crop,mean,std", [
    (517, 389, 256, 256, HALF, HALF), (300, 451, 512, 512, HALF, HALF), (350, 350, (512, 512), None, HALF, HALF),
    (256, 256, 256, 256, HALF, HALF), (320, 240, (224, 224), None, CLIP_MEAN, CLIP_STD), (99, 64, 33, 33, HALF, HALF),
    (64, 99, 33, 33, CLIP_MEAN, CLIP_STD)])
def test_oracle_pipeline_matches_torchvision(h, w, size, crop, mean, std):
    img = synthetic_photo(np.random.default_rng(h + 7 * w), h, w)
    ref = torchvision_pipeline(img, size, crop, mean, std)
    got = po.preprocess(img, size, crop, mean, std)
    assert got.dtype == np.float32 and np.array_equal(got, ref)


def test_oracle_matches_golden(golden):
    g, cases = golden
    for c in cases:
        src, name = g[c["name"] + ".src"], c["name"]
        rh, rw = po.resized_output_size(src.shape[0], src.shape[1], c["size"])
        u8 = po.resize_bicubic_u8(src, rh, rw)
        if c["crop"] is not None:
            t, l = po.center_crop_offsets(rh, rw, c["crop"], c["crop"])
            u8 = u8[t:t + c["crop"], l:l + c["crop"]]
        assert np.array_equal(u8, g[name + ".u8"]), name
        assert np.array_equal(po.preprocess(src, c["size"], c["crop"], c["mean"], c["std"]), g[name + ".tensor"]), name
    assert np.array_equal(po.postprocess(g["post.x"][0]), g["post.u8"])


def test_golden_is_what_the_libraries_produce_today(golden):
    """The fixtures were made by Pillow / torchvision; if the installed versions ever disagree with them, say so here
    rather than in a parity failure."""
    g, cases = golden
    for c in cases:
        size = tuple(c["size"]) if isinstance(c["size"], list) else c["size"]
        ref = torchvision_pipeline(g[c["name"] + ".src"], size, c["crop"], c["mean"], c["std"])
        assert np.array_equal(ref, g[c["name"] + ".tensor"]), c["name"]


def test_oracle_postprocess_matches_topilimage():
    import torchvision.transforms as T

    rng = np.random.default_rng(5)
    noise = torch.from_numpy(rng.uniform(-1, 1, (3, 20, 32)).astype(np.float32))
    # every u8 level boundary: the exact pre-images k/255 and their fp32 neighbours on both sides
    levels = torch.arange(256, dtype=torch.float32) / 255 * 2 - 1
    grid = torch.stack([levels, torch.nextafter(levels, torch.tensor(2.0)), torch.nextafter(levels, torch.tensor(-2.0))])
    x = torch.cat([noise, grid.clamp(-1, 1).reshape(3, 8, 32)], dim=1).contiguous()
    half = torch.tensor(HALF).view(1, -1, 1, 1)
    ref = np.asarray(T.ToPILImage()((x[None] * half + half)[0]))
    assert np.array_equal(po.postprocess(x.numpy()), ref)


def test_save_image_truncates_like_the_reference(tmp_path):
    """MingUniVisionForConditionalGeneration._save_image == tensor_to_pil + save (modeling_bailing_moe.py:84-90, :1787-1795)."""
    from PIL import Image

    from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration as Wrapper

    x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (3, 32, 48)).astype(np.float32))
    Wrapper._save_image(x, str(tmp_path / "o"), 0)
    Wrapper._save_image(x, str(tmp_path / "o"), 2)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "o.png")), po.postprocess(x.numpy()))
    assert os.path.exists(tmp_path / "o_2.png")


# ---------------------------------------------------------------------------------------------------------------------
# 2. the device code, emulated on the CPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("preemu") / "libpreemu.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-Wall", "-Werror", "-o",
                    str(so), os.path.join(ROOT, "tests", "native", "preprocess_emu.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emu_workspace_bytes.restype = C.c_longlong
    return lib


def emu_preprocess(lib, imgs, size, crop, mean=HALF, std=HALF):
    imgs = np.ascontiguousarray(imgs)
    n, h, w, _ = imgs.shape
    rh, rw = po.resized_output_size(h, w, size)
    if crop is None:
        top, left, oh, ow = 0, 0, rh, rw
    else:
        (top, left), oh, ow = po.center_crop_offsets(rh, rw, crop, crop), crop, crop
    out = np.full((n, 3, oh, ow), np.nan, dtype=np.float32)
    plan = (C.c_int * 8)()
    rc = lib.emu_image_preprocess(imgs.ctypes.data_as(C.c_void_p), n, h, w, rh, rw, top, left, oh, ow,
                                  (C.c_float * 3)(*mean), (C.c_float * 3)(*std), out.ctypes.data_as(C.c_void_p), plan)
    assert rc == 0, rc
    return out, dict(zip(("do_h", "do_v", "ksize_h", "ksize_v", "row0", "rows", "tile_w", "tile_rows"), plan))


@pytest.mark.parametrize("in_size,out_size", [(64, 32), (389, 256), (47, 90), (1000, 17), (512, 512), (3000, 256),
                                              (5, 1), (1, 5), (4096, 1023)])
def test_device_coefficients_equal_pillows(emu, in_size, out_size):
    bounds, kk, ksize = po.precompute_coeffs(in_size, out_size)
    b = np.zeros((out_size, 2), dtype=np.int32)
    k = np.full((out_size, ksize), -7, dtype=np.int32)
    assert emu.emu_axis_coeffs(in_size, out_size, b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p)) == ksize
    assert np.array_equal(b, bounds) and np.array_equal(k, kk)
    assert (np.abs(kk.astype(np.int64)).sum(axis=1) * 255 < 2 ** 31 - 2 ** 21).all()  # the int32 accumulator cannot overflow


@pytest.mark.parametrize("h,w,size,crop,mean,std", [
    (517, 389, 256, 256, HALF, HALF),               # portrait: rows cropped (row0 > 0)
    (300, 451, 512, 512, HALF, HALF),               # enlarging
    (350, 350, (512, 512), None, HALF, HALF),       # MingTokUndProcessor shape, no crop
    (256, 256, 256, 256, HALF, HALF),               # nothing to resample
    (320, 240, (224, 224), None, CLIP_MEAN, CLIP_STD),
    (97, 131, 64, 64, HALF, HALF), (131, 97, 64, 64, HALF, HALF),
    (600, 800, 256, 256, HALF, HALF), (1000, 1500, 128, 128, HALF, HALF),
    (64, 640, 64, 64, HALF, HALF), (640, 64, 64, 64, HALF, HALF),   # crop only
    (33, 33, (7, 5), None, HALF, HALF),
    (64, 100, (64, 50), None, HALF, HALF),          # horizontal pass only
    (100, 64, (50, 64), None, HALF, HALF),          # vertical pass only
    (40, 2600, (40, 3), None, HALF, HALF),          # ~870-fold reduction: the horizontal tile narrows to fit 48 KB
    (3, 301, (5, 299), None, HALF, HALF),           # odd byte alignments of every staged row segment
])
def test_emulated_device_path_is_bit_exact(emu, h, w, size, crop, mean, std):
    rng = np.random.default_rng(h * 31 + w)
    imgs = np.stack([synthetic_photo(rng, h, w) for _ in range(2)])
    got, plan = emu_preprocess(emu, imgs, size, crop, mean, std)
    assert not np.isnan(got).any(), "an output element was never written"
    for i in range(2):
        assert np.array_equal(got[i], po.preprocess(imgs[i], size, crop, mean, std)), plan
    assert np.array_equal(got[0], torchvision_pipeline(imgs[0], size, crop, mean, std)), plan
    rh, rw = po.resized_output_size(h, w, size)
    assert plan["do_h"] == int(rw != w) and plan["do_v"] == int(rh != h)


def test_height_first_geometry_is_refused(emu):
    """Images more than 100 times taller than wide whose height shrinks: Pillow resizes them height-first (the oracle
    follows, test_oracle_resize_matches_pillow); the device path refuses instead of answering differently."""
    assert po.vertical_pass_first(306, 3, 95) and not po.vertical_pass_first(300, 3, 95)
    assert not po.vertical_pass_first(306, 3, 400) and not po.vertical_pass_first(3, 306, 1)
    assert emu.emu_workspace_bytes(1, 306, 3, 95, 203, 0, 0, 95, 203) == -1
    assert emu.emu_workspace_bytes(1, 300, 3, 95, 203, 0, 0, 95, 203) > 0
    assert emu.emu_workspace_bytes(1, 306, 3, 400, 203, 0, 0, 400, 203) > 0
    from ming_univision_b200 import _lib

    lib, nbytes = _lib.load(), C.c_int64(-1)
    assert lib.mb_image_preprocess_workspace_bytes(1, 306, 3, 95, 203, 0, 0, 95, 203, C.byref(nbytes)) == -1
    assert b"height-first" in lib.mb_last_error() and nbytes.value == -1


def test_emulated_u8_output_is_pillows_resize(emu):
    """out_kind 2: the resized + cropped u8 image itself == PIL's Image.resize (+ crop), e.g. fetch_image's resize
    (bailingmm_utils.py:162) on the device."""
    from PIL import Image

    rng = np.random.default_rng(8)
    for h, w, oh, ow, top, left, ch, cw in [(300, 400, 224, 308, 0, 0, 224, 308), (97, 131, 64, 86, 0, 11, 64, 64),
                                            (50, 60, 120, 150, 7, 9, 100, 100), (64, 64, 64, 32, 0, 0, 64, 32)]:
        img = synthetic_photo(rng, h, w)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))[top:top + ch, left:left + cw]
        f32 = np.empty((1, 3, ch, cw), dtype=np.float32)
        u8 = np.full((1, ch, cw, 3), 9, dtype=np.uint8)
        rc = emu.emu_image_preprocess_ex(np.ascontiguousarray(img[None]).ctypes.data_as(C.c_void_p), 1, h, w, oh, ow, top,
                                         left, ch, cw, (C.c_float * 3)(*HALF), (C.c_float * 3)(*HALF),
                                         f32.ctypes.data_as(C.c_void_p), u8.ctypes.data_as(C.c_void_p), None)
        assert rc == 0 and np.array_equal(u8[0], ref)


def test_emulated_empty_batch(emu):
    out, plan = emu_preprocess(emu, np.zeros((0, 50, 60, 3), dtype=np.uint8), 32, 32)
    assert out.shape == (0, 3, 32, 32) and plan["do_h"] == 1 and plan["do_v"] == 1


def test_emulated_plan_narrows_the_tile(emu):
    imgs = np.zeros((1, 40, 2600, 3), dtype=np.uint8)
    _, plan = emu_preprocess(emu, imgs, (40, 3), None)
    assert plan["tile_w"] * plan["tile_rows"] < 256 and plan["ksize_h"] > 3000


def test_emulated_golden(emu, golden):
    g, cases = golden
    for c in cases:
        size = tuple(c["size"]) if isinstance(c["size"], list) else c["size"]
        got, _ = emu_preprocess(emu, g[c["name"] + ".src"][None], size, c["crop"], c["mean"], c["std"])
        assert np.array_equal(got[0], g[c["name"] + ".tensor"]), c["name"]
    x = np.ascontiguousarray(g["post.x"])
    out = np.zeros((1,) + g["post.u8"].shape, dtype=np.uint8)
    emu.emu_image_postprocess(x.ctypes.data_as(C.c_void_p), 1, x.shape[2], x.shape[3], (C.c_float * 3)(*HALF),
                              (C.c_float * 3)(*HALF), out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out[0], g["post.u8"])


def test_emulated_postprocess_every_level(emu):
    levels = torch.arange(256, dtype=torch.float32) / 255 * 2 - 1
    grid = torch.stack([levels, torch.nextafter(levels, torch.tensor(2.0)), torch.nextafter(levels, torch.tensor(-2.0))])
    x = np.ascontiguousarray(grid.clamp(-1, 1).reshape(1, 3, 16, 16).numpy())
    out = np.zeros((1, 16, 16, 3), dtype=np.uint8)
    emu.emu_image_postprocess(x.ctypes.data_as(C.c_void_p), 1, 16, 16, (C.c_float * 3)(*HALF), (C.c_float * 3)(*HALF),
                              out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out[0], po.postprocess(x[0]))


# ---------------------------------------------------------------------------------------------------------------------
# 3. host logic through the real library
# ---------------------------------------------------------------------------------------------------------------------
def test_size_and_crop_rules_match_torchvision():
    import torchvision.transforms.functional as F

    from ming_univision_b200 import ops

    for h, w, s in [(517, 389, 256), (389, 517, 256), (300, 300, 512), (1, 1000, 7), (1333, 800, 512), (801, 800, 512)]:
        assert list(ops.resized_output_size(h, w, s)) == F._compute_resized_output_size((h, w), [s])
        assert ops.resized_output_size(h, w, s) == po.resized_output_size(h, w, s)
    assert ops.resized_output_size(10, 20, (7, 9)) == (7, 9)
    for full, crop in [(341, 256), (342, 256), (343, 256), (256, 256), (259, 256), (261, 256)]:
        x = torch.arange(full).view(1, full, 1).expand(1, full, full)
        top = int(F.center_crop(x, [crop, crop])[0, 0, 0])
        assert ops.center_crop_window(full, full, crop, crop) == (top, top)  # half-to-even rounding included
    with pytest.raises(ValueError):
        ops.center_crop_window(10, 300, 256, 256)


def test_workspace_bytes_through_the_c_abi(emu):
    from ming_univision_b200 import _lib

    lib = _lib.load()
    for geom in [(2, 517, 389, 340, 256, 42, 0, 256, 256), (1, 256, 256, 256, 256, 0, 0, 256, 256),
                 (64, 300, 451, 512, 769, 0, 128, 512, 512), (1, 40, 2600, 40, 3, 0, 0, 40, 3)]:
        nbytes = C.c_int64(-1)
        assert lib.mb_image_preprocess_workspace_bytes(*geom, C.byref(nbytes)) == 0
        assert nbytes.value == emu.emu_workspace_bytes(*geom) and nbytes.value % 16 == 0
    nbytes = C.c_int64(-1)
    # crop window outside the resized image, empty sizes: MB_ERR_SHAPE with a message, nothing written
    assert lib.mb_image_preprocess_workspace_bytes(1, 100, 100, 64, 64, 10, 0, 64, 64, C.byref(nbytes)) == -1
    assert b"invalid geometry" in lib.mb_last_error() and nbytes.value == -1
    assert lib.mb_image_preprocess_workspace_bytes(1, 0, 100, 64, 64, 0, 0, 64, 64, C.byref(nbytes)) == -1


def test_processors_refuse_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ming_univision_b200 import ops
    from ming_univision_b200.mingtok.utils import CenterCropProcessor

    with pytest.raises(TypeError, match="CUDA uint8"):
        ops.image_preprocess(torch.zeros((8, 8, 3), dtype=torch.uint8), 4, 4)
    with pytest.raises(TypeError, match="CUDA float32"):
        ops.image_postprocess(torch.zeros((3, 8, 8)))
    proc = CenterCropProcessor(image_size=4)
    assert (proc.image_size, proc.mean, proc.std) == (4, HALF, HALF)
    with pytest.raises((RuntimeError, AssertionError)):  # no device to move the pixels to, no CPU fallback
        proc(np.zeros((8, 8, 3), dtype=np.uint8))
    lib = __import__("ming_univision_b200._lib", fromlist=["load"]).load()
    assert lib.mb_image_postprocess_u8(None, 1, 1, 8, 8, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, None, None) == -3  # MB_ERR_ARCH


# ---------------------------------------------------------------------------------------------------------------------
# 4. the reference's own fixture for this path (read in the build container only; absent on the GPU box)
# ---------------------------------------------------------------------------------------------------------------------
REF_ASSET = "/root/reference/mingtok/asset/mingtok.png"


@pytest.mark.skipif(not os.path.exists(REF_ASSET), reason="reference checkout not present")
def test_reference_asset_through_all_three_processors(emu):
    """mingtok/asset/mingtok.png (a real 512 x 512 photograph, the input of test_infer_recon_image.py:14) through the
    demo's CenterCropProcessor(512) (nothing to resample: the crop-only path), the generation processor at 256 and the
    understanding processor's Resize((1024, 1024)): oracle and emulated device code == torchvision, bit for bit."""
    from PIL import Image

    img = np.array(Image.open(REF_ASSET).convert("RGB"))
    assert img.shape == (512, 512, 3)
    for size, crop in ((512, 512), (256, 256), ((1024, 1024), None), (200, 200)):
        ref = torchvision_pipeline(img, size, crop, HALF, HALF)
        assert np.array_equal(po.preprocess(img, size, crop), ref), (size, crop)
        got, _ = emu_preprocess(emu, img[None], size, crop)
        assert np.array_equal(got[0], ref), (size, crop)
    # and the way back (test_infer_recon_image.py:24-28): normalised tensor -> PIL bytes
    import torchvision.transforms as T

    x = torch.from_numpy(torchvision_pipeline(img, 512, 512, HALF, HALF))
    half = torch.tensor(HALF).view(1, -1, 1, 1)
    back = np.asarray(T.ToPILImage()((x[None] * half + half)[0]))
    assert np.array_equal(po.postprocess(x.numpy()), back)
    assert np.abs(back.astype(int) - img.astype(int)).max() <= 1  # truncation may lose one level, never more


def test_device_code_is_memory_safe(tmp_path):
    """The same device code under AddressSanitizer + UBSan with exact-size buffers (tests/native/preprocess_asan.cpp)."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "preprocess_asan"
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            "-ffp-contract=off", "-std=c++17", "-Wall", "-o", str(exe),
                            os.path.join(ROOT, "tests", "native", "preprocess_asan.cpp")], capture_output=True, text=True)
    if build.returncode != 0 and "asan" in (build.stderr or "").lower():
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "asan ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_reference_processor_names_and_defaults():
    """processing_bailingmm.py:80-123: class names, constructor defaults (CLIP statistics, 224) and the 0.5 / 0.5
    instances BailingMMProcessor creates (:175-176)."""
    from ming_univision_b200.processing_bailingmm import (MingTokCenterCropProcessor, MingTokUndProcessor,
                                                          install_gpu_image_processors)

    und, gen = MingTokUndProcessor(), MingTokCenterCropProcessor()
    for p in (und, gen):
        assert p.image_size == 224 and p.mean == CLIP_MEAN and p.std == CLIP_STD

    class Holder:
        vis_processor = gen_processor = None

    h = install_gpu_image_processors(Holder())
    assert (h.vis_processor.image_size, h.gen_processor.image_size) == (1024, 512)
    assert h.vis_processor.mean == HALF and h.gen_processor.std == HALF


def test_emulated_unpatchify_to_u8_equals_unpatchify_clamp_then_tensor_to_pil(emu):
    """mb_unpatchify_to_u8 == the reference's tail composed: unpatchify (vision_transformer.py:515-527: rows
    [B, g*g, p*p*3], channel-last inside the patch) -> clamp_(-1, 1) (modeling_mingtok.py:194) -> tensor_to_pil."""
    B, g, p = 2, 5, 4
    gen = torch.Generator().manual_seed(4)
    x = (torch.randn((B, g * g, p * p * 3), generator=gen) * 0.8).to(torch.bfloat16)
    x[0, 0, :4] = torch.tensor([-1.5, 1.5, 1.0, -1.0], dtype=torch.bfloat16)
    # the oracle's restatement of :515-527 (oracle/mingtok_oracle.py, pixel_decoder_forward)
    img = torch.einsum("nhwpqc->nchpwq", x.float().reshape(B, g, g, p, p, 3)).reshape(B, 3, g * p, g * p).clamp(-1, 1)
    want = np.stack([po.postprocess(img[i].numpy()) for i in range(B)])
    bits = np.ascontiguousarray(x.view(torch.int16).numpy().view(np.uint16))
    out = np.full((B, g * p, g * p, 3), 7, dtype=np.uint8)
    emu.emu_unpatchify_to_u8(bits.ctypes.data_as(C.c_void_p), B, g, p, (C.c_float * 3)(*HALF), (C.c_float * 3)(*HALF),
                             out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, want)
