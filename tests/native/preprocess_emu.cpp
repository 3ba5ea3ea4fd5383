// CPU emulation of the launches in ming_univision_b200/csrc/preprocess.cu: the SAME per-thread bodies
// (preprocess_core.h), with the grids, blocks and the shared-memory barrier walked by plain loops.  Test infrastructure
// only (built by tests/test_preprocess_cpu.py with g++ -ffp-contract=off into a temporary directory); it lets the
// CPU suite check every index computation and rounding step of the device path against Pillow and torchvision.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ming_univision_b200/csrc/preprocess_core.h"

using mbpre::Plan;

extern "C" long long emu_workspace_bytes(int n, int in_h, int in_w, int res_h, int res_w, int crop_top, int crop_left,
                                         int out_h, int out_w) {
  Plan p;
  if (mbpre::make_plan(n, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w, &p) != 0) return -1;
  return p.total_bytes;
}

// out: fp32 [n, 3, out_h, out_w] (out_kind 1); u8_out (optional): the resized + cropped u8 image [n, out_h, out_w, 3]
// (out_kind 2).
// plan_out (optional, 8 ints): do_h, do_v, ksize_h, ksize_v, row0, rows, tile_w, tile_rows.
extern "C" int emu_image_preprocess_ex(const uint8_t* src, int n, int in_h, int in_w, int res_h, int res_w,
                                       int crop_top, int crop_left, int out_h, int out_w, const float* mean,
                                       const float* stdv, float* out, uint8_t* u8_out, int* plan_out) {
  Plan p;
  const int bad = mbpre::make_plan(n, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w, &p);
  if (bad) return -bad;
  if (plan_out) {
    const int v[8] = {p.do_h, p.do_v, p.ksize_h, p.ksize_v, p.row0, p.rows, p.tile_w, p.tile_rows};
    memcpy(plan_out, v, sizeof(v));
  }
  // the workspace is filled with garbage first: nothing may depend on its initial contents
  // (exact size, 16-byte aligned like a device allocation: an access past the end is visible to ASAN)
  uint8_t* base = static_cast<uint8_t*>(aligned_alloc(16, static_cast<size_t>(p.total_bytes) + (p.total_bytes == 0 ? 16 : 0)));
  if (!base) return -101;
  memset(base, 0xA5, static_cast<size_t>(p.total_bytes));
  struct Free { uint8_t* p; ~Free() { free(p); } } free_ws{base};
  int32_t* bounds_h = reinterpret_cast<int32_t*>(base + p.off_bounds_h);
  int32_t* kk_h = reinterpret_cast<int32_t*>(base + p.off_kk_h);
  int32_t* bounds_v = reinterpret_cast<int32_t*>(base + p.off_bounds_v);
  int32_t* kk_v = reinterpret_cast<int32_t*>(base + p.off_kk_v);
  uint8_t* temp = base + p.off_temp;

  if (p.do_h || p.do_v) {  // resample_coeffs_kernel<<<(out_w + out_h + 127) / 128, 128>>>
    const int blocks = (out_w + out_h + 127) / 128;
    for (int b = 0; b < blocks; ++b)
      for (int t = 0; t < 128; ++t) mbpre::coeff_entry(p, b * 128 + t, bounds_h, kk_h, bounds_v, kk_v);
  }
  if (p.do_h) {  // resample_h_kernel<<<grid, kHThreads, smem_row_bytes * tile_rows>>>
    const int gx = (out_w + p.tile_w - 1) / p.tile_w, gy = (p.rows + p.tile_rows - 1) / p.tile_rows;
    uint8_t* smem = static_cast<uint8_t*>(aligned_alloc(16, static_cast<size_t>(p.smem_row_bytes) * p.tile_rows));
    if (!smem) return -101;
    struct FreeS { uint8_t* p; ~FreeS() { free(p); } } free_smem{smem};
    for (int bz = 0; bz < n; ++bz)
      for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
          memset(smem, 0x5A, static_cast<size_t>(p.smem_row_bytes) * p.tile_rows);
          for (int t = 0; t < mbpre::kHThreads; ++t)
            if (!mbpre::h_phase_load(p, src, bounds_h, bx, by, bz, t, mbpre::kHThreads, smem)) return -100;
          // __syncthreads()
          for (int t = 0; t < mbpre::kHThreads; ++t)
            mbpre::h_phase_compute(p, src, bounds_h, kk_h, bx, by, bz, t, mbpre::kHThreads, smem, temp);
        }
  }
  // resample_v_normalize_kernel<true><<<(ceil(out_w / kVThreads), out_h, n), kVThreads>>>
  const int vx = (out_w + mbpre::kVThreads - 1) / mbpre::kVThreads;
  const long long plane = static_cast<long long>(out_h) * out_w;
  for (int img = 0; img < n; ++img)
    for (int yy = 0; yy < out_h; ++yy)
      for (int b = 0; b < vx; ++b)
        for (int t = 0; t < mbpre::kVThreads; ++t) {
          const int xl = b * mbpre::kVThreads + t;
          if (xl >= out_w) continue;
          float v[3];
          mbpre::v_pixel(p, src, temp, bounds_v, kk_v, img, yy, xl, mean, stdv, v);
          for (int c = 0; c < 3; ++c) out[(img * 3LL + c) * plane + static_cast<long long>(yy) * out_w + xl] = v[c];
          if (u8_out) {  // resample_v_normalize_kernel<2>: the u8 image itself, HWC
            uint8_t u[3];
            mbpre::v_pixel_u8(p, src, temp, bounds_v, kk_v, img, yy, xl, u);
            for (int c = 0; c < 3; ++c) u8_out[((static_cast<long long>(img) * out_h + yy) * out_w + xl) * 3 + c] = u[c];
          }
        }
  return 0;
}

extern "C" int emu_image_preprocess(const uint8_t* src, int n, int in_h, int in_w, int res_h, int res_w, int crop_top,
                                    int crop_left, int out_h, int out_w, const float* mean, const float* stdv,
                                    float* out, int* plan_out) {
  return emu_image_preprocess_ex(src, n, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w, mean, stdv, out,
                                 nullptr, plan_out);
}

// image_to_u8_kernel<true>
extern "C" void emu_image_postprocess(const float* img, int n, int h, int w, const float* mean, const float* stdv,
                                      uint8_t* out) {
  const long long plane = static_cast<long long>(h) * w, n_px = plane * n;
  for (long long i = 0; i < n_px; ++i) {
    const long long im = i / plane, r = i - im * plane;
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = mbpre::denormalize_to_u8(img[(im * 3 + c) * plane + r], mean[c], stdv[c]);
  }
}

// Coefficient table of one axis, for a direct comparison with the oracle's tables.
extern "C" int emu_axis_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk) {
  const mbpre::AxisGeom g = mbpre::axis_geom(in_size, out_size);
  for (int xx = 0; xx < out_size; ++xx) mbpre::axis_coeffs(g, in_size, xx, bounds + 2 * xx, kk + static_cast<long long>(xx) * g.ksize);
  return g.ksize;
}

// unpatchify_to_u8_kernel: grid-stride loop over the pixels of [B, g*p, g*p]
extern "C" void emu_unpatchify_to_u8(const uint16_t* x, int B, int g, int p, const float* mean, const float* stdv,
                                     uint8_t* out) {
  const long long total = static_cast<long long>(B) * g * p * g * p;
  const long long threads = 7 * 256;  // a grid smaller than the problem, so the stride loop is exercised
  for (long long t = 0; t < threads; ++t)
    for (long long idx = t; idx < total; idx += threads) mbpre::unpatchify_u8_pixel(x, g, p, idx, mean, stdv, out);
}
