// Image pre- / post-processing on the GPU: Pillow's antialiased bicubic resize in its 8-bit fixed
// point, torchvision's centre crop, ToTensor and Normalize on the way in; denormalise + ToPILImage's truncation on the
// way out.  Byte work bounded by HBM / launch latency: no tensor cores.  The arithmetic is in preprocess_core.h (shared
// with the CPU emulation that tests it bit-exactly against Pillow); this file only maps threads onto it.
//
//   mb_image_preprocess_u8 : coefficient tables (one thread per output column / row, IEEE double, no FMA contraction)
//                            -> horizontal pass (input row segments staged in shared memory, one output pixel per
//                               thread, only the columns and rows the crop keeps) -> u8 scratch [n, rows, out_w, 3]
//                            -> vertical pass fused with crop + /255 + (x - mean) / std + NCHW store (bf16 or fp32),
//                               or with a plain u8 HWC store (out_kind 2: PIL's Image.resize + crop on the device)
//   mb_image_postprocess_u8: [n, 3, h, w] in [-1, 1] -> [n, h, w, 3] u8
//   mb_unpatchify_to_u8    : the pixel decoder's head rows -> unpatchify + clamp + the same u8 conversion in ONE pass
#include <cuda_bf16.h>

#include "common.h"
#include "preprocess_core.h"

namespace {

using mbpre::Plan;

__global__ void __launch_bounds__(128) resample_coeffs_kernel(Plan p, int32_t* bounds_h, int32_t* kk_h,
                                                              int32_t* bounds_v, int32_t* kk_v) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  mbpre::coeff_entry(p, idx, bounds_h, kk_h, bounds_v, kk_v);
}

__global__ void __launch_bounds__(mbpre::kHThreads) resample_h_kernel(Plan p, const uint8_t* __restrict__ src,
                                                                       const int32_t* __restrict__ bounds_h,
                                                                       const int32_t* __restrict__ kk_h,
                                                                       uint8_t* __restrict__ temp) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
  if (!mbpre::h_phase_load(p, src, bounds_h, bx, by, bz, threadIdx.x, blockDim.x, smem)) __trap();
  __syncthreads();
  mbpre::h_phase_compute(p, src, bounds_h, kk_h, bx, by, bz, threadIdx.x, blockDim.x, smem, temp);
}

// kOut: 0 = bf16 NCHW, 1 = fp32 NCHW (both normalised), 2 = u8 HWC (the resized + cropped image itself)
template <int kOut>
__global__ void __launch_bounds__(mbpre::kVThreads) resample_v_normalize_kernel(
    Plan p, const uint8_t* __restrict__ src, const uint8_t* __restrict__ temp, const int32_t* __restrict__ bounds_v,
    const int32_t* __restrict__ kk_v, float m0, float m1, float m2, float s0, float s1, float s2, void* out_) {
  const int xl = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, img = blockIdx.z;
  if (xl >= p.out_w) return;
  if (kOut == 2) {
    uint8_t u[3];
    mbpre::v_pixel_u8(p, src, temp, bounds_v, kk_v, img, yy, xl, u);
    uint8_t* o = static_cast<uint8_t*>(out_) + ((static_cast<int64_t>(img) * p.out_h + yy) * p.out_w + xl) * 3;
    o[0] = u[0], o[1] = u[1], o[2] = u[2];
    return;
  }
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  float v[3];
  mbpre::v_pixel(p, src, temp, bounds_v, kk_v, img, yy, xl, mean, stdv, v);
  const int64_t plane = static_cast<int64_t>(p.out_h) * p.out_w;
  const int64_t o = static_cast<int64_t>(img) * 3 * plane + static_cast<int64_t>(yy) * p.out_w + xl;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (kOut == 1) static_cast<float*>(out_)[o + c * plane] = v[c];
    else static_cast<__nv_bfloat16*>(out_)[o + c * plane] = __float2bfloat16_rn(v[c]);
  }
}

template <bool kF32>
__global__ void __launch_bounds__(256) image_to_u8_kernel(const void* __restrict__ img_, uint8_t* __restrict__ out,
                                                          int64_t n_px, int64_t plane, float m0, float m1, float m2,
                                                          float s0, float s1, float s2) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // pixel index over [n, h, w]
  if (i >= n_px) return;
  const int64_t img = i / plane, r = i - img * plane;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int64_t src = (img * 3 + c) * plane + r;
    const float x = kF32 ? static_cast<const float*>(img_)[src]
                         : __bfloat162float(static_cast<const __nv_bfloat16*>(img_)[src]);
    out[i * 3 + c] = mbpre::denormalize_to_u8(x, mean[c], stdv[c]);
  }
}

__global__ void __launch_bounds__(256) unpatchify_to_u8_kernel(const uint16_t* __restrict__ x,
                                                               uint8_t* __restrict__ out, int g, int p,
                                                               int64_t total_pix, float m0, float m1, float m2,
                                                               float s0, float s1, float s2) {
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total_pix;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    mbpre::unpatchify_u8_pixel(x, g, p, idx, mean, stdv, out);
}

}  // namespace

using namespace mb;

extern "C" int mb_image_preprocess_workspace_bytes(int n, int in_h, int in_w, int res_h, int res_w, int crop_top,
                                                   int crop_left, int out_h, int out_w, int64_t* bytes) {
  Plan p;
  const int bad = mbpre::make_plan(n, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w, &p);
  MB_CHECK_ARG(bad != 5, MB_ERR_SHAPE,
               "mb_image_preprocess_workspace_bytes: %dx%d is more than 100 times taller than wide and shrinks "
               "vertically; Pillow resizes such images height-first, which this path does not reproduce", in_h, in_w);
  MB_CHECK_ARG(bad == 0 && bytes != nullptr, MB_ERR_SHAPE,
               "mb_image_preprocess_workspace_bytes: invalid geometry (check %d): in %dx%d resized %dx%d crop (%d,%d) "
               "%dx%d", bad, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w);
  *bytes = p.total_bytes;
  return MB_OK;
}

extern "C" int mb_image_preprocess_u8(const void* src_, int n, int in_h, int in_w, int res_h, int res_w, int crop_top,
                                      int crop_left, int out_h, int out_w, float mean0, float mean1, float mean2,
                                      float std0, float std1, float std2, void* out, int out_kind, void* workspace,
                                      int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_image_preprocess_u8: no sm_100 device");
  Plan p;
  const int bad = mbpre::make_plan(n, in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w, &p);
  MB_CHECK_ARG(bad != 5, MB_ERR_SHAPE,
               "mb_image_preprocess_u8: %dx%d is more than 100 times taller than wide and shrinks vertically; Pillow "
               "resizes such images height-first, which this path does not reproduce", in_h, in_w);
  MB_CHECK_ARG(bad == 0, MB_ERR_SHAPE,
               "mb_image_preprocess_u8: invalid geometry (check %d): in %dx%d resized %dx%d crop (%d,%d) %dx%d", bad,
               in_h, in_w, res_h, res_w, crop_top, crop_left, out_h, out_w);
  MB_CHECK_ARG(out_kind >= 0 && out_kind <= 2, MB_ERR_SHAPE, "mb_image_preprocess_u8: out_kind must be 0, 1 or 2");
  MB_CHECK_ARG(out_kind == 2 || (std0 != 0.f && std1 != 0.f && std2 != 0.f), MB_ERR_SHAPE,
               "mb_image_preprocess_u8: std must be non-zero");
  MB_CHECK_ARG(workspace_bytes >= p.total_bytes, MB_ERR_SHAPE,
               "mb_image_preprocess_u8: workspace of %lld bytes, %lld needed", (long long)workspace_bytes,
               (long long)p.total_bytes);
  MB_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, MB_ERR_ALIGN,
               "mb_image_preprocess_u8: workspace must be 16-byte aligned");
  MB_CHECK_ARG(out_h <= 65535 && n <= 65535 && (p.rows + p.tile_rows - 1) / p.tile_rows <= 65535, MB_ERR_SHAPE,
               "mb_image_preprocess_u8: more than 65535 rows or images per call");
  if (n == 0) return MB_OK;
  const uint8_t* src = static_cast<const uint8_t*>(src_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int32_t* bounds_h = reinterpret_cast<int32_t*>(ws + p.off_bounds_h);
  int32_t* kk_h = reinterpret_cast<int32_t*>(ws + p.off_kk_h);
  int32_t* bounds_v = reinterpret_cast<int32_t*>(ws + p.off_bounds_v);
  int32_t* kk_v = reinterpret_cast<int32_t*>(ws + p.off_kk_v);
  uint8_t* temp = ws + p.off_temp;

  if (p.do_h || p.do_v) {
    resample_coeffs_kernel<<<(out_w + out_h + 127) / 128, 128, 0, stream>>>(p, bounds_h, kk_h, bounds_v, kk_v);
    MB_CHECK_CUDA(cudaGetLastError());
  }
  if (p.do_h) {
    const dim3 grid((out_w + p.tile_w - 1) / p.tile_w, (p.rows + p.tile_rows - 1) / p.tile_rows, n);
    resample_h_kernel<<<grid, mbpre::kHThreads, static_cast<size_t>(p.smem_row_bytes) * p.tile_rows, stream>>>(
        p, src, bounds_h, kk_h, temp);
    MB_CHECK_CUDA(cudaGetLastError());
  }
  const dim3 vgrid((out_w + mbpre::kVThreads - 1) / mbpre::kVThreads, out_h, n);
  if (out_kind == 1)
    resample_v_normalize_kernel<1><<<vgrid, mbpre::kVThreads, 0, stream>>>(
        p, src, temp, bounds_v, kk_v, mean0, mean1, mean2, std0, std1, std2, out);
  else if (out_kind == 2)
    resample_v_normalize_kernel<2><<<vgrid, mbpre::kVThreads, 0, stream>>>(
        p, src, temp, bounds_v, kk_v, mean0, mean1, mean2, std0, std1, std2, out);
  else
    resample_v_normalize_kernel<0><<<vgrid, mbpre::kVThreads, 0, stream>>>(
        p, src, temp, bounds_v, kk_v, mean0, mean1, mean2, std0, std1, std2, out);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_image_postprocess_u8(const void* img, int img_is_fp32, int n, int h, int w, float mean0, float mean1,
                                       float mean2, float std0, float std1, float std2, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_image_postprocess_u8: no sm_100 device");
  MB_CHECK_ARG(n >= 0 && h >= 1 && w >= 1, MB_ERR_SHAPE, "mb_image_postprocess_u8: invalid shape %d x %d x %d", n, h, w);
  const int64_t plane = static_cast<int64_t>(h) * w, n_px = plane * n;
  if (n_px == 0) return MB_OK;
  const unsigned blocks = static_cast<unsigned>((n_px + 255) / 256);
  if (img_is_fp32)
    image_to_u8_kernel<true><<<blocks, 256, 0, stream>>>(img, static_cast<uint8_t*>(out), n_px, plane, mean0, mean1,
                                                         mean2, std0, std1, std2);
  else
    image_to_u8_kernel<false><<<blocks, 256, 0, stream>>>(img, static_cast<uint8_t*>(out), n_px, plane, mean0, mean1,
                                                          mean2, std0, std1, std2);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_unpatchify_to_u8(const void* x, void* out, int B, int g, int p, float mean0, float mean1,
                                   float mean2, float std0, float std1, float std2, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_unpatchify_to_u8: no sm_100 device");
  MB_CHECK_ARG(B >= 0 && g >= 1 && p >= 1, MB_ERR_SHAPE, "mb_unpatchify_to_u8: invalid shape B=%d g=%d p=%d", B, g, p);
  const int64_t total = static_cast<int64_t>(B) * g * p * g * p;  // pixels; each thread writes 3 adjacent bytes
  if (total == 0) return MB_OK;
  const int64_t want = (total + 255) / 256;
  const int cap = num_sms() * 16;
  unpatchify_to_u8_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, stream>>>(
      static_cast<const uint16_t*>(x), static_cast<uint8_t*>(out), g, p, total, mean0, mean1, mean2, std0, std1, std2);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}
