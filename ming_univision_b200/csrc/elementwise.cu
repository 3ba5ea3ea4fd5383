// HBM-bound row / elementwise operators of the MingTok path: LayerNorm (warp per row, shuffle reductions),
// im2row, cls-row fill, group mean, affine, in-projection with repeat shortcut, pixel shuffle, unpatchify+clamp.
// All use 16-byte vector accesses where the layout allows and grids sized from the row / element count.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]); q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]); q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, NV 16-byte vectors per lane held in registers, two-pass fp32 statistics.
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ gamma,
                 const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int dim,
                 float eps, int act, int rows_per_group, int64_t group_stride_x) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* xr = (rows_per_group > 0)
                                ? x + static_cast<int64_t>(row / rows_per_group) * group_stride_x +
                                      static_cast<int64_t>(row % rows_per_group) * ldx
                                : x + static_cast<int64_t>(row) * ldx;
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    if (col < dim) {
      const uint4 q = *reinterpret_cast<const uint4*>(xr + col);
      unpack8(q, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    if (col < dim) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(dim) + eps);
  __nv_bfloat16* yr = y + static_cast<int64_t>(row) * ldy;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    if (col < dim) {
      float g[8], b[8], o[8];
      if (gamma != nullptr) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + col)), g);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 1.f;
      }
      if (beta != nullptr) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(beta + col)), b);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        if (act == 1) o[j] = gelu_erf(bf16_round(o[j]));
      }
      *reinterpret_cast<uint4*>(yr + col) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// im2row for the non-overlapping patch conv
// ------------------------------------------------------------------------------------------------------------
template <bool kFp32In>
__global__ void patchify_kernel(const void* __restrict__ img_, __nv_bfloat16* __restrict__ rows, int B, int C, int Hh,
                                int Ww, int P, int64_t total_vec) {
  const int gw = Ww / P, gh = Hh / P;
  const int kcols = C * P * P;
  const int vec_per_row = kcols / 8;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total_vec;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / vec_per_row;
    const int col = static_cast<int>(idx % vec_per_row) * 8;
    const int b = static_cast<int>(row / (gh * gw));
    const int pr = static_cast<int>(row % (gh * gw));
    const int gy = pr / gw, gx = pr % gw;
    const int c = col / (P * P);
    const int py = (col % (P * P)) / P;
    const int px = col % P;
    const int64_t src = ((static_cast<int64_t>(b) * C + c) * Hh + (gy * P + py)) * Ww + gx * P + px;
    float f[8];
    if (kFp32In) {
      const float4* s = reinterpret_cast<const float4*>(static_cast<const float*>(img_) + src);
      const float4 a = s[0], d = s[1];
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = d.x; f[5] = d.y; f[6] = d.z; f[7] = d.w;
    } else {
      unpack8(*reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(img_) + src), f);
    }
    *reinterpret_cast<uint4*>(rows + row * kcols + col) = pack8(f);
  }
}

__global__ void fill_cls_row_kernel(__nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ cls,
                                    const __nv_bfloat16* __restrict__ pos_cls, int B, int n_plus_1, int dim) {
  const int b = blockIdx.x;
  __nv_bfloat16* dst = x + (static_cast<int64_t>(b) * n_plus_1 + (n_plus_1 - 1)) * dim;
  for (int c = threadIdx.x; c < dim; c += blockDim.x)
    dst[c] = __float2bfloat16_rn(__bfloat162float(cls[c]) + __bfloat162float(pos_cls[c]));
}

__global__ void group_mean_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                  int rows, int dim, int groups) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<int64_t>(rows) * groups) return;
  const int64_t r = idx / groups;
  const int c = static_cast<int>(idx % groups);
  const int g = dim / groups;
  const __nv_bfloat16* p = x + r * ldx + c * g;
  float s = 0.f;
  for (int j = 0; j < g; ++j) s += __bfloat162float(p[j]);
  out[idx] = __float2bfloat16_rn(s / static_cast<float>(g));
}

template <bool kFp32In>
__global__ void affine_kernel(const void* __restrict__ x_, __nv_bfloat16* __restrict__ y, int64_t n, float scale,
                              float shift) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = kFp32In ? static_cast<const float*>(x_)[i] : __bfloat162float(static_cast<const __nv_bfloat16*>(x_)[i]);
    y[i] = __float2bfloat16_rn(v * scale + shift);
  }
}

// in_proj (K = in_dim <= 64) + repeat shortcut.  The whole weight ([dim, in_dim] bf16, 64 KB for 1024 x 32) is staged
// TRANSPOSED in shared memory once per CTA and reused for kRowsPerCta rows, so it is read from L2 ~130 times per call
// instead of once per row (the first version spent 256 us re-reading it 4160 times).
constexpr int kInprojRows = 32;
__global__ void __launch_bounds__(256)
inproj_repeat_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ W,
                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, int rows, int in_dim,
                     int dim) {
  // One CTA = kInprojRows rows x all `dim` columns.  Thread j owns output column j (and j + 256, ...): it reads its weight
  // row W[j][0:in_dim] straight from global memory (64 contiguous bytes, L2-resident after the first CTA) into registers
  // and walks the CTA's rows, whose inputs sit in shared memory (broadcast reads); consecutive threads write
  // consecutive columns, so the stores are fully coalesced.  (The first version transposed W through shared memory in
  // every CTA with bank-conflicted scatter writes: 200 us for 4160 rows; this one is bound by the 8.5 MB it writes.)
  extern __shared__ __align__(16) uint8_t inproj_smem[];
  float* sx = reinterpret_cast<float*>(inproj_smem);  // [kInprojRows][in_dim]
  pdl_launch_dependents();
  pdl_wait();
  const int r0 = blockIdx.x * kInprojRows;
  const int nr = min(kInprojRows, rows - r0);
  for (int i = threadIdx.x; i < nr * in_dim; i += blockDim.x)
    sx[i] = __bfloat162float(x[static_cast<int64_t>(r0) * in_dim + i]);
  __syncthreads();
  const int rep = dim / in_dim;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    const float bj = b ? __bfloat162float(b[j]) : 0.f;
    float wj[64];
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8) {
      if (k8 * 8 < in_dim) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(W + static_cast<int64_t>(j) * in_dim + k8 * 8));
        const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
        wj[k8 * 8 + 0] = f0.x; wj[k8 * 8 + 1] = f0.y; wj[k8 * 8 + 2] = f1.x; wj[k8 * 8 + 3] = f1.y;
        wj[k8 * 8 + 4] = f2.x; wj[k8 * 8 + 5] = f2.y; wj[k8 * 8 + 6] = f3.x; wj[k8 * 8 + 7] = f3.y;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) wj[k8 * 8 + i] = 0.f;
      }
    }
    const int src = j / rep;
    for (int r = 0; r < nr; ++r) {
      const float* xr = sx + r * in_dim;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k < in_dim) acc = fmaf(wj[k], xr[k], acc);
      const float lin = bf16_round(acc + bj);
      out[static_cast<int64_t>(r0 + r) * dim + j] = __float2bfloat16_rn(lin + xr[src]);
    }
  }
}

__global__ void pixel_shuffle_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                     int g, int f, int C, int64_t total_vec) {
  const int cv = C / 8;
  const int gf = g * f;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total_vec;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(idx % cv);
    const int64_t tok = idx / cv;  // output token over B * gf * gf
    const int b = static_cast<int>(tok / (gf * gf));
    const int t = static_cast<int>(tok % (gf * gf));
    const int oy = t / gf, ox = t % gf;
    const int h = oy / f, xx = oy % f, w = ox / f, yy = ox % f;
    const int64_t src = ((static_cast<int64_t>(b) * g * g + h * g + w) * (f * f) + (xx * f + yy)) * C + c8 * 8;
    *reinterpret_cast<uint4*>(out + tok * C + c8 * 8) = *reinterpret_cast<const uint4*>(in + src);
  }
}

template <bool kFp32Out>
__global__ void unpatchify_clamp_kernel(const __nv_bfloat16* __restrict__ x, void* __restrict__ img_, int B, int g,
                                        int p, int64_t total_pix) {
  // one thread per output pixel: its 3 channels are adjacent in x (channel-last inside the patch) and go to the 3
  // planes of the NCHW image; consecutive threads walk along the image row, so both sides are coalesced
  const int HW = g * p;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total_pix;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int xw = static_cast<int>(idx % HW);
    const int yh = static_cast<int>((idx / HW) % HW);
    const int b = static_cast<int>(idx / (static_cast<int64_t>(HW) * HW));
    const int h = yh / p, pp = yh % p, w = xw / p, q = xw % p;
    const int64_t src = ((static_cast<int64_t>(b) * g * g + h * g + w) * (p * p) + pp * p + q) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __bfloat162float(x[src + c]);
      v = fminf(fmaxf(v, -1.f), 1.f);
      const int64_t dst = ((static_cast<int64_t>(b) * 3 + c) * HW + yh) * HW + xw;
      if (kFp32Out)
        static_cast<float*>(img_)[dst] = v;
      else
        static_cast<__nv_bfloat16*>(img_)[dst] = __float2bfloat16_rn(v);
    }
  }
}

// stats[r] = (sum_c x[r, c], sum_c x[r, c]^2)   one warp per row — seeds the LayerNorm-folded GEMM chain of a stage
__global__ void __launch_bounds__(256)
row_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float* __restrict__ stats, int rows, int dim) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + static_cast<int64_t>(row) * ldx;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane * 8; c < dim; c += 256) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(xr + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { stats[2 * row] = s1; stats[2 * row + 1] = s2; }
}

static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace mb

using namespace mb;

extern "C" int mb_layernorm(const void* x, int64_t ldx, const void* gamma, const void* beta, void* y, int64_t ldy,
                            int rows, int dim, float eps, int act, int rows_per_group, int64_t group_stride_x,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_layernorm: no sm_100 device");
  MB_CHECK_ARG(rows >= 0 && dim >= 8 && dim % 8 == 0 && dim <= 4096, MB_ERR_SHAPE,
               "mb_layernorm: dim must be a multiple of 8 in [8, 4096] (dim=%d)", dim);
  MB_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && group_stride_x % 8 == 0, MB_ERR_ALIGN,
               "mb_layernorm: ldx/ldy/group_stride_x must be multiples of 8");
  if (rows == 0) return MB_OK;
  const int nv = (dim + 255) / 256;
  const int wpb = 8;
  const dim3 grid((rows + wpb - 1) / wpb), block(wpb * 32);
  const __nv_bfloat16* xx = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* gg = static_cast<const __nv_bfloat16*>(gamma);
  const __nv_bfloat16* bb = static_cast<const __nv_bfloat16*>(beta);
  __nv_bfloat16* yy = static_cast<__nv_bfloat16*>(y);
#define MB_LN(NV_)                                                                                           \
  MB_CHECK_CUDA(launch_pdl(layernorm_kernel<NV_>, grid, block, 0, stream, xx, ldx, gg, bb, yy, ldy, rows, dim, eps,  \
                           act, rows_per_group, group_stride_x))
  if (nv <= 1) MB_LN(1);
  else if (nv <= 2) MB_LN(2);
  else if (nv <= 3) MB_LN(3);
  else if (nv <= 4) MB_LN(4);
  else if (nv <= 8) MB_LN(8);
  else if (nv <= 12) MB_LN(12);
  else MB_LN(16);
#undef MB_LN
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_row_stats(const void* x, int64_t ldx, float* stats, int rows, int dim, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_row_stats: no sm_100 device");
  MB_CHECK_ARG(dim % 8 == 0 && ldx % 8 == 0, MB_ERR_ALIGN, "mb_row_stats: dim and ldx must be multiples of 8");
  if (rows == 0) return MB_OK;
  row_stats_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, stats, rows, dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_patchify(const void* img, int img_is_fp32, void* rows, int B, int C, int Hh, int Ww, int P,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_patchify: no sm_100 device");
  MB_CHECK_ARG(B >= 0 && C >= 1 && P % 8 == 0 && Hh % P == 0 && Ww % P == 0, MB_ERR_SHAPE,
               "mb_patchify: H, W must be multiples of P and P of 8 (H=%d W=%d P=%d)", Hh, Ww, P);
  const int64_t total_vec = static_cast<int64_t>(B) * (Hh / P) * (Ww / P) * C * P * P / 8;
  if (total_vec == 0) return MB_OK;
  const int grid = grid_for(total_vec, 256);
  if (img_is_fp32)
    patchify_kernel<true><<<grid, 256, 0, stream>>>(img, static_cast<__nv_bfloat16*>(rows), B, C, Hh, Ww, P, total_vec);
  else
    patchify_kernel<false><<<grid, 256, 0, stream>>>(img, static_cast<__nv_bfloat16*>(rows), B, C, Hh, Ww, P, total_vec);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_fill_cls_row(void* x, const void* cls, const void* pos_cls, int B, int n_plus_1, int dim,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_fill_cls_row: no sm_100 device");
  if (B == 0) return MB_OK;
  fill_cls_row_kernel<<<B, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(cls),
                                             static_cast<const __nv_bfloat16*>(pos_cls), B, n_plus_1, dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_group_mean(const void* x, int64_t ldx, void* out, int rows, int dim, int groups, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_group_mean: no sm_100 device");
  MB_CHECK_ARG(groups >= 1 && dim % groups == 0, MB_ERR_SHAPE, "mb_group_mean: dim %% groups != 0");
  const int64_t n = static_cast<int64_t>(rows) * groups;
  if (n == 0) return MB_OK;
  group_mean_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), rows, dim, groups);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_affine(const void* x, int x_is_fp32, void* y, int64_t n, float scale, float shift, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_affine: no sm_100 device");
  if (n == 0) return MB_OK;
  if (x_is_fp32)
    affine_kernel<true><<<grid_for(n, 256), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y), n, scale, shift);
  else
    affine_kernel<false><<<grid_for(n, 256), 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y), n, scale, shift);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_inproj_repeat(const void* x, const void* W, const void* b, void* out, int rows, int in_dim, int dim,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_inproj_repeat: no sm_100 device");
  MB_CHECK_ARG(in_dim >= 1 && in_dim <= 64 && dim % in_dim == 0, MB_ERR_SHAPE,
               "mb_inproj_repeat: need in_dim <= 64 and dim %% in_dim == 0 (in_dim=%d dim=%d)", in_dim, dim);
  if (rows == 0) return MB_OK;
  MB_CHECK_ARG(in_dim % 8 == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, MB_ERR_ALIGN,
               "mb_inproj_repeat: in_dim %% 8 == 0 and a 16-byte aligned weight are required");
  const size_t smem = static_cast<size_t>(kInprojRows) * in_dim * 4;
  MB_CHECK_CUDA(launch_pdl(inproj_repeat_kernel, dim3((rows + kInprojRows - 1) / kInprojRows), dim3(256), smem, stream,
                           static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(W),
                           static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(out), rows, in_dim, dim));
  return MB_OK;
}

extern "C" int mb_pixel_shuffle(const void* in, void* out, int B, int g, int f, int C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_pixel_shuffle: no sm_100 device");
  MB_CHECK_ARG(C % 8 == 0 && g >= 1 && f >= 1, MB_ERR_SHAPE, "mb_pixel_shuffle: C %% 8 != 0");
  const int64_t total_vec = static_cast<int64_t>(B) * g * g * f * f * (C / 8);
  if (total_vec == 0) return MB_OK;
  pixel_shuffle_kernel<<<grid_for(total_vec, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), B, g, f, C, total_vec);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_unpatchify_clamp(const void* x, void* img, int out_is_fp32, int B, int g, int p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_unpatchify_clamp: no sm_100 device");
  const int64_t total = static_cast<int64_t>(B) * g * p * g * p;  // pixels; each thread writes the 3 channels
  if (total == 0) return MB_OK;
  if (out_is_fp32)
    unpatchify_clamp_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), img,
                                                                            B, g, p, total);
  else
    unpatchify_clamp_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                             img, B, g, p, total);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}
