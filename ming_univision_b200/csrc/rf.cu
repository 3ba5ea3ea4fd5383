// Row-wise helpers of the rectified-flow head (mingunivision/diff_loss_rf_swiglu.py): adaLN modulation, the
// step-batched SiLU(t_emb + c) conditioning rows, and the CFG-combine + Euler update.  M = CFG rows (<= 3) here, so
// these are latency-sized kernels; the weight streaming happens in gemv.cu.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += scratch[w];
  return t;
}

// y[m] = bf16( (LN(x[m]) * gamma + beta) * bf16(1 + scale[m]) + shift[m] )     one CTA (256 threads) per row
// modulate(): diff_loss_rf_swiglu.py:184-185, used at :270 (ResBlock) and :290 (FinalLayer, gamma/beta = NULL).
__global__ void __launch_bounds__(256)
adaln_modulate_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ gamma,
                      const __nv_bfloat16* __restrict__ beta, const __nv_bfloat16* __restrict__ shift, int64_t lds,
                      const __nv_bfloat16* __restrict__ scale, int64_t ldsc, __nv_bfloat16* __restrict__ y, int64_t ldy,
                      int dim, float eps) {
  __shared__ float scratch[8];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const __nv_bfloat16* xr = x + m * ldx;
  float s = 0.f;
  for (int i = threadIdx.x; i < dim; i += 256) s += __bfloat162float(xr[i]);
  const float mean = block_sum_256(s, scratch) / dim;
  float q = 0.f;
  for (int i = threadIdx.x; i < dim; i += 256) {
    const float d = __bfloat162float(xr[i]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum_256(q, scratch) / dim + eps);
  for (int i = threadIdx.x; i < dim; i += 256) {
    float v = (__bfloat162float(xr[i]) - mean) * rstd;
    if (gamma) v = v * __bfloat162float(gamma[i]) + (beta ? __bfloat162float(beta[i]) : 0.f);
    const float sc = bf16_round(1.f + __bfloat162float(scale[m * ldsc + i]));
    v = v * sc + __bfloat162float(shift[m * lds + i]);
    y[m * ldy + i] = __float2bfloat16_rn(v);
  }
}

// out[(s*B + b), :] = bf16(silu(bf16(temb[s, :] + c[b, :])))   — the input of every adaLN_modulation Linear
// (y = t + c, then nn.SiLU: diff_loss_rf_swiglu.py:376, 262-265, 283-286), for all sampling steps at once.
__global__ void silu_add_rows_kernel(const __nv_bfloat16* __restrict__ temb, const __nv_bfloat16* __restrict__ c,
                                     __nv_bfloat16* __restrict__ out, int steps, int B, int dim) {
  const int64_t total = static_cast<int64_t>(steps) * B * dim;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % dim);
    const int row = static_cast<int>(i / dim);
    const int s = row / B, b = row % B;
    const float yv = bf16_round(__bfloat162float(temb[static_cast<int64_t>(s) * dim + d]) +
                                __bfloat162float(c[static_cast<int64_t>(b) * dim + d]));
    out[i] = __float2bfloat16_rn(silu(yv));
  }
}

// CFG combine + explicit Euler step (RectifiedFlowLoss.sample, diff_loss_rf_swiglu.py:138-179).
// v: bf16 [B, C]; the B rows are B / cfg_rows independent samples of cfg_rows adjacent rows ordered (cond, uncond
// [, text_uncond]); x: fp32 [B, C], every row of a sample receives the same update.
// Rounding points mirror the reference's bf16 tensor arithmetic with Python-float scalars.
__global__ void rf_euler_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ x_bf16,
                                const __nv_bfloat16* __restrict__ v, int B, int cfg_rows, int C, float dt, float text_cfg,
                                float image_cfg) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (B / cfg_rows) * C) return;
  const int c = i % C, r0 = (i / C) * cfg_rows;
  const __nv_bfloat16* vs = v + static_cast<int64_t>(r0) * C;
  float step;
  if (cfg_rows == 3) {
    const float vc = __bfloat162float(vs[c]), vu = __bfloat162float(vs[C + c]), vt = __bfloat162float(vs[2 * C + c]);
    const float t3 = bf16_round(vu + bf16_round(image_cfg * bf16_round(vt - vu)));
    const float vg = bf16_round(t3 + bf16_round(text_cfg * bf16_round(vc - vt)));
    step = bf16_round(vg * dt);
  } else if (cfg_rows == 2) {
    const float vc = __bfloat162float(vs[c]), vu = __bfloat162float(vs[C + c]);
    const float vg = bf16_round(vu + bf16_round(text_cfg * bf16_round(vc - vu)));
    step = bf16_round(vg * dt);
  } else {
    step = bf16_round(__bfloat162float(vs[c]) * dt);
  }
  for (int b = 0; b < cfg_rows; ++b) {
    const float nx = x[(r0 + b) * C + c] + step;
    x[(r0 + b) * C + c] = nx;
    x_bf16[(r0 + b) * C + c] = __float2bfloat16_rn(nx);
  }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_adaln_modulate(const void* x, int64_t ldx, const void* gamma, const void* beta, const void* shift,
                                 int64_t ld_shift, const void* scale, int64_t ld_scale, void* y, int64_t ldy, int rows,
                                 int dim, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_adaln_modulate: no sm_100 device");
  MB_CHECK_ARG(rows >= 0 && dim >= 1 && shift && scale, MB_ERR_SHAPE, "mb_adaln_modulate: bad arguments");
  if (rows == 0) return MB_OK;
  MB_CHECK_CUDA(launch_pdl(adaln_modulate_kernel, dim3(rows), dim3(256), 0, stream,
                           static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(gamma),
                           static_cast<const __nv_bfloat16*>(beta), static_cast<const __nv_bfloat16*>(shift), ld_shift,
                           static_cast<const __nv_bfloat16*>(scale), ld_scale, static_cast<__nv_bfloat16*>(y), ldy,
                           dim, eps));
  return MB_OK;
}

extern "C" int mb_silu_add_rows(const void* temb, const void* c, void* out, int steps, int B, int dim, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_silu_add_rows: no sm_100 device");
  const int64_t total = static_cast<int64_t>(steps) * B * dim;
  if (total == 0) return MB_OK;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  silu_add_rows_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(temb),
                                                 static_cast<const __nv_bfloat16*>(c),
                                                 static_cast<__nv_bfloat16*>(out), steps, B, dim);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_rf_euler_step(void* x_f32, void* x_bf16, const void* v, int B, int cfg_rows, int C, float dt,
                                float text_cfg, float image_cfg, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_rf_euler_step: no sm_100 device");
  MB_CHECK_ARG(B >= 1 && B <= 8 && C >= 1 && cfg_rows >= 1 && cfg_rows <= 3 && B % cfg_rows == 0, MB_ERR_SHAPE,
               "mb_rf_euler_step: bad shape B=%d cfg_rows=%d C=%d", B, cfg_rows, C);
  const int n = (B / cfg_rows) * C;
  MB_CHECK_CUDA(launch_pdl(rf_euler_kernel, dim3((n + 63) / 64), dim3(64), 0, stream, static_cast<float*>(x_f32),
                           static_cast<__nv_bfloat16*>(x_bf16), static_cast<const __nv_bfloat16*>(v), B, cfg_rows, C, dt,
                           text_cfg, image_cfg));
  return MB_OK;
}
