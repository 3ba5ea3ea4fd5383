// Weight-streaming skinny GEMM for the decode regime (M <= 8 rows: CFG rows of the RF head, the AR step, the cached
// semantic-decoder step):  out[M, N'] = epilogue(A[M, K] @ W[N, K]^T + bias).
//
// HBM-bound by construction: every weight byte is read exactly once with 16-byte, fully coalesced, L1-bypassing
// loads; the tiny activation matrix lives in shared memory (bf16), accumulation is fp32 in registers, the K-split is
// reduced with warp shuffles.
//
// Algorithmic bytes per launch: N*K*2 (weights) — the roofline figure bench_rf.py reports against HBM peak.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

constexpr int kGemvThreads = 256;

struct GemvParams {
  const __nv_bfloat16* A; int64_t lda;
  const __nv_bfloat16* W; int64_t ldw;
  const __nv_bfloat16* bias;
  __nv_bfloat16* out; int64_t ldo;
  const __nv_bfloat16* res; int64_t ldr;    // RESIDUAL / GATED: residual stream
  const __nv_bfloat16* gate; int64_t ldg;   // GATED: per-element gate
  float* out_f32;                            // optional fp32 copy of the output (may be null)
  int M, N, K;
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void unpack8f(const uint4& q, float (&f)[8]) {
  const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// EPI: 0 bias, 1 gelu, 2 swiglu (W = [2H, K] reference layout, out has H columns), 3 residual, 4 silu, 5 gated residual
//
// Work decomposition: a CTA has 8 warps; KSPLIT consecutive warps share one output column and split its K range, so a
// CTA produces 8 / KSPLIT columns per iteration.  KSPLIT is chosen on the host so that even a narrow layer
// (w3: N = 3072, K = 8192) spreads over every warp slot of the GPU and each lane has all of its 16-byte loads in flight
// at once.  Activations are staged ONCE per CTA in shared memory as fp32 (no per-use bf16 unpacking), MT is the exact
// row count (no padding to a power of two): at M = 3 the inner loop is 8 unpack + 24 FMA + 6 LDS.128 per 16-byte
// weight chunk, ~50 % of the issue slots at HBM speed.
template <int MT, int EPI, int KSPLIT>
__global__ void __launch_bounds__(kGemvThreads)
gemv_bf16_kernel(const GemvParams p) {
  constexpr int kRows = (EPI == MB_EPI_SWIGLU) ? 2 : 1;  // weight rows per output column
  constexpr int kUnroll = (EPI == MB_EPI_SWIGLU) ? 4 : 8;
  constexpr int kColsPerIter = (kGemvThreads / 32) / KSPLIT;
  extern __shared__ __align__(16) uint8_t gemv_smem[];
  float* sA = reinterpret_cast<float*>(gemv_smem);  // [MT][K] fp32
  __shared__ float red[2][kGemvThreads / 32][kRows * MT];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, kchunks = K >> 3;
  const int n_out = (EPI == MB_EPI_SWIGLU) ? p.N / 2 : p.N;

  for (int i = tid; i < MT * kchunks; i += kGemvThreads) {
    const int m = i / kchunks, c = i % kchunks;
    float f[8];
    unpack8f(*reinterpret_cast<const uint4*>(p.A + m * p.lda + c * 8), f);
    float4* dst = reinterpret_cast<float4*>(sA + m * K + c * 8);
    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();

  const int col_slot = warp / KSPLIT, ks = warp % KSPLIT;
  const int per = (kchunks + KSPLIT - 1) / KSPLIT;
  const int c_beg = ks * per, c_end = min(kchunks, c_beg + per);
  const int iters = (n_out + gridDim.x * kColsPerIter - 1) / (gridDim.x * kColsPerIter);
  for (int it = 0; it < iters; ++it) {
    const int n_raw = (it * gridDim.x + blockIdx.x) * kColsPerIter + col_slot;
    const bool valid = n_raw < n_out;
    const int n = valid ? n_raw : n_out - 1;
    const uint4* wrow[kRows];
    wrow[0] = reinterpret_cast<const uint4*>(p.W + static_cast<int64_t>(n) * p.ldw);
    if constexpr (kRows == 2) wrow[1] = reinterpret_cast<const uint4*>(p.W + static_cast<int64_t>(n + n_out) * p.ldw);
    float acc[kRows][MT];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;

    for (int c0 = c_beg + lane; c0 < c_end; c0 += 32 * kUnroll) {
      uint4 w[kRows][kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + 32 * u;
#pragma unroll
        for (int r = 0; r < kRows; ++r) w[r][u] = (c < c_end) ? ldg_stream(wrow[r] + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + 32 * u;
        if (c < c_end) {
          float wf[kRows][8];
#pragma unroll
          for (int r = 0; r < kRows; ++r) unpack8f(w[r][u], wf[r]);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const float4 a0 = *reinterpret_cast<const float4*>(sA + m * K + c * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(sA + m * K + c * 8 + 4);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
              float s = acc[r][m];
              s = fmaf(wf[r][0], a0.x, s); s = fmaf(wf[r][1], a0.y, s); s = fmaf(wf[r][2], a0.z, s);
              s = fmaf(wf[r][3], a0.w, s); s = fmaf(wf[r][4], a1.x, s); s = fmaf(wf[r][5], a1.y, s);
              s = fmaf(wf[r][6], a1.z, s); s = fmaf(wf[r][7], a1.w, s);
              acc[r][m] = s;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r][m] += __shfl_xor_sync(0xffffffffu, acc[r][m], o);
      }
    if constexpr (KSPLIT > 1) {
      // combine the K-slices of the KSPLIT warps that share this column (double-buffered scratch, one barrier / iter)
      float* slot = red[it & 1][warp];
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
          for (int m = 0; m < MT; ++m) slot[r * MT + m] = acc[r][m];
      }
      __syncthreads();
      if (ks == 0) {
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < KSPLIT; ++j) v += red[it & 1][warp + j][r * MT + m];
            acc[r][m] = v;
          }
      }
    }
    if (ks == 0 && valid) {
      // lane m finalises row m (every lane holds the full sums)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (lane == m) {
          float o;
          if constexpr (EPI == MB_EPI_SWIGLU) {
            const float x1 = bf16_round(acc[0][m] + (p.bias ? __bfloat162float(p.bias[n]) : 0.f));
            const float x2 = bf16_round(acc[kRows - 1][m] + (p.bias ? __bfloat162float(p.bias[n + n_out]) : 0.f));
            o = bf16_round(silu(x1)) * x2;
          } else {
            float v = acc[0][m] + (p.bias ? __bfloat162float(p.bias[n]) : 0.f);
            if constexpr (EPI == MB_EPI_GELU) v = gelu_erf(bf16_round(v));
            if constexpr (EPI == MB_EPI_SILU) v = silu(bf16_round(v));
            if constexpr (EPI == MB_EPI_RESIDUAL) v = bf16_round(v) + __bfloat162float(p.res[m * p.ldr + n]);
            if constexpr (EPI == MB_EPI_GATED) {
              // x + gate * h  (diff_loss_rf_swiglu.py:272): bf16 product, bf16 sum
              const float gh = bf16_round(__bfloat162float(p.gate[m * p.ldg + n]) * bf16_round(v));
              v = __bfloat162float(p.res[m * p.ldr + n]) + gh;
            }
            o = v;
          }
          p.out[m * p.ldo + n] = __float2bfloat16_rn(o);
          if (p.out_f32) p.out_f32[m * n_out + n] = bf16_round(o);
        }
      }
    }
  }
}

template <int MT, int KSPLIT>
static int launch_gemv_ks(const GemvParams& p, int epi, int grid, size_t smem, cudaStream_t stream) {
#define MB_GEMV_CASE(E_)                                                                                  \
  case E_: {                                                                                              \
    static bool attr_set = false;                                                                         \
    if (!attr_set) {                                                                                      \
      MB_CHECK_CUDA(cudaFuncSetAttribute(gemv_bf16_kernel<MT, E_, KSPLIT>,                                \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));       \
      attr_set = true;                                                                                    \
    }                                                                                                     \
    gemv_bf16_kernel<MT, E_, KSPLIT><<<grid, kGemvThreads, smem, stream>>>(p);                            \
    break;                                                                                                \
  }
  switch (epi) {
    MB_GEMV_CASE(MB_EPI_BIAS)
    MB_GEMV_CASE(MB_EPI_GELU)
    MB_GEMV_CASE(MB_EPI_SWIGLU)
    MB_GEMV_CASE(MB_EPI_RESIDUAL)
    MB_GEMV_CASE(MB_EPI_SILU)
    MB_GEMV_CASE(MB_EPI_GATED)
    default:
      set_error("mb_gemv_bf16: unknown epilogue %d", epi);
      return MB_ERR_SHAPE;
  }
#undef MB_GEMV_CASE
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

template <int MT>
static int launch_gemv(const GemvParams& p, int epi, int ksplit, int grid, size_t smem, cudaStream_t stream) {
  switch (ksplit) {
    case 1: return launch_gemv_ks<MT, 1>(p, epi, grid, smem, stream);
    case 2: return launch_gemv_ks<MT, 2>(p, epi, grid, smem, stream);
    case 4: return launch_gemv_ks<MT, 4>(p, epi, grid, smem, stream);
    default: return launch_gemv_ks<MT, 8>(p, epi, grid, smem, stream);
  }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_gemv_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                            int64_t ldo, int M, int N, int K, int epi, const void* residual, int64_t ldr,
                            const void* gate, int64_t ldg, void* out_f32, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_gemv_bf16: no sm_100 device");
  MB_CHECK_ARG(M >= 0 && M <= 8 && N >= 1 && K >= 8, MB_ERR_SHAPE, "mb_gemv_bf16: need 0 <= M <= 8 (M=%d N=%d K=%d)", M,
               N, K);
  if (M == 0) return MB_OK;
  MB_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, MB_ERR_ALIGN, "mb_gemv_bf16: K, lda, ldw must be multiples of 8");
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, MB_ERR_ALIGN,
               "mb_gemv_bf16: A and W must be 16-byte aligned");
  if (epi == MB_EPI_SWIGLU) MB_CHECK_ARG(N % 2 == 0, MB_ERR_SHAPE, "mb_gemv_bf16: SWIGLU needs an even N");
  if (epi == MB_EPI_RESIDUAL || epi == MB_EPI_GATED)
    MB_CHECK_ARG(residual != nullptr, MB_ERR_SHAPE, "mb_gemv_bf16: residual epilogue without a residual pointer");
  if (epi == MB_EPI_GATED) MB_CHECK_ARG(gate != nullptr, MB_ERR_SHAPE, "mb_gemv_bf16: GATED epilogue without a gate");
  const size_t smem = static_cast<size_t>(M) * K * 4;
  MB_CHECK_ARG(smem <= 200 * 1024, MB_ERR_SHAPE, "mb_gemv_bf16: M*K too large for shared memory (M=%d K=%d)", M, K);

  GemvParams p;
  p.A = static_cast<const __nv_bfloat16*>(A); p.lda = lda;
  p.W = static_cast<const __nv_bfloat16*>(W); p.ldw = ldw;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  p.res = static_cast<const __nv_bfloat16*>(residual); p.ldr = ldr;
  p.gate = static_cast<const __nv_bfloat16*>(gate); p.ldg = ldg;
  p.out_f32 = static_cast<float*>(out_f32);
  p.M = M; p.N = N; p.K = K;
  const int n_out = (epi == MB_EPI_SWIGLU) ? N / 2 : N;
  // K-split: enough (column, K-slice) work items to occupy ~16 warps on every SM, but >= 2 chunks per lane and slice
  const int kchunks = K / 8;
  int ksplit = 1;
  while (ksplit < 8 && static_cast<long>(n_out) * ksplit < static_cast<long>(num_sms()) * 16 &&
         kchunks / (ksplit * 2) >= 64)
    ksplit *= 2;
  const int cols_per_iter = (kGemvThreads / 32) / ksplit;
  const int units = (n_out + cols_per_iter - 1) / cols_per_iter;
  const int per_sm = smem <= 64 * 1024 ? 3 : (smem <= 100 * 1024 ? 2 : 1);
  int grid = num_sms() * per_sm;
  if (grid > units) grid = units;
  switch (M) {
    case 1: return launch_gemv<1>(p, epi, ksplit, grid, smem, stream);
    case 2: return launch_gemv<2>(p, epi, ksplit, grid, smem, stream);
    case 3: return launch_gemv<3>(p, epi, ksplit, grid, smem, stream);
    case 4: return launch_gemv<4>(p, epi, ksplit, grid, smem, stream);
    case 5: return launch_gemv<5>(p, epi, ksplit, grid, smem, stream);
    case 6: return launch_gemv<6>(p, epi, ksplit, grid, smem, stream);
    case 7: return launch_gemv<7>(p, epi, ksplit, grid, smem, stream);
    default: return launch_gemv<8>(p, epi, ksplit, grid, smem, stream);
  }
}
