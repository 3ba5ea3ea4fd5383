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

__device__ __forceinline__ float dot8(const uint4& w, const uint4& a) {
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  float s = w0.x * a0.x;
  s = fmaf(w0.y, a0.y, s); s = fmaf(w1.x, a1.x, s); s = fmaf(w1.y, a1.y, s);
  s = fmaf(w2.x, a2.x, s); s = fmaf(w2.y, a2.y, s); s = fmaf(w3.x, a3.x, s); s = fmaf(w3.y, a3.y, s);
  return s;
}

// EPI: 0 bias, 1 gelu, 2 swiglu (W = [2H, K] reference layout, out has H columns), 3 residual, 4 silu, 5 gated residual
//
// One WARP owns one output column at a time (two weight rows for SwiGLU: gate row n and up row n + H); the 32 lanes
// split K in 16-byte chunks and keep kUnroll independent loads per weight row in flight.  Warps never synchronise with
// each other after the activation matrix has been staged, so ~24 resident warps per SM x 8 outstanding 16-byte loads
// per lane keep ~100 KB per SM in flight — what it takes to cover HBM latency at 6.5 TB/s.
template <int MT, int EPI>
__global__ void __launch_bounds__(kGemvThreads)
gemv_bf16_kernel(const GemvParams p) {
  constexpr int kRows = (EPI == MB_EPI_SWIGLU) ? 2 : 1;  // weight rows per output column
  constexpr int kUnroll = (EPI == MB_EPI_SWIGLU) ? 4 : 8;
  extern __shared__ __align__(16) uint8_t gemv_smem[];
  const uint4* sA = reinterpret_cast<const uint4*>(gemv_smem);  // [MT][K / 8] chunks of 8 bf16

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, kchunks = K >> 3;
  const int n_out = (EPI == MB_EPI_SWIGLU) ? p.N / 2 : p.N;

  // stage A (bf16) in shared memory, zero-filling rows >= M
  for (int i = tid; i < MT * kchunks; i += kGemvThreads) {
    const int m = i / kchunks, c = i % kchunks;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (m < p.M) v = *reinterpret_cast<const uint4*>(p.A + m * p.lda + c * 8);
    reinterpret_cast<uint4*>(gemv_smem)[i] = v;
  }
  __syncthreads();

  const int warps_total = gridDim.x * (kGemvThreads / 32);
  for (int n = blockIdx.x * (kGemvThreads / 32) + warp; n < n_out; n += warps_total) {
    const uint4* wrow[kRows];
    wrow[0] = reinterpret_cast<const uint4*>(p.W + static_cast<int64_t>(n) * p.ldw);
    if constexpr (kRows == 2) wrow[1] = reinterpret_cast<const uint4*>(p.W + static_cast<int64_t>(n + n_out) * p.ldw);
    float acc[kRows][MT];
#pragma unroll
    for (int r = 0; r < kRows; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;

    for (int c0 = lane; c0 < kchunks; c0 += 32 * kUnroll) {
      uint4 w[kRows][kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + 32 * u;
#pragma unroll
        for (int r = 0; r < kRows; ++r) w[r][u] = (c < kchunks) ? ldg_stream(wrow[r] + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + 32 * u;
        if (c < kchunks) {
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const uint4 a = sA[m * kchunks + c];
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r][m] += dot8(w[r][u], a);
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
    // lane m finalises row m (all lanes hold the full sums after the xor-reduction)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (lane == m && m < p.M) {
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

template <int MT>
static int launch_gemv(const GemvParams& p, int epi, int grid, size_t smem, cudaStream_t stream) {
#define MB_GEMV_CASE(E_)                                                                                       \
  case E_: {                                                                                                   \
    static bool attr_set = false;                                                                              \
    if (!attr_set) {                                                                                           \
      MB_CHECK_CUDA(cudaFuncSetAttribute(gemv_bf16_kernel<MT, E_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         200 * 1024));                                                         \
      attr_set = true;                                                                                         \
    }                                                                                                          \
    gemv_bf16_kernel<MT, E_><<<grid, kGemvThreads, smem, stream>>>(p);                                         \
    break;                                                                                                     \
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
  const int mt = M <= 1 ? 1 : M <= 2 ? 2 : M <= 4 ? 4 : 8;
  const size_t smem = static_cast<size_t>(mt) * K * 2;
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
  const int units = (n_out + (kGemvThreads / 32) - 1) / (kGemvThreads / 32);  // CTAs needed for one column per warp
  const int per_sm = smem <= 64 * 1024 ? 3 : (smem <= 100 * 1024 ? 2 : 1);
  int grid = num_sms() * per_sm;
  if (grid > units) grid = units;
  switch (mt) {
    case 1: return launch_gemv<1>(p, epi, grid, smem, stream);
    case 2: return launch_gemv<2>(p, epi, grid, smem, stream);
    case 4: return launch_gemv<4>(p, epi, grid, smem, stream);
    default: return launch_gemv<8>(p, epi, grid, smem, stream);
  }
}
