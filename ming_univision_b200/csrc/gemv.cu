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
  // Fused input normalisation (mb_gemv_bf16_norm): the activation rows are normalised while they are staged in shared
  // memory, so the separate row kernel (and its launch + dependency gap in front of EVERY streaming GEMM of the RF
  // head / the LLM step) disappears.  Each CTA redoes the M <= 8 row statistics: ~K flops per row, nothing next to
  // the weight tile it streams.
  //   pro = 1  adaLN    : bf16( (LN(a) * gamma + beta) * bf16(1 + scale[m]) + shift[m] )   diff_loss_rf_swiglu.py:184-185, 270, 290
  //   pro = 2  RMSNorm  : bf16( gamma * bf16(a * rsqrt(mean(a^2) + eps)) )                  modeling_bailing_moe.py:131-136
  int pro;
  const __nv_bfloat16* pro_gamma;
  const __nv_bfloat16* pro_beta;
  const __nv_bfloat16* pro_shift; int64_t ld_shift;
  const __nv_bfloat16* pro_scale; int64_t ld_scale;
  float pro_eps;
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// EPI: 0 bias, 1 gelu, 2 swiglu (W = [2H, K] reference layout, out has H columns), 3 residual, 4 silu, 5 gated residual
//
// The streaming kernel must spend almost no issue slots per weight byte, or it cannot keep enough loads in flight to
// reach HBM speed (a CUDA-core version needed ~110 instructions per 32 weight bytes per lane and topped out at ~40 % of
// the copy bandwidth).  So the dot products run on the tensor cores, fed DIRECTLY from coalesced 16-byte global loads:
//   * a warp owns a tile of 16 weight rows (x2 for SwiGLU: gate rows n.. and up rows n + H..); lane (g = lane / 4,
//     t = lane % 4) loads 16 bytes of row g and of row g + 8 at K offset 32 kg + 8 t  — 64 contiguous bytes per row
//     across the 4 lanes of a quad, every byte of W read exactly once;
//   * those 8 + 8 bf16 are the A fragments of two mma.m16n8k16: the MMA's k index is a PERMUTATION of the memory order
//     (lane t's elements {0,1} and {2,3} take k' = 2t, 2t+1 and 2t+8, 2t+9), which is harmless for a dot product as
//     long as the B fragment uses the same permutation — and it does, because lane (g, t) reads activation row g (a
//     token) at the same K offset 32 kg + 8 t from shared memory with one conflict-free 16-byte load;
//   * n = 8 MMA columns = up to 8 tokens (rows >= M read a shared zero row).
// Per 1 KB of weights a warp issues 2 LDG.128 + 1 LDS.128 + 2 HMMA.  KSPLIT warps of a CTA share one row tile and split
// its K range (narrow layers still fill the GPU); their partial sums meet in shared memory.
template <int EPI, int KSPLIT, bool NORM>
__global__ void __launch_bounds__(kGemvThreads, 2)
gemv_bf16_kernel(const GemvParams p) {
  constexpr int kTiles = (EPI == MB_EPI_SWIGLU) ? 2 : 1;  // 16-row tiles per work item
  constexpr int kUnroll = (EPI == MB_EPI_SWIGLU) ? 4 : 8;
  constexpr int kItemsPerIter = (kGemvThreads / 32) / KSPLIT;
  extern __shared__ __align__(16) uint8_t gemv_smem[];
  __shared__ float red[2][kGemvThreads / 32][kTiles * 4][32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K, kchunks = K >> 3;
  const int kgroups = (kchunks + 3) >> 2;               // groups of 32 k
  const int row_bytes = kgroups * 64 + 64;               // +64 B: consecutive token rows land in different bank halves
  const int n_out = (EPI == MB_EPI_SWIGLU) ? p.N / 2 : p.N;

  const int item_slot = warp / KSPLIT, ks = warp % KSPLIT;
  const int per = (kgroups + KSPLIT - 1) / KSPLIT;
  const int kg_beg = ks * per, kg_end = min(kgroups, kg_beg + per);
  const int n_items = (n_out + 15) / 16;
  const int iters = (n_items + gridDim.x * kItemsPerIter - 1) / (gridDim.x * kItemsPerIter);

  auto row_ptrs = [&](int it, const __nv_bfloat16* (&wr)[kTiles][2], int& n0, bool& item_valid) {
    const int item = (it * gridDim.x + blockIdx.x) * kItemsPerIter + item_slot;
    item_valid = item < n_items;
    n0 = (item_valid ? item : 0) * 16;
#pragma unroll
    for (int tl = 0; tl < kTiles; ++tl) {
      const int base = tl * n_out;  // the rows this lane streams (clamped at the matrix edge; never stored if clamped)
      wr[tl][0] = p.W + static_cast<int64_t>(base + min(n0 + g, n_out - 1)) * p.ldw + t * 8;
      wr[tl][1] = p.W + static_cast<int64_t>(base + min(n0 + g + 8, n_out - 1)) * p.ldw + t * 8;
    }
  };
  auto load_batch = [&](const __nv_bfloat16* (&wr)[kTiles][2], int kg0, uint4 (&w)[kTiles][2][kUnroll]) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int kg = kg0 + u;
      const bool ok = kg < kg_end && (kg * 4 + t) < kchunks;
#pragma unroll
      for (int tl = 0; tl < kTiles; ++tl) {
        w[tl][0][u] = ok ? ldg_stream(reinterpret_cast<const uint4*>(wr[tl][0] + kg * 32)) : make_uint4(0, 0, 0, 0);
        w[tl][1][u] = ok ? ldg_stream(reinterpret_cast<const uint4*>(wr[tl][1] + kg * 32)) : make_uint4(0, 0, 0, 0);
      }
    }
  };

  // PDL: let the next kernel of the stream start launching, and put this warp's FIRST batch of weight loads in flight
  // before anything else: weights do not depend on the predecessor, so their DRAM latency overlaps the predecessor's
  // tail, the wait, and the staging of the activations below
  pdl_launch_dependents();
  const __nv_bfloat16* wr0[kTiles][2];
  int n0_first;
  bool valid_first;
  row_ptrs(0, wr0, n0_first, valid_first);
  uint4 w[kTiles][2][kUnroll];
  load_batch(wr0, kg_beg, w);
  pdl_wait();

  // stage the activations: rows 0..M-1, then one zero row shared by the unused MMA columns
  if constexpr (!NORM) {
    for (int i = tid; i < (p.M + 1) * (row_bytes / 16); i += kGemvThreads) {
      const int m = i / (row_bytes / 16), c = i % (row_bytes / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (m < p.M && c < kchunks) v = *reinterpret_cast<const uint4*>(p.A + m * p.lda + c * 8);
      *reinterpret_cast<uint4*>(gemv_smem + m * row_bytes + c * 16) = v;
    }
  } else {
    // All 256 threads cooperate on every row: each thread keeps its 16-byte chunks of ALL rows in registers (one pass
    // over global memory, all loads in flight together), the row statistics are block reductions (shuffle + one shared
    // array), and the normalised rows go straight to the staging area.  ~2 us, most of it hidden behind the first
    // weight batch that is already in flight.
    constexpr int kMaxChunks = 2;   // per thread and row: K <= 2 * 256 * 8 = 4096
    constexpr int kMaxRowsReg = 2;  // rows held in registers per round (register budget: two CTAs per SM)
    __shared__ float red_s[kMaxRowsReg][kGemvThreads / 32];
    for (int m0 = 0; m0 < p.M; m0 += kMaxRowsReg) {
      const int nm = min(kMaxRowsReg, p.M - m0);
      uint4 raw[kMaxRowsReg][kMaxChunks];  // the rows stay packed (bf16) in registers and are unpacked per pass
      float part[kMaxRowsReg];
      auto unpack8 = [](const uint4& q, float (&f)[8]) {
        const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
        f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; f[4] = f2.x; f[5] = f2.y; f[6] = f3.x; f[7] = f3.y;
      };
#pragma unroll
      for (int r = 0; r < kMaxRowsReg; ++r) {
        part[r] = 0.f;
#pragma unroll
        for (int cc = 0; cc < kMaxChunks; ++cc) {
          const int c = tid + cc * kGemvThreads;
          const bool ok = r < nm && c < kchunks;
          raw[r][cc] = ok ? *reinterpret_cast<const uint4*>(p.A + (m0 + r) * p.lda + c * 8) : make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int r = 0; r < kMaxRowsReg; ++r) {
#pragma unroll
        for (int cc = 0; cc < kMaxChunks; ++cc) {
          float f[8];
          unpack8(raw[r][cc], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) part[r] += (p.pro == 1) ? f[i] : f[i] * f[i];
        }
      }
      auto block_sums = [&](float (&x)[kMaxRowsReg]) {  // x[r] <- sum over the CTA
#pragma unroll
        for (int r = 0; r < kMaxRowsReg; ++r) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) x[r] += __shfl_xor_sync(0xffffffffu, x[r], o);
        }
        __syncthreads();  // red_s free (previous use fully read)
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < kMaxRowsReg; ++r) red_s[r][warp] = x[r];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kMaxRowsReg; ++r) {
          float t2 = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < kGemvThreads / 32; ++w2) t2 += red_s[r][w2];
          x[r] = t2;
        }
      };
      block_sums(part);
      float mean[kMaxRowsReg], rstd[kMaxRowsReg];
      if (p.pro == 1) {
        float sq[kMaxRowsReg];
#pragma unroll
        for (int r = 0; r < kMaxRowsReg; ++r) {
          mean[r] = part[r] / K;
          sq[r] = 0.f;
#pragma unroll
          for (int cc = 0; cc < kMaxChunks; ++cc) {
            if (tid + cc * kGemvThreads < kchunks) {
              float f[8];
              unpack8(raw[r][cc], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float d = f[i] - mean[r];
                sq[r] += d * d;
              }
            }
          }
        }
        block_sums(sq);
#pragma unroll
        for (int r = 0; r < kMaxRowsReg; ++r) rstd[r] = rsqrtf(sq[r] / K + p.pro_eps);
      } else {
#pragma unroll
        for (int r = 0; r < kMaxRowsReg; ++r) {
          mean[r] = 0.f;
          rstd[r] = rsqrtf(part[r] / K + p.pro_eps);
        }
      }
#pragma unroll
      for (int cc = 0; cc < kMaxChunks; ++cc) {
        const int c = tid + cc * kGemvThreads;
        if (c >= kchunks) continue;
        __nv_bfloat16 gm[8], bt[8];
        if (p.pro_gamma) *reinterpret_cast<uint4*>(gm) = *reinterpret_cast<const uint4*>(p.pro_gamma + c * 8);
        if (p.pro_beta) *reinterpret_cast<uint4*>(bt) = *reinterpret_cast<const uint4*>(p.pro_beta + c * 8);
#pragma unroll
        for (int r = 0; r < kMaxRowsReg; ++r) {
          if (r >= nm) continue;
          const int m = m0 + r;
          __nv_bfloat16 sh[8], sc[8];
          if (p.pro == 1) {
            *reinterpret_cast<uint4*>(sh) = *reinterpret_cast<const uint4*>(p.pro_shift + m * p.ld_shift + c * 8);
            *reinterpret_cast<uint4*>(sc) = *reinterpret_cast<const uint4*>(p.pro_scale + m * p.ld_scale + c * 8);
          }
          float o[8], f[8];
          unpack8(raw[r][cc], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (p.pro == 1) {
              float x = (f[i] - mean[r]) * rstd[r];
              if (p.pro_gamma) x = x * __bfloat162float(gm[i]) + (p.pro_beta ? __bfloat162float(bt[i]) : 0.f);
              o[i] = x * bf16_round(1.f + __bfloat162float(sc[i])) + __bfloat162float(sh[i]);
            } else {
              o[i] = __bfloat162float(gm[i]) * (f[i] * rstd[r]);  // one rounding, as BailingMoeRMSNorm (:131-136)
            }
          }
          uint4 o4;
          o4.x = pack_bf16x2(o[0], o[1]); o4.y = pack_bf16x2(o[2], o[3]);
          o4.z = pack_bf16x2(o[4], o[5]); o4.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(gemv_smem + m * row_bytes + c * 16) = o4;
        }
      }
    }
    // padding chunks of the data rows and the shared zero row
    for (int i = tid; i < (p.M + 1) * (row_bytes / 16); i += kGemvThreads) {
      const int m = i / (row_bytes / 16), c = i % (row_bytes / 16);
      if (m == p.M || c >= kchunks) *reinterpret_cast<uint4*>(gemv_smem + m * row_bytes + c * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  __syncthreads();
  const uint8_t* xrow = gemv_smem + min(g, p.M) * row_bytes + t * 16;

  for (int it = 0; it < iters; ++it) {
    const __nv_bfloat16* wr[kTiles][2];
    int n0;
    bool item_valid;
    row_ptrs(it, wr, n0, item_valid);
    float acc[kTiles][4];
#pragma unroll
    for (int tl = 0; tl < kTiles; ++tl) acc[tl][0] = acc[tl][1] = acc[tl][2] = acc[tl][3] = 0.f;

    for (int kg0 = kg_beg; kg0 < kg_end; kg0 += kUnroll) {
      if (!(it == 0 && kg0 == kg_beg)) load_batch(wr, kg0, w);  // the very first batch is already in flight
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int kg = kg0 + u;
        if (kg < kg_end) {
          const uint4 xb = *reinterpret_cast<const uint4*>(xrow + kg * 64);
#pragma unroll
          for (int tl = 0; tl < kTiles; ++tl) {
            mma16816(acc[tl], w[tl][0][u].x, w[tl][1][u].x, w[tl][0][u].y, w[tl][1][u].y, xb.x, xb.y);
            mma16816(acc[tl], w[tl][0][u].z, w[tl][1][u].z, w[tl][0][u].w, w[tl][1][u].w, xb.z, xb.w);
          }
        }
      }
    }
    if constexpr (KSPLIT > 1) {
      // combine the K-slices of the KSPLIT warps that share this item (double-buffered scratch, one barrier / iter)
#pragma unroll
      for (int tl = 0; tl < kTiles; ++tl)
#pragma unroll
        for (int e = 0; e < 4; ++e) red[it & 1][warp][tl * 4 + e][lane] = acc[tl][e];
      __syncthreads();
      if (ks == 0) {
#pragma unroll
        for (int tl = 0; tl < kTiles; ++tl)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < KSPLIT; ++j) v += red[it & 1][warp + j][tl * 4 + e][lane];
            acc[tl][e] = v;
          }
      }
    }
    if (ks == 0 && item_valid) {
      // accumulator element e: row n0 + g + 8 (e >> 1), token 2 t + (e & 1)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + g + 8 * (e >> 1);
        const int m = 2 * t + (e & 1);
        if (n < n_out && m < p.M) {
          float o;
          if constexpr (EPI == MB_EPI_SWIGLU) {
            const float x1 = bf16_round(acc[0][e] + (p.bias ? __bfloat162float(p.bias[n]) : 0.f));
            const float x2 = bf16_round(acc[kTiles - 1][e] + (p.bias ? __bfloat162float(p.bias[n + n_out]) : 0.f));
            o = bf16_round(silu(x1)) * x2;
          } else {
            float v = acc[0][e] + (p.bias ? __bfloat162float(p.bias[n]) : 0.f);
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

template <int KSPLIT, bool NORM>
static int launch_gemv_ks(const GemvParams& p, int epi, int grid, size_t smem, cudaStream_t stream) {
#define MB_GEMV_CASE(E_)                                                                                       \
  case E_: {                                                                                                   \
    static bool attr_set = false;                                                                              \
    if (!attr_set) {                                                                                           \
      MB_CHECK_CUDA(cudaFuncSetAttribute(gemv_bf16_kernel<E_, KSPLIT, NORM>,                                   \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));            \
      attr_set = true;                                                                                         \
    }                                                                                                          \
    MB_CHECK_CUDA(launch_pdl(gemv_bf16_kernel<E_, KSPLIT, NORM>, dim3(grid), dim3(kGemvThreads), smem, stream, p)); \
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

static int gemv_impl(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                     int64_t ldo, int M, int N, int K, int epi, const void* residual, int64_t ldr,
                     const void* gate, int64_t ldg, void* out_f32, int pro, const void* gamma, const void* beta,
                     const void* shift, int64_t ld_shift, const void* scale, int64_t ld_scale, float eps,
                     void* stream_) {
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
  const int kgroups_h = (K / 8 + 3) / 4;
  const size_t smem = static_cast<size_t>(M + 1) * (kgroups_h * 64 + 64);
  MB_CHECK_ARG(smem <= 160 * 1024, MB_ERR_SHAPE, "mb_gemv_bf16: M*K too large for shared memory (M=%d K=%d)", M, K);

  GemvParams p;
  p.A = static_cast<const __nv_bfloat16*>(A); p.lda = lda;
  p.W = static_cast<const __nv_bfloat16*>(W); p.ldw = ldw;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  p.res = static_cast<const __nv_bfloat16*>(residual); p.ldr = ldr;
  p.gate = static_cast<const __nv_bfloat16*>(gate); p.ldg = ldg;
  p.out_f32 = static_cast<float*>(out_f32);
  p.M = M; p.N = N; p.K = K;
  p.pro = pro;
  p.pro_gamma = static_cast<const __nv_bfloat16*>(gamma);
  p.pro_beta = static_cast<const __nv_bfloat16*>(beta);
  p.pro_shift = static_cast<const __nv_bfloat16*>(shift); p.ld_shift = ld_shift;
  p.pro_scale = static_cast<const __nv_bfloat16*>(scale); p.ld_scale = ld_scale;
  p.pro_eps = eps;
  const int n_out = (epi == MB_EPI_SWIGLU) ? N / 2 : N;
  // K-split: enough (row tile, K-slice) work items to give every SM ~8 busy warps, but >= 8 k-groups per slice
  const int n_items = (n_out + 15) / 16;
  int ksplit = 1;
  while (ksplit < 8 && static_cast<long>(n_items) * ksplit < static_cast<long>(num_sms()) * 8 &&
         kgroups_h / (ksplit * 2) >= 8)
    ksplit *= 2;
  const int items_per_iter = (kGemvThreads / 32) / ksplit;
  const int units = (n_items + items_per_iter - 1) / items_per_iter;
  // (a software-pipelined variant with two weight batches in flight per warp needs ~170 registers, i.e. one CTA per
  // SM, and measured SLOWER: 8.25 -> 10.97 ms per RF sample — occupancy hides the HBM latency better than depth)
  const int per_sm = smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1);
  int grid = num_sms() * per_sm;
  if (grid > units) grid = units;
  if (pro != 0) {  // fused-normalisation instances (only the epilogues / K-splits the RF head and the LLM step use)
    MB_CHECK_ARG(epi == MB_EPI_BIAS || epi == MB_EPI_SWIGLU, MB_ERR_SHAPE,
                 "mb_gemv_bf16_norm: BIAS and SWIGLU epilogues only");

    switch (ksplit) {
      case 1: return launch_gemv_ks<1, true>(p, epi, grid, smem, stream);
      case 2: return launch_gemv_ks<2, true>(p, epi, grid, smem, stream);
      case 4: return launch_gemv_ks<4, true>(p, epi, grid, smem, stream);
      default: return launch_gemv_ks<8, true>(p, epi, grid, smem, stream);
    }
  }
  switch (ksplit) {
    case 1: return launch_gemv_ks<1, false>(p, epi, grid, smem, stream);
    case 2: return launch_gemv_ks<2, false>(p, epi, grid, smem, stream);
    case 4: return launch_gemv_ks<4, false>(p, epi, grid, smem, stream);
    default: return launch_gemv_ks<8, false>(p, epi, grid, smem, stream);
  }
}

extern "C" int mb_gemv_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                            int64_t ldo, int M, int N, int K, int epi, const void* residual, int64_t ldr,
                            const void* gate, int64_t ldg, void* out_f32, void* stream_) {
  return gemv_impl(A, lda, W, ldw, bias, out, ldo, M, N, K, epi, residual, ldr, gate, ldg, out_f32, 0, nullptr, nullptr,
                   nullptr, 0, nullptr, 0, 0.f, stream_);
}

extern "C" int mb_gemv_bf16_norm(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                                 int64_t ldo, int M, int N, int K, int epi, const void* residual, int64_t ldr,
                                 const void* gate, int64_t ldg, void* out_f32, int norm, const void* gamma,
                                 const void* beta, const void* shift, int64_t ld_shift, const void* scale,
                                 int64_t ld_scale, float eps, void* stream_) {
  MB_CHECK_ARG(norm == 1 || norm == 2, MB_ERR_SHAPE, "mb_gemv_bf16_norm: norm must be 1 (adaLN) or 2 (RMSNorm)");
  MB_CHECK_ARG(K <= 4096, MB_ERR_SHAPE, "mb_gemv_bf16_norm: K <= 4096 (the rows are held in registers; K=%d)", K);
  if (norm == 1)
    MB_CHECK_ARG(shift != nullptr && scale != nullptr && ld_shift % 8 == 0 && ld_scale % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(shift) | reinterpret_cast<uintptr_t>(scale) |
                       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                 MB_ERR_ALIGN, "mb_gemv_bf16_norm: adaLN needs 16-byte aligned shift / scale rows (ld %% 8 == 0)");
  if (norm == 2)
    MB_CHECK_ARG(gamma != nullptr && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0, MB_ERR_ALIGN,
                 "mb_gemv_bf16_norm: RMSNorm needs a 16-byte aligned weight");
  return gemv_impl(A, lda, W, ldw, bias, out, ldo, M, N, K, epi, residual, ldr, gate, ldg, out_f32, norm, gamma, beta,
                   shift, ld_shift, scale, ld_scale, eps, stream_);
}
