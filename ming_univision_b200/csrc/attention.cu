// Fused softmax(QK^T * scale)V forward for head_dim 64 / 128 (bf16 in, fp32 softmax + accumulate, bf16 out).
// One CTA = 64 query rows of one (batch, head); 4 warps x 16 rows; K/V streamed through a double-buffered,
// XOR-swizzled shared-memory ring with cp.async; online softmax in registers.
// Warp-level mma.sync kernel: backend 2 of mb_attn_hd64 / mb_attn_fwd (A/B tests, shapes the tcgen05 kernel declines);
// the default backend is the tcgen05 / TMEM / TMA kernel in attention_tc.cu.  This file also holds the cached q_len = 1
// decode step of the semantic decoder.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

struct AttnParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  // strides in elements
  int64_t q_bs, q_ts, q_hs;
  int64_t k_bs, k_ts, k_hs;
  int64_t v_bs, v_ts, v_hs;
  int64_t o_bs, o_ts, o_hs;
  int Sq, Sk, Hq, Hkv;
  float scale_log2;  // scale * log2(e)
  int causal;        // bottom-right aligned: query i sees keys j <= i + (Sk - Sq)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Tile [rows][HD] bf16 in shared memory, 16-byte chunks XOR-swizzled by (row & 7) within each 128-byte segment.
template <int HD>
__device__ __forceinline__ __nv_bfloat16* tile_ptr(__nv_bfloat16* base, int row, int chunk) {
  // chunk = 16-byte chunk index along the row (HD/8 chunks per row)
  const int seg = chunk >> 3, c = chunk & 7;
  return base + row * HD + seg * 64 + ((c ^ (row & 7)) << 3);
}

template <int HD, int ROWS>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* smem_tile, const __nv_bfloat16* gbase, int64_t tok_stride,
                                                int row0, int nrows_total, int tid) {
  constexpr int kChunksPerRow = HD / 8;
  constexpr int kChunks = ROWS * kChunksPerRow;
#pragma unroll
  for (int i = 0; i < kChunks / 128; ++i) {
    const int idx = tid + i * 128;
    const int r = idx / kChunksPerRow, c = idx % kChunksPerRow;
    const int grow = row0 + r;
    const bool ok = grow < nrows_total;
    const __nv_bfloat16* src = gbase + static_cast<int64_t>(ok ? grow : 0) * tok_stride + c * 8;
    cp_async16(tile_ptr<HD>(smem_tile, r, c), src, ok);
  }
}

// MTW = m16 tiles per warp: 1 -> 64 query rows per CTA, 2 -> 128.  With two tiles every K / V fragment fetched with
// ldmatrix feeds twice as many MMAs (32 FLOP per shared-memory byte instead of 16), which is what lifts the kernel
// off the shared-memory bandwidth limit on the long-sequence (pixel-decoder) shapes.
template <int HD, int MTW>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const AttnParams p) {
  constexpr int kBM = 64 * MTW, kBN = 64;
  constexpr int kKSteps = HD / 16;   // k-steps of QK^T
  constexpr int kDTiles = HD / 8;    // n-tiles of the output
  extern __shared__ __align__(128) uint8_t attn_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(attn_smem);
  __nv_bfloat16* sK = sQ + kBM * HD;            // 2 buffers
  __nv_bfloat16* sV = sK + 2 * kBN * HD;        // 2 buffers

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hkv = h / (p.Hq / p.Hkv);
  const int q0 = qb * kBM;
  const int wrow0 = warp * 16 * MTW;  // first row of this warp inside the CTA tile
  const int shift = p.Sk - p.Sq;      // causal offset

  const __nv_bfloat16* gq = p.q + b * p.q_bs + h * p.q_hs;
  const __nv_bfloat16* gk = p.k + b * p.k_bs + hkv * p.k_hs;
  const __nv_bfloat16* gv = p.v + b * p.v_bs + hkv * p.v_hs;

  int n_blocks = (p.Sk + kBN - 1) / kBN;
  if (p.causal) {
    const int last_key = min(p.Sk - 1, q0 + kBM - 1 + shift);
    n_blocks = min(n_blocks, last_key / kBN + 1);
    if (last_key < 0) n_blocks = 0;
  }
  pdl_launch_dependents();
  pdl_wait();

  load_tile_async<HD, kBM>(sQ, gq, p.q_ts, q0, p.Sq, tid);
  if (n_blocks > 0) {
    load_tile_async<HD, kBN>(sK, gk, p.k_ts, 0, p.Sk, tid);
    load_tile_async<HD, kBN>(sV, gv, p.v_ts, 0, p.Sk, tid);
  }
  cp_async_commit();

  uint32_t qf[MTW][kKSteps][4];
  float o[MTW][kDTiles][4];
  float m_run[MTW][2], l_run[MTW][2];
#pragma unroll
  for (int mt = 0; mt < MTW; ++mt) {
#pragma unroll
    for (int i = 0; i < kDTiles; ++i) o[mt][i][0] = o[mt][i][1] = o[mt][i][2] = o[mt][i][3] = 0.f;
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
  }

  for (int nb = 0; nb < n_blocks; ++nb) {
    const int buf = nb & 1;
    if (nb + 1 < n_blocks) {
      load_tile_async<HD, kBN>(sK + (buf ^ 1) * kBN * HD, gk, p.k_ts, (nb + 1) * kBN, p.Sk, tid);
      load_tile_async<HD, kBN>(sV + (buf ^ 1) * kBN * HD, gv, p.v_ts, (nb + 1) * kBN, p.Sk, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (nb == 0) {
      // Q fragments: rows wrow0 + 16 mt + (lane & 15), chunk = 2*ks + (lane >> 4)
#pragma unroll
      for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks)
          ldmatrix_x4(qf[mt][ks], tile_ptr<HD>(sQ, wrow0 + 16 * mt + (lane & 15), 2 * ks + (lane >> 4)));
    }
    __nv_bfloat16* tK = sK + buf * kBN * HD;
    __nv_bfloat16* tV = sV + buf * kBN * HD;

    // ---- S = Q K^T (MTW x 16 x 64 per warp); each K fragment is used by all MTW row tiles
    float s[MTW][8][4];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[mt][j][0] = s[mt][j][1] = s[mt][j][2] = s[mt][j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kKSteps; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        // 4 matrices: keys (16jp .. 16jp+7 | +8..15) x d-chunks (2ks | 2ks+1)
        uint32_t kf[4];
        const int krow = jp * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int kchunk = 2 * ks + ((lane >> 3) & 1);
        ldmatrix_x4(kf, tile_ptr<HD>(tK, krow, kchunk));
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt) {
          mma_bf16_16816(s[mt][2 * jp], qf[mt][ks], kf[0], kf[1]);
          mma_bf16_16816(s[mt][2 * jp + 1], qf[mt][ks], kf[2], kf[3]);
        }
      }
    }

    // ---- mask + online softmax.  Scores stay UNSCALED; the softmax scale is folded into one FMA per element:
    // p = ex2(s * scale_log2 - m * scale_log2).
    const int key0 = nb * kBN;
    const bool full_tile = (key0 + kBN <= p.Sk) && (!p.causal || key0 + kBN - 1 <= q0 + wrow0 + shift);
    uint32_t pa[MTW][kBN / 16][4];
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt) {
      const int qrow0 = q0 + wrow0 + 16 * mt + g;
      if (!full_tile) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = key0 + j * 8 + 2 * t + (e & 1);
            const int qrow = qrow0 + ((e >> 1) << 3);
            bool ok = key < p.Sk;
            if (p.causal) ok = ok && (key <= qrow + shift);
            if (!ok) s[mt][j][e] = -INFINITY;
          }
        }
      }
      float m_new[2] = {m_run[mt][0], m_run[mt][1]};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m_new[0] = fmaxf(m_new[0], fmaxf(s[mt][j][0], s[mt][j][1]));
        m_new[1] = fmaxf(m_new[1], fmaxf(s[mt][j][2], s[mt][j][3]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
        m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
      }
      float corr[2], mneg[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float msafe = (m_new[r] == -INFINITY) ? 0.f : m_new[r];
        corr[r] = fast_ex2((m_run[mt][r] - msafe) * p.scale_log2);  // m_run = -inf -> 0
        mneg[r] = -msafe * p.scale_log2;
        m_run[mt][r] = m_new[r];
        l_run[mt][r] *= corr[r];
      }
      float rowsum[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = fast_ex2(fmaf(s[mt][j][e], p.scale_log2, mneg[e >> 1]));
          s[mt][j][e] = pv;
          rowsum[e >> 1] += pv;
        }
      }
      l_run[mt][0] += rowsum[0];
      l_run[mt][1] += rowsum[1];
#pragma unroll
      for (int i = 0; i < kDTiles; ++i) {
        o[mt][i][0] *= corr[0]; o[mt][i][1] *= corr[0];
        o[mt][i][2] *= corr[1]; o[mt][i][3] *= corr[1];
      }
      // P in registers: the C-fragments of key tiles 2ks, 2ks+1 form the A-fragment of k-step ks of the PV product
#pragma unroll
      for (int ks = 0; ks < kBN / 16; ++ks) {
        pa[mt][ks][0] = pack_bf16x2(s[mt][2 * ks][0], s[mt][2 * ks][1]);
        pa[mt][ks][1] = pack_bf16x2(s[mt][2 * ks][2], s[mt][2 * ks][3]);
        pa[mt][ks][2] = pack_bf16x2(s[mt][2 * ks + 1][0], s[mt][2 * ks + 1][1]);
        pa[mt][ks][3] = pack_bf16x2(s[mt][2 * ks + 1][2], s[mt][2 * ks + 1][3]);
      }
    }

    // ---- O += P V; each V^T fragment (ldmatrix.trans) is used by all MTW row tiles
#pragma unroll
    for (int ks = 0; ks < kBN / 16; ++ks) {
#pragma unroll
      for (int dp = 0; dp < kDTiles / 2; ++dp) {
        uint32_t vf[4];
        const int vrow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int vchunk = 2 * dp + (lane >> 4);
        ldmatrix_x4_trans(vf, tile_ptr<HD>(tV, vrow, vchunk));
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt) {
          mma_bf16_16816(o[mt][2 * dp], pa[mt][ks], vf[0], vf[1]);
          mma_bf16_16816(o[mt][2 * dp + 1], pa[mt][ks], vf[2], vf[3]);
        }
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  if (n_blocks == 0) {
    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- finalise: O / l, stage through this warp's own Q rows, 16-byte coalesced stores
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < MTW; ++mt) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 1);
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 2);
    }
    const float inv0 = l_run[mt][0] > 0.f ? 1.f / l_run[mt][0] : 0.f;
    const float inv1 = l_run[mt][1] > 0.f ? 1.f / l_run[mt][1] : 0.f;
#pragma unroll
    for (int i = 0; i < kDTiles; ++i) {
      // element (row g, cols 8i + 2t, +1) and (row g+8, ...)
      __nv_bfloat16* p0 = tile_ptr<HD>(sQ, wrow0 + 16 * mt + g, i) + 2 * t;
      __nv_bfloat16* p1 = tile_ptr<HD>(sQ, wrow0 + 16 * mt + g + 8, i) + 2 * t;
      *reinterpret_cast<uint32_t*>(p0) = pack_bf16x2(o[mt][i][0] * inv0, o[mt][i][1] * inv0);
      *reinterpret_cast<uint32_t*>(p1) = pack_bf16x2(o[mt][i][2] * inv1, o[mt][i][3] * inv1);
    }
  }
  __syncwarp();
  __nv_bfloat16* go = p.o + b * p.o_bs + h * p.o_hs;
  constexpr int kChunksPerRow = HD / 8;
#pragma unroll
  for (int i = 0; i < 16 * MTW * kChunksPerRow / 32; ++i) {
    const int idx = lane + i * 32;
    const int r = idx / kChunksPerRow, c = idx % kChunksPerRow;
    const int qrow = q0 + wrow0 + r;
    if (qrow < p.Sq)
      *reinterpret_cast<uint4*>(go + static_cast<int64_t>(qrow) * p.o_ts + c * 8) =
          *reinterpret_cast<const uint4*>(tile_ptr<HD>(sQ, wrow0 + r, c));
  }
}

template <int HD, int MTW>
static int launch_attn_mt(const AttnParams& p, int B, cudaStream_t stream) {
  constexpr int smem = (64 * MTW + 4 * 64) * HD * 2;
  static bool attr_set = false;
  if (!attr_set) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD, MTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((p.Sq + 64 * MTW - 1) / (64 * MTW), p.Hq, B);
  MB_CHECK_CUDA(launch_pdl(attn_fwd_kernel<HD, MTW>, grid, dim3(128), smem, stream, p));
  return MB_OK;
}

template <int HD>
static int launch_attn(const AttnParams& p, int B, cudaStream_t stream) {
  // 128-row query tiles for long sequences (measured: +4 % at S = 256, +10 % at S = 1024; slower at S = 65 where half
  // of a 128-row tile is padding).  head_dim 64 only: two row tiles at head_dim 128 would need > 255 registers.
  if (HD == 64 && p.Sq >= 192) return launch_attn_mt<64, 2>(p, B, stream);
  return launch_attn_mt<HD, 1>(p, B, stream);
}

// ------------------------------------------------------------------------------------------------------------
// Decode step (q_len = 1) against a static KV cache, head_dim 64: one warp per (batch, head).
// Appends the new K/V at position t, then attends to 0..t.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_hd64_decode_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ kcache,
                        __nv_bfloat16* __restrict__ vcache, __nv_bfloat16* __restrict__ out, int B, int H, int t_host,
                        int Tmax, float scale, const int32_t* __restrict__ t_dev) {
  const int t = (t_dev ? *t_dev : 0) + t_host;
  const int warp_global = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp_global >= B * H) return;
  const int b = warp_global / H, h = warp_global % H;
  const __nv_bfloat16* q = qkv + (static_cast<int64_t>(b) * 3 * H + h) * 64;
  const __nv_bfloat16* kn = q + static_cast<int64_t>(H) * 64;
  const __nv_bfloat16* vn = kn + static_cast<int64_t>(H) * 64;
  __nv_bfloat16* kc = kcache + (static_cast<int64_t>(b) * H + h) * Tmax * 64;
  __nv_bfloat16* vc = vcache + (static_cast<int64_t>(b) * H + h) * Tmax * 64;
  // each lane owns dims 2*lane, 2*lane+1
  const float2 qv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(q + 2 * lane));
  const uint32_t knew = *reinterpret_cast<const uint32_t*>(kn + 2 * lane);
  const uint32_t vnew = *reinterpret_cast<const uint32_t*>(vn + 2 * lane);
  *reinterpret_cast<uint32_t*>(kc + static_cast<int64_t>(t) * 64 + 2 * lane) = knew;
  *reinterpret_cast<uint32_t*>(vc + static_cast<int64_t>(t) * 64 + 2 * lane) = vnew;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int j = 0; j <= t; ++j) {
    const uint32_t kk = (j == t) ? knew : *reinterpret_cast<const uint32_t*>(kc + static_cast<int64_t>(j) * 64 + 2 * lane);
    const uint32_t vv = (j == t) ? vnew : *reinterpret_cast<const uint32_t*>(vc + static_cast<int64_t>(j) * 64 + 2 * lane);
    const float2 kf = unpack_bf16x2(kk);
    float s = qv.x * kf.x + qv.y * kf.y;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    s *= scale;
    const float m_new = fmaxf(m, s);
    const float corr = __expf(m - m_new);
    const float pj = __expf(s - m_new);
    const float2 vf = unpack_bf16x2(vv);
    l = l * corr + pj;
    o0 = o0 * corr + pj * vf.x;
    o1 = o1 * corr + pj * vf.y;
    m = m_new;
  }
  const float inv = 1.f / l;
  *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * H + h) * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
}

// attention_tc.cu: tcgen05 / TMEM kernel.  Returns 1 if it launched, 0 if the problem is not eligible, < 0 on error.
int launch_attn_tc(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, int64_t k_bs, int64_t k_ts,
                   int64_t k_hs, const void* v, int64_t v_bs, int64_t v_ts, int64_t v_hs, void* out, int64_t o_bs,
                   int64_t o_ts, int64_t o_hs, int B, int Sq, int Sk, int Hq, int Hkv, int hd, float scale, int causal,
                   cudaStream_t stream);

}  // namespace mb

using namespace mb;

extern "C" int mb_attn_hd64(const void* qkv, void* out, int B, int S, int H, float scale, int causal, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_attn_hd64: no sm_100 device");
  MB_CHECK_ARG(B >= 0 && S >= 0 && H >= 1 && H <= 65535 && B <= 65535, MB_ERR_SHAPE, "mb_attn_hd64: bad shape");
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               MB_ERR_ALIGN, "mb_attn_hd64: qkv/out must be 16-byte aligned");
  if (B == 0 || S == 0) return MB_OK;
  AttnParams p;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  const int64_t ts = static_cast<int64_t>(3) * H * 64;
  p.q = base; p.k = base + static_cast<int64_t>(H) * 64; p.v = base + static_cast<int64_t>(2) * H * 64;
  p.o = static_cast<__nv_bfloat16*>(out);
  p.q_bs = p.k_bs = p.v_bs = ts * S;
  p.q_ts = p.k_ts = p.v_ts = ts;
  p.q_hs = p.k_hs = p.v_hs = 64;
  p.o_bs = static_cast<int64_t>(S) * H * 64; p.o_ts = static_cast<int64_t>(H) * 64; p.o_hs = 64;
  p.Sq = S; p.Sk = S; p.Hq = H; p.Hkv = H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  {
    const int rc = launch_attn_tc(p.q, p.q_bs, p.q_ts, p.q_hs, p.k, p.k_bs, p.k_ts, p.k_hs, p.v, p.v_bs, p.v_ts, p.v_hs,
                                  p.o, p.o_bs, p.o_ts, p.o_hs, B, S, S, H, H, 64, scale, causal, stream);
    if (rc != 0) return rc < 0 ? rc : MB_OK;
  }
  return launch_attn<64>(p, B, stream);
}

extern "C" int mb_attn_hd64_decode(const void* qkv, void* kcache, void* vcache, void* out, int B, int H, int t,
                                   int Tmax, float scale, const int32_t* t_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_attn_hd64_decode: no sm_100 device");
  MB_CHECK_ARG(B >= 0 && H >= 1 && t >= 0 && (t_dev != nullptr || t < Tmax), MB_ERR_SHAPE,
               "mb_attn_hd64_decode: position t=%d outside the cache (Tmax=%d)", t, Tmax);
  if (B == 0) return MB_OK;
  const int warps = B * H;
  attn_hd64_decode_kernel<<<(warps + 3) / 4, 128, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(kcache),
      static_cast<__nv_bfloat16*>(vcache), static_cast<__nv_bfloat16*>(out), B, H, t, Tmax, scale, t_dev);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

// General strided entry: q[B, Sq, Hq, hd], k/v[B, Sk, Hkv, hd] with explicit (batch, token, head) element strides;
// causal masks are bottom-right aligned (query i sees keys j <= i + Sk - Sq).
extern "C" int mb_attn_fwd(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, int64_t k_bs,
                           int64_t k_ts, int64_t k_hs, const void* v, int64_t v_bs, int64_t v_ts, int64_t v_hs,
                           void* out, int64_t o_bs, int64_t o_ts, int64_t o_hs, int B, int Sq, int Sk, int Hq, int Hkv,
                           int hd, float scale, int causal, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_attn_fwd: no sm_100 device");
  MB_CHECK_ARG((hd == 64 || hd == 128) && Hkv >= 1 && Hq % Hkv == 0 && B >= 0 && B <= 65535 && Hq <= 65535,
               MB_ERR_SHAPE, "mb_attn_fwd: head_dim must be 64 or 128 and Hq %% Hkv == 0");
  MB_CHECK_ARG(q_ts % 8 == 0 && k_ts % 8 == 0 && v_ts % 8 == 0 && o_ts % 8 == 0 && q_hs % 8 == 0 && k_hs % 8 == 0 &&
                   v_hs % 8 == 0 && o_hs % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 && o_bs % 8 == 0,
               MB_ERR_ALIGN, "mb_attn_fwd: all strides must be multiples of 8 elements");
  if (B == 0 || Sq == 0) return MB_OK;
  AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v); p.o = static_cast<__nv_bfloat16*>(out);
  p.q_bs = q_bs; p.q_ts = q_ts; p.q_hs = q_hs;
  p.k_bs = k_bs; p.k_ts = k_ts; p.k_hs = k_hs;
  p.v_bs = v_bs; p.v_ts = v_ts; p.v_hs = v_hs;
  p.o_bs = o_bs; p.o_ts = o_ts; p.o_hs = o_hs;
  p.Sq = Sq; p.Sk = Sk; p.Hq = Hq; p.Hkv = Hkv;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  if (Sk >= 1) {
    const int rc = launch_attn_tc(q, q_bs, q_ts, q_hs, k, k_bs, k_ts, k_hs, v, v_bs, v_ts, v_hs, out, o_bs, o_ts, o_hs, B,
                                  Sq, Sk, Hq, Hkv, hd, scale, causal, stream);
    if (rc != 0) return rc < 0 ? rc : MB_OK;
  }
  return hd == 64 ? launch_attn<64>(p, B, stream) : launch_attn<128>(p, B, stream);
}
