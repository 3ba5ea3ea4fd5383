// tcgen05 / TMEM / TMA flash-attention forward for sm_100a:  O = softmax(scale * Q K^T [+ causal mask]) V.
// bf16 in, fp32 scores / softmax statistics / accumulation, bf16 out; head_dim 64 or 128; GQA (Hq % Hkv == 0).
// Replaces flash_attn_func / the eager softmax path of the reference (mingtok/vision_transformer/layers/attention.py:
// 61-74, 94-108, 138-163, 213-239; mingunivision/modeling_bailing_moe.py:946-1007).
//
// Persistent CTAs, each walking a static list of work items (one item = 128 query rows of one (batch, head)).  Roles:
//   warps 0..7  softmax   : two threads per query row (= TMEM lane), each taking 64 of the 128 keys of a block and half
//                           of the output columns.  Per 128-key block: read the score row from TMEM
//                           (tcgen05.ld), running max / sum, p = exp2(c s - c m) -> bf16 -> 128B-swizzled shared memory
//                           (the A operand of the P.V MMA); the running output O_row = alpha O_row + (P V)_row lives in
//                           registers and is updated ONE BLOCK LATE (P V of block j is fetched from TMEM after the
//                           softmax of block j + 1), so the tensor-core latency of P.V is off the softmax critical path.
//   warp 8      TMA       : Q tiles (double buffered) and K_0, V_0, K_1, V_1, ... through a 3-slot ring; runs ahead of
//                           the compute across work items, so the load latency of the next item is hidden.
//   warp 9      MMA       : S = Q K_j^T  (M 128 x N keys x K head_dim, both operands K-major) into TMEM cols [0, 128)
//                           O_j = P V_j  (M 128 x N 64 per 64-wide head_dim box x K keys; V is the MN-major B operand
//                           straight from its [keys][head_dim] TMA tile, no transpose) into one of two TMEM buffers.
// Two CTAs per SM at head_dim 64 (112 KB shared memory, 256 TMEM columns each) overlap one CTA's MUFU-bound softmax
// with the other's tensor-core work.  Key blocks are clipped to a multiple of 16 keys (MMA N / K granularity), so
// S = 65 costs 80 keys, not 128.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

constexpr int kAttThreads = 320;  // warps 0..7 softmax (2 threads per query row), warp 8 TMA, warp 9 MMA
constexpr int kAttRing = 3;

struct AttnTcParams {
  __nv_bfloat16* o;
  int64_t o_bs, o_ts, o_hs;  // element strides of out[b][s][h][:]
  int Sq, Sk, Hq, Hkv, B;
  float scale_log2;  // softmax scale * log2(e)
  int causal;        // bottom-right aligned: query i sees keys j <= i + (Sk - Sq)
  long long* dbg;    // optional phase timestamps of CTA 0 (mb_attn_set_debug; development only)
};

template <int HD>
struct AttnTcCfg {
  static constexpr int kBoxes = HD / 64;                 // 64-element (128-byte) column boxes per row
  static constexpr int kTileBytes = 128 * 128 * kBoxes;  // one Q / K / V tile: 128 rows x HD bf16
  static constexpr int kPBytes = 128 * 128 * 2;          // P tile: 128 rows x 128 keys bf16 = two K-major 64-key boxes
  // exchange area of the two threads that share a query row: [2 halves][128] bf16 row maxima, reused as [128] fp32 for
  // the row-sum hand-over at the end of a work item.  (Every byte counts: two CTAs must fit the SM's 228 KB.)
  static constexpr int kXchBytes = 512;
  static constexpr int kSmemBytes = kTileBytes * (2 + kAttRing) + kPBytes + kXchBytes + 128 /*barriers*/;
  static constexpr int kTmemCols = (HD == 64) ? 256 : 512;  // S: [0, 128), O buffers: [128, 128 + HD), [128 + HD, 128 + 2 HD)
  static constexpr int kMinBlocks = (HD == 64) ? 2 : 1;
  static_assert(kMinBlocks * (kSmemBytes + 1024) <= 233472, "shared memory budget (228 KB per SM, 1 KB reserved per CTA)");
};

template <int HD>
__global__ void __launch_bounds__(kAttThreads, AttnTcCfg<HD>::kMinBlocks)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const AttnTcParams p) {
  using Cfg = AttnTcCfg<HD>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;                                   // [2] Q tiles
  uint8_t* ring_s = smem + 2 * Cfg::kTileBytes;          // [3] K / V tiles
  uint8_t* p_s = ring_s + kAttRing * Cfg::kTileBytes;
  __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(p_s + Cfg::kPBytes);
  float* lxch = reinterpret_cast<float*>(xch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + Cfg::kPBytes + Cfg::kXchBytes);
  uint64_t* q_full = bars;             // [2] TMA -> MMA
  uint64_t* q_empty = bars + 2;        // [2] MMA -> TMA
  uint64_t* kv_full = bars + 4;        // [3] TMA -> MMA
  uint64_t* kv_empty = bars + 7;       // [3] MMA -> TMA
  uint64_t* s_full = bars + 10;        // MMA -> softmax (scores of block g in TMEM)
  uint64_t* p_full = bars + 11;        // softmax -> MMA (P of block g in shared memory; S of block g consumed)
  uint64_t* o_full = bars + 12;        // [2] MMA -> softmax (P V of block g in TMEM buffer g & 1; P smem consumed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.Sq + 127) >> 7;
  const int n_items = nq * p.Hq * p.B;
  const int shift = p.Sk - p.Sq;  // causal: query row r sees keys <= r + shift
  const int gqa = p.Hq / p.Hkv;
  // work item -> (q tile, head, batch); q tiles of one (batch, head) are adjacent, so CTAs running side by side share
  // their K / V through the L2
  auto kv_end_of = [&](int q0) { return p.causal ? min(p.Sk, q0 + 128 + shift) : p.Sk; };

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("attn_tc_kernel: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < kAttRing; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);  // one arrival per softmax warp
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  pdl_launch_dependents();
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      pdl_wait();
      int qi = 0, ld = 0;  // running Q-tile and K/V-tile counters (slots and parities continue across work items)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++qi) {
        const int q0 = (item % nq) * 128, h = (item / nq) % p.Hq, b = item / (nq * p.Hq);
        const int hk = h / gqa;
        const int nblk = (kv_end_of(q0) + 127) >> 7;
        const int qs = qi & 1;
        mbar_wait(&q_empty[qs], ((qi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], Cfg::kTileBytes);
#pragma unroll
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(&tmap_q, &q_full[qs], q_s + qs * Cfg::kTileBytes + bx * 16384, bx * 64, h, q0, b);
        for (int i = 0; i < 2 * nblk; ++i, ++ld) {
          const int slot = ld % kAttRing;
          mbar_wait(&kv_empty[slot], ((ld / kAttRing) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[slot], Cfg::kTileBytes);
          const CUtensorMap* tm = (i & 1) ? &tmap_v : &tmap_k;
#pragma unroll
          for (int bx = 0; bx < Cfg::kBoxes; ++bx)
            tma_load_4d(tm, &kv_full[slot], ring_s + slot * Cfg::kTileBytes + bx * 16384, bx * 64, hk, (i >> 1) * 128, b);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      int qi = 0, ld = 0, g = 0;  // g: running key-block counter (parities of s_full / p_full / o_full)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++qi) {
        const int q0 = (item % nq) * 128;
        const int kv_end = kv_end_of(q0);
        const int nblk = (kv_end + 127) >> 7;
        const int qs = qi & 1;
        const uint8_t* qt = q_s + qs * Cfg::kTileBytes;
        mbar_wait(&q_full[qs], (qi >> 1) & 1);
        for (int j = 0; j < nblk; ++j, ++g) {
          const int nk = min(128, kv_end - j * 128);
          const int nk16 = (nk + 15) & ~15;
          {  // S_g = Q K_j^T   (the S columns are free: p_full of block g - 1 was waited for before its P V)
            const int slot = ld % kAttRing;
            const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && g < 64;
            if (stamp) p.dbg[g * 16 + 8] = clock64();
            mbar_wait(&kv_full[slot], (ld / kAttRing) & 1);
            tc_fence_after();
            if (stamp) p.dbg[g * 16 + 9] = clock64();
            const uint32_t idesc = umma_idesc_bf16(128, nk16);
            // descriptors are built ONCE per block; the k steps only add constants (the issuing thread is on the
            // critical path of every block: a descriptor rebuilt per MMA cost ~120 cycles per instruction)
            const uint64_t da0 = umma_desc_sw128_kmajor(smem_u32(qt));
            const uint64_t db0 = umma_desc_sw128_kmajor(smem_u32(ring_s + slot * Cfg::kTileBytes));
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks) {
              const uint64_t off = static_cast<uint64_t>((ks >> 2) * (16384 >> 4) + 2 * (ks & 3));
              umma_bf16(tmem_s, da0 + off, db0 + off, idesc, ks != 0);
            }
            umma_commit(&kv_empty[slot]);
            if (j == nblk - 1) umma_commit(&q_empty[qs]);
            umma_commit(s_full);
            if (stamp) p.dbg[g * 16 + 10] = clock64();
            ++ld;
          }
          {  // O_g = P_g V_j into TMEM buffer g & 1 (fresh accumulator; the softmax warps keep the running O)
            const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && g < 64;
            mbar_wait(p_full, g & 1);
            tc_fence_after();
            if (stamp) p.dbg[g * 16 + 11] = clock64();
            const int slot = ld % kAttRing;
            mbar_wait(&kv_full[slot], (ld / kAttRing) & 1);
            tc_fence_after();
            if (stamp) p.dbg[g * 16 + 12] = clock64();
            constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | kUmmaBMajorMN;
            const uint64_t da0 = umma_desc_sw128_kmajor(smem_u32(p_s));
            const uint64_t db0 = umma_desc_sw128_mnmajor(smem_u32(ring_s + slot * Cfg::kTileBytes), 16384);
            const uint32_t d_o = tmem_o + (g & 1) * HD;
            const int nks = nk16 >> 4;
#pragma unroll
            for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                if (ks < nks) {
                  // A: 64-key boxes of P are 16 KB apart, 32 B per K step inside a box; B: 16 keys = two 8-row groups
                  // of the [keys][64] tile = 2048 B per K step, 64-column boxes 16 KB apart
                  const uint64_t offa = static_cast<uint64_t>((ks >> 2) * (16384 >> 4) + 2 * (ks & 3));
                  const uint64_t offb = static_cast<uint64_t>(bx * (16384 >> 4) + ks * (2048 >> 4));
                  umma_bf16(d_o + bx * 64, da0 + offa, db0 + offb, idesc, ks != 0);
                }
              }
            }
            umma_commit(&kv_empty[slot]);
            umma_commit(&o_full[g & 1]);
            if (stamp) p.dbg[g * 16 + 13] = clock64();
            ++ld;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / accumulate
    // TWO threads per query row: warps w and w + 4 both own TMEM lane quadrant w & 3; half hf = w >> 2 takes keys
    // [64 hf, 64 hf + 64) of every 128-key block (= one 64-key K-major box of P) and columns [HD/2 hf, HD/2 (hf + 1))
    // of the output.  The halves exchange their partial row maximum through shared memory once per block (named
    // barrier of the 64 threads that share a quadrant); the partial row sums are combined once per work item.
    const int quad = warp & 3, hf = warp >> 2;
    const int r_local = quad * 32 + lane;       // 0..127 == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale_log2;
    constexpr int HO = HD / 2;                  // output columns per thread
    pdl_wait();  // `out` may still be read by the predecessor kernel; also orders our stores after it
    int g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int q0 = (item % nq) * 128, h = (item / nq) % p.Hq, b = item / (nq * p.Hq);
      const int kv_end = kv_end_of(q0);
      const int nblk = (kv_end + 127) >> 7;
      const int row = q0 + r_local;
      const int key_lim = p.causal ? min(p.Sk, row + shift + 1) : p.Sk;  // keys [0, key_lim) are visible to this row
      float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float o[HO];
#pragma unroll
      for (int i = 0; i < HO; ++i) o[i] = 0.f;
      // O_run = alpha_prev * O_run + (P V)_{g-1}: the update for block g - 1, executed after the softmax of block g
      auto accumulate_prev = [&](int gp) {
        mbar_wait(&o_full[gp & 1], (gp >> 1) & 1);
        tc_fence_after();
        const uint32_t src = tmem_o + (gp & 1) * HD + hf * HO + lane_base;
#pragma unroll
        for (int hh = 0; hh < HO / 32; ++hh) {
          uint32_t t0[32];
          tmem_ld_32x32b_x32(src + hh * 32, t0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[hh * 32 + i] = fmaf(o[hh * 32 + i], alpha_prev, __uint_as_float(t0[i]));
        }
      };
      for (int j = 0; j < nblk; ++j, ++g) {
        const int k0 = j * 128 + hf * 64;                      // first key of this thread's half block
        const int nk = min(64, kv_end - k0);                   // may be <= 0: nothing to do for this half
        const int nch = nk > 0 ? (nk + 31) >> 5 : 0;           // 32-key chunks of this half the MMAs produce / consume
        const uint32_t s_addr = tmem_s + lane_base + hf * 64;
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && g < 64;
        if (stamp) p.dbg[g * 16 + 0] = clock64();
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        if (stamp) p.dbg[g * 16 + 1] = clock64();
        // pass 1: maximum over the visible keys of this half, then exchange with the other half of the row
        float mx = -INFINITY;
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t s0[32];
          tmem_ld_32x32b_x32(s_addr + ch * 32, s0);
          tmem_ld_wait();
          const int kb = k0 + ch * 32;
          if (kb + 32 <= key_lim) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s0[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kb + i < key_lim) mx = fmaxf(mx, __uint_as_float(s0[i]));
          }
        }
        // both halves must end up with the SAME reference value: they exchange the partial maxima rounded UP to bf16 and
        // each takes the maximum of the two rounded values (any m >= the true maximum is a valid softmax reference)
        if (stamp) p.dbg[g * 16 + 2] = clock64();
        const __nv_bfloat16 mx_b = __float2bfloat16_ru(mx);
        xch[hf * 128 + r_local] = mx_b;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float m_new = fmaxf(m, fmaxf(__bfloat162float(mx_b), __bfloat162float(xch[(hf ^ 1) * 128 + r_local])));
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // xch is rewritten by the next block
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far (rows past Sq only): avoid NaN
        const float alpha = fast_ex2((m - m_use) * c);           // m = -inf on the first block -> 0
        const float mc = m_use * c;
        // the P tile is free once P V of the previous block has completed (long done in steady state)
        if (stamp) p.dbg[g * 16 + 3] = clock64();
        if (j > 0) mbar_wait(&o_full[(g - 1) & 1], ((g - 1) >> 1) & 1);
        if (stamp) p.dbg[g * 16 + 4] = clock64();
        // pass 2: p = exp2(c s - c m), partial row sum, bf16 P into this half's swizzled 64-key box
        float sum = 0.f;
        uint8_t* prow = p_s + hf * 16384 + r_local * 128;
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t s[32];
          tmem_ld_32x32b_x32(s_addr + ch * 32, s);
          tmem_ld_wait();
          const int kb = k0 + ch * 32;
          float pv[32];
          if (kb + 32 <= key_lim) {
#pragma unroll
            for (int i = 0; i < 32; ++i) pv[i] = fast_ex2(fmaf(__uint_as_float(s[i]), c, -mc));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              pv[i] = (kb + i < key_lim) ? fast_ex2(fmaf(__uint_as_float(s[i]), c, -mc)) : 0.f;
          }
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 32; ++i) s4[i & 3] += pv[i];
          sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16x2(pv[8 * q + 0], pv[8 * q + 1]);
            v.y = pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]);
            v.z = pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]);
            v.w = pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]);
            const int chunk16 = ch * 4 + q;
            *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (r_local & 7)) << 4)) = v;
          }
        }
        if (stamp) p.dbg[g * 16 + 5] = clock64();
        fence_proxy_async_smem();  // P stores -> visible to the tensor core's (async proxy) reads
        tc_fence_before();         // the TMEM reads of S are complete (tmem_ld_wait) before the MMA warp reuses S
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (stamp) p.dbg[g * 16 + 6] = clock64();
        // while the tensor core runs P V of this block (and S of the next): fold in P V of the PREVIOUS block
        if (j > 0) accumulate_prev(g - 1);
        if (stamp) p.dbg[g * 16 + 7] = clock64();
        l = l * alpha + sum;
        m = m_new;
        alpha_prev = alpha;
      }
      accumulate_prev(g - 1);  // the last block of this work item
      tc_fence_before();
      // row sum = this half's partial sum + the other half's (same alpha sequence, so the partials simply add)
      if (hf == 0) lxch[r_local] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      if (hf == 1) {
        l += lxch[r_local];
        lxch[r_local] = l;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      if (hf == 0) l = lxch[r_local];
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // the area is rewritten by the next work item
      if (row < p.Sq) {
        const float inv = 1.0f / l;
        __nv_bfloat16* dst = p.o + static_cast<int64_t>(b) * p.o_bs + static_cast<int64_t>(row) * p.o_ts +
                             static_cast<int64_t>(h) * p.o_hs + hf * HO;
#pragma unroll
        for (int q = 0; q < HO / 8; ++q) {
          uint4 v;
          v.x = pack_bf16x2(o[8 * q + 0] * inv, o[8 * q + 1] * inv);
          v.y = pack_bf16x2(o[8 * q + 2] * inv, o[8 * q + 3] * inv);
          v.z = pack_bf16x2(o[8 * q + 4] * inv, o[8 * q + 5] * inv);
          v.w = pack_bf16x2(o[8 * q + 6] * inv, o[8 * q + 7] * inv);
          reinterpret_cast<uint4*>(dst)[q] = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int HD>
static int launch_attn_tc_hd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& p,
                             int B, cudaStream_t stream) {
  using Cfg = AttnTcCfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  const long items = static_cast<long>((p.Sq + 127) / 128) * p.Hq * B;
  const long resident = static_cast<long>(num_sms()) * Cfg::kMinBlocks;
  cfg.gridDim = dim3(static_cast<unsigned>(items < resident ? items : resident));
  cfg.blockDim = dim3(kAttThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_tc_kernel<HD>, tq, tk, tv, p));
  return MB_OK;
}

static long long* g_attn_dbg = nullptr;
static int g_attn_backend = getenv("MB_ATTN_BACKEND") ? atoi(getenv("MB_ATTN_BACKEND")) : 0;  // 0 auto, 1 tcgen05, 2 mma.sync

// 0: not eligible (caller falls back to the mma.sync kernel), 1: launched, < 0: error.
int launch_attn_tc(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, int64_t k_bs, int64_t k_ts,
                   int64_t k_hs, const void* v, int64_t v_bs, int64_t v_ts, int64_t v_hs, void* out, int64_t o_bs,
                   int64_t o_ts, int64_t o_hs, int B, int Sq, int Sk, int Hq, int Hkv, int hd, float scale, int causal,
                   cudaStream_t stream) {
  if (g_attn_backend == 2) return 0;
  if (hd != 64 && hd != 128) return 0;
  if (Sk < Sq && causal) return 0;  // rows without any visible key: keep the legacy kernel's convention
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) != 0)
    return 0;
  CUtensorMap tq, tk, tv;
  const uint32_t box[4] = {64, 1, 128, 1};
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(Hq), static_cast<uint64_t>(Sq),
                              static_cast<uint64_t>(B)};
    const uint64_t st[3] = {static_cast<uint64_t>(q_hs), static_cast<uint64_t>(q_ts), static_cast<uint64_t>(q_bs)};
    if (!make_tmap_4d_bf16(&tq, q, dims, st, box)) return MB_ERR_CUDA;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(hd), static_cast<uint64_t>(Hkv), static_cast<uint64_t>(Sk),
                              static_cast<uint64_t>(B)};
    const uint64_t sk[3] = {static_cast<uint64_t>(k_hs), static_cast<uint64_t>(k_ts), static_cast<uint64_t>(k_bs)};
    const uint64_t sv[3] = {static_cast<uint64_t>(v_hs), static_cast<uint64_t>(v_ts), static_cast<uint64_t>(v_bs)};
    if (!make_tmap_4d_bf16(&tk, k, dims, sk, box)) return MB_ERR_CUDA;
    if (!make_tmap_4d_bf16(&tv, v, dims, sv, box)) return MB_ERR_CUDA;
  }
  AttnTcParams p;
  p.o = static_cast<__nv_bfloat16*>(out);
  p.o_bs = o_bs; p.o_ts = o_ts; p.o_hs = o_hs;
  p.Sq = Sq; p.Sk = Sk; p.Hq = Hq; p.Hkv = Hkv; p.B = B;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.dbg = g_attn_dbg;
  const int rc = (hd == 64) ? launch_attn_tc_hd<64>(tq, tk, tv, p, B, stream)
                            : launch_attn_tc_hd<128>(tq, tk, tv, p, B, stream);
  return rc == MB_OK ? 1 : rc;
}

}  // namespace mb

// Development aid: device buffer of >= 64 * 16 int64 that receives clock64() phase stamps of CTA 0 (NULL = off).
extern "C" int mb_attn_set_debug(void* dev_buf) {
  mb::g_attn_dbg = static_cast<long long*>(dev_buf);
  return MB_OK;
}

extern "C" int mb_attn_set_backend(int backend) {
  MB_CHECK_ARG(backend >= 0 && backend <= 2, MB_ERR_SHAPE, "mb_attn_set_backend: 0 auto, 1 tcgen05, 2 mma.sync");
  mb::g_attn_backend = backend;
  return MB_OK;
}
