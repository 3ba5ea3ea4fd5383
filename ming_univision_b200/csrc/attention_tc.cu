// tcgen05 / TMEM / TMA flash-attention forward for sm_100a:  O = softmax(scale * Q K^T [+ causal mask]) V.
// bf16 in, fp32 scores / softmax statistics / accumulation, bf16 out; head_dim 64 or 128; GQA (Hq % Hkv == 0).
// Replaces flash_attn_func / the eager softmax path of the reference (mingtok/vision_transformer/layers/attention.py:
// 61-74, 94-108, 138-163, 213-239; mingunivision/modeling_bailing_moe.py:946-1007).
//
// Persistent CTAs, each walking a static list of work items (one item = 128 query rows of one (batch, head)); the keys
// of an item are processed in blocks of BN (96 at head_dim 64, 128 at head_dim 128).  Roles:
//   warps 0..7  softmax   : two threads per query row (= TMEM lane), each taking half of the keys of a block and half
//                           of the output columns.  Per block: read the score row from TMEM (tcgen05.ld), running
//                           max / sum, p = exp2(c s - c m) -> bf16 -> 128B-swizzled shared memory (the A operand of the
//                           P.V MMA); the running output O_row = alpha O_row + (P V)_row lives in registers and is
//                           updated ONE BLOCK LATE (P V of block g is fetched from TMEM during block g + 1).
//   warp 8      TMA       : Q tiles (double buffered) and K_0, V_0, K_1, V_1, ... through a 2-slot ring; runs ahead of
//                           the compute across work items, so the load latency of the next item is hidden.
//   warp 9      MMA       : S_g = Q K^T (M 128 x N keys x K head_dim, both operands K-major) into one of TWO score
//                           buffers in TMEM, issued one block AHEAD (S_{g+1} runs on the tensor core while the softmax
//                           warps work on S_g); O_g = P_g V (M 128 x N 64 per 64-wide head_dim box x K keys; V is the
//                           MN-major B operand straight from its [keys][head_dim] TMA tile, no transpose).
// So neither MMA nor its issue latency is on the softmax warps' critical path.  TMEM: 2 BN (scores) + HD (P V) columns
// = 256 at head_dim 64, which lets two CTAs share an SM (105 KB shared memory each) and overlap their MUFU phases.
// P is double-buffered in shared memory, so the softmax of block g + 1 never waits for P V of block g.
// Key blocks are clipped to a multiple of 16 keys (MMA N / K granularity), so S = 65 costs 80 keys, not 96.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

constexpr int kAttThreads = 320;  // warps 0..7 softmax (2 threads per query row), warp 8 TMA, warp 9 MMA
constexpr int kAttRing = 2;  // K_g always lands in slot 0, V_g in slot 1

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
  static constexpr int kBN = (HD == 64) ? 96 : 128;      // keys per block
  static constexpr int kBoxes = HD / 64;                 // 64-element (128-byte) column boxes per row
  static constexpr int kQBytes = 128 * 128 * kBoxes;     // Q tile: 128 rows x HD bf16
  static constexpr int kKVBytes = kBN * 128 * kBoxes;    // K / V tile: BN rows x HD bf16
  // P tiles: TWO buffers (block g writes buffer g & 1 while P V of block g - 1 may still read the other one).  A tile is
  // two K-major 64-key boxes of 128-byte rows.  With 96-key blocks the second box holds only 32 keys = the first 64
  // bytes of every row, so both buffers SHARE one second box: buffer 0 uses 16-byte chunks 0..3 of each row, buffer 1
  // chunks 4..7 (logical chunks; the 128-byte swizzle permutes chunks within a row, so the two never collide).
  static constexpr bool kShareBox1 = kBN <= 96;
  static constexpr int kPBytes = kShareBox1 ? 3 * 16384 : 4 * 16384;
  // exchange area of the two threads that share a query row: [2 halves][128] bf16 row maxima, reused as [128] fp32 for
  // the row-sum hand-over at the end of a work item.  (Every byte counts: two CTAs must fit the SM's 228 KB.)
  static constexpr int kXchBytes = 512;
  static constexpr int kSmemBytes = 2 * kQBytes + kAttRing * kKVBytes + kPBytes + kXchBytes + 160 /*barriers*/;
  static constexpr int kTmemCols = (HD == 64) ? 256 : 512;  // S0: [0, BN), S1: [BN, 2 BN), O: [2 BN, 2 BN + HD)
  static constexpr int kMinBlocks = (HD == 64) ? 2 : 1;
  static_assert(2 * kBN + HD <= kTmemCols, "TMEM budget");
  static_assert(kMinBlocks * (kSmemBytes + 1024) <= 233472, "shared memory budget (228 KB per SM, 1 KB reserved per CTA)");
};

// Cursor over the key blocks of this CTA's work items, in processing order.
struct AttnCursor {
  int item, j, nblk, kv_end, qi;
  bool valid;
};

template <int HD>
__global__ void __launch_bounds__(kAttThreads, AttnTcCfg<HD>::kMinBlocks)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const AttnTcParams p) {
  using Cfg = AttnTcCfg<HD>;
  constexpr int BN = Cfg::kBN;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;                                   // [2] Q tiles
  uint8_t* ring_s = smem + 2 * Cfg::kQBytes;             // [4] K / V tiles
  uint8_t* p_s = ring_s + kAttRing * Cfg::kKVBytes;
  __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(p_s + Cfg::kPBytes);
  float* lxch = reinterpret_cast<float*>(xch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + Cfg::kPBytes + Cfg::kXchBytes);
  uint64_t* q_full = bars;             // [2] TMA -> MMA
  uint64_t* q_empty = bars + 2;        // [2] MMA -> TMA
  uint64_t* kv_full = bars + 4;        // [4] TMA -> MMA
  uint64_t* kv_empty = bars + 8;       // [4] MMA -> TMA
  uint64_t* s_full = bars + 12;        // [2] MMA -> softmax (scores of block g in TMEM buffer g & 1)
  uint64_t* p_full = bars + 14;        // softmax -> MMA (P of block g in shared memory; S_g and O_{g-1} consumed)
  uint64_t* o_full = bars + 15;        // MMA -> softmax (P V of block g in TMEM; P shared memory consumed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.Sq + 127) >> 7;
  const int n_items = nq * p.Hq * p.B;
  const int shift = p.Sk - p.Sq;  // causal: query row r sees keys <= r + shift
  const int gqa = p.Hq / p.Hkv;
  // work item -> (q tile, head, batch); q tiles of one (batch, head) are adjacent, so CTAs running side by side share
  // their K / V through the L2
  auto kv_end_of = [&](int q0) { return p.causal ? min(p.Sk, q0 + 128 + shift) : p.Sk; };
  auto cursor_at = [&](int item, int qi) {
    AttnCursor cu;
    cu.item = item; cu.j = 0; cu.qi = qi;
    cu.valid = item < n_items;
    cu.kv_end = cu.valid ? kv_end_of((item % nq) * 128) : 0;
    cu.nblk = (cu.kv_end + BN - 1) / BN;
    return cu;
  };
  auto advance = [&](AttnCursor& cu) {
    if (++cu.j == cu.nblk) cu = cursor_at(cu.item + gridDim.x, cu.qi + 1);
  };

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
      mbar_init(&s_full[i], 1);
    }
    for (int i = 0; i < kAttRing; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(p_full, 8);  // one arrival per softmax warp
    mbar_init(o_full, 1);
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
  const uint32_t tmem_o = tmem_base + 2 * BN;

  pdl_launch_dependents();
  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      pdl_wait();
      // K stream (slot 0) runs ONE BLOCK AHEAD of the V stream (slot 1), mirroring the order in which the MMA warp
      // consumes them (S_{g+1} is issued before P V of block g): K_{g+1} is requested as soon as S_g has read its slot,
      // V_g as soon as P V of block g - 1 has.  The Q tile of a work item travels with the K tile of its first block.
      auto load_k = [&](const AttnCursor& cu, int g) {
        const int q0 = (cu.item % nq) * 128, h = (cu.item / nq) % p.Hq, b = cu.item / (nq * p.Hq);
        if (cu.j == 0) {
          const int qs = cu.qi & 1;
          mbar_wait_sleep(&q_empty[qs], ((cu.qi >> 1) & 1) ^ 1, 100);
          mbar_arrive_expect_tx(&q_full[qs], Cfg::kQBytes);
#pragma unroll
          for (int bx = 0; bx < Cfg::kBoxes; ++bx)
            tma_load_4d(&tmap_q, &q_full[qs], q_s + qs * Cfg::kQBytes + bx * 16384, bx * 64, h, q0, b);
        }
        mbar_wait_sleep(&kv_empty[0], (g & 1) ^ 1, 100);
        mbar_arrive_expect_tx(&kv_full[0], Cfg::kKVBytes);
#pragma unroll
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(&tmap_k, &kv_full[0], ring_s + bx * (BN * 128), bx * 64, h / gqa, cu.j * BN, b);
      };
      auto load_v = [&](const AttnCursor& cu, int g) {
        const int h = (cu.item / nq) % p.Hq, b = cu.item / (nq * p.Hq);
        mbar_wait_sleep(&kv_empty[1], (g & 1) ^ 1, 100);
        mbar_arrive_expect_tx(&kv_full[1], Cfg::kKVBytes);
#pragma unroll
        for (int bx = 0; bx < Cfg::kBoxes; ++bx)
          tma_load_4d(&tmap_v, &kv_full[1], ring_s + Cfg::kKVBytes + bx * (BN * 128), bx * 64, h / gqa, cu.j * BN, b);
      };
      AttnCursor ck = cursor_at(blockIdx.x, 0), cv = ck;
      int gk = 0, gv = 0;
      if (ck.valid) { load_k(ck, gk++); advance(ck); }
      while (cv.valid) {
        if (ck.valid) { load_k(ck, gk++); advance(ck); }
        load_v(cv, gv++);
        advance(cv);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      // S_g = Q K^T into score buffer g & 1.  Its previous content S_{g-2} was consumed before p_full of block g - 2,
      // which this thread has waited for (ahead of P V of block g - 2) by the time it gets here.
      auto issue_s = [&](const AttnCursor& cu, int g) {
        const int qs = cu.qi & 1;
        if (cu.j == 0) mbar_wait(&q_full[qs], (cu.qi >> 1) & 1);
        constexpr int slot = 0;  // K tiles
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && g < 64;
        if (stamp) p.dbg[g * 16 + 8] = clock64();
        mbar_wait(&kv_full[slot], g & 1);
        tc_fence_after();
        const int nk16 = (min(BN, cu.kv_end - cu.j * BN) + 15) & ~15;
        const uint32_t idesc = umma_idesc_bf16(128, nk16);
        // descriptors as (low word, shared high word): K steps / boxes are 32-bit adds on the low word (see ptx.cuh)
        const uint64_t da0 = umma_desc_sw128_kmajor(smem_u32(q_s + qs * Cfg::kQBytes));
        const uint64_t db0 = umma_desc_sw128_kmajor(smem_u32(ring_s + slot * Cfg::kKVBytes));
        const uint32_t hi = static_cast<uint32_t>(da0 >> 32);
        const uint32_t a_lo = static_cast<uint32_t>(da0), b_lo = static_cast<uint32_t>(db0);
        const uint32_t d_s = tmem_base + (g & 1) * BN;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          // 64-column boxes: Q boxes are 16 KB apart, K boxes BN * 128 B apart; 32 B per K step inside a box
          const uint32_t offa = (ks >> 2) * (16384 >> 4) + 2 * (ks & 3);
          const uint32_t offb = (ks >> 2) * ((BN * 128) >> 4) + 2 * (ks & 3);
          umma_bf16_lo(d_s, a_lo + offa, b_lo + offb, hi, idesc, ks != 0);
        }
        umma_commit(&kv_empty[slot]);
        if (cu.j == cu.nblk - 1) umma_commit(&q_empty[qs]);
        umma_commit(&s_full[g & 1]);
        if (stamp) p.dbg[g * 16 + 9] = clock64();
      };
      // O_g = P_g V (fresh accumulator; the softmax warps keep the running O in registers)
      auto issue_pv = [&](const AttnCursor& cu, int g) {
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && g < 64;
        if (stamp) p.dbg[g * 16 + 10] = clock64();
        mbar_wait_sleep(p_full, g & 1, 20);
        tc_fence_after();
        if (stamp) p.dbg[g * 16 + 11] = clock64();
        constexpr int slot = 1;  // V tiles
        mbar_wait(&kv_full[slot], g & 1);
        tc_fence_after();
        if (stamp) p.dbg[g * 16 + 12] = clock64();
        const int nks = (min(BN, cu.kv_end - cu.j * BN) + 15) >> 4;
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64) | kUmmaBMajorMN;
        const int pb = g & 1;  // P buffer of this block
        const uint64_t da0 = umma_desc_sw128_kmajor(smem_u32(p_s + pb * 16384));                       // keys 0..63
        const uint64_t da1 = umma_desc_sw128_kmajor(smem_u32(p_s + 32768 + (Cfg::kShareBox1 ? 0 : pb * 16384)));  // 64..
        const uint64_t db0 = umma_desc_sw128_mnmajor(smem_u32(ring_s + slot * Cfg::kKVBytes), BN * 128);
        const uint32_t hi = static_cast<uint32_t>(da0 >> 32);  // identical for both layouts (SBO, version, swizzle)
        const uint32_t a_lo0 = static_cast<uint32_t>(da0);
        const uint32_t a_lo1 = static_cast<uint32_t>(da1) + (Cfg::kShareBox1 ? pb * 4 : 0);  // shared box: chunks 4..7
        const uint32_t b_lo = static_cast<uint32_t>(db0);
#pragma unroll
        for (int bx = 0; bx < Cfg::kBoxes; ++bx) {
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            if (ks < nks) {
              // A: 32 B per K step inside a 64-key box; B: 16 keys = two 8-row groups of the [keys][64] tile = 2048 B
              // per K step, 64-column boxes BN * 128 B apart
              const uint32_t a_lo = (ks < 4 ? a_lo0 : a_lo1) + 2 * (ks & 3);
              const uint32_t offb = bx * ((BN * 128) >> 4) + ks * (2048 >> 4);
              umma_bf16_lo(tmem_o + bx * 64, a_lo, b_lo + offb, hi, idesc, ks != 0);
            }
          }
        }
        umma_commit(&kv_empty[slot]);
        umma_commit(o_full);
        if (stamp) p.dbg[g * 16 + 13] = clock64();
      };
      AttnCursor cs = cursor_at(blockIdx.x, 0), cp = cs;
      int gs = 0, gp = 0;
      if (cs.valid) { issue_s(cs, gs++); advance(cs); }
      while (cp.valid) {
        if (cs.valid) { issue_s(cs, gs++); advance(cs); }  // one block ahead of the softmax warps
        issue_pv(cp, gp++);
        advance(cp);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / accumulate
    // TWO threads per query row: warps w and w + 4 both own TMEM lane quadrant w & 3; half hf = w >> 2 takes keys
    // [BN/2 hf, BN/2 (hf + 1)) of every block and columns [HD/2 hf, HD/2 (hf + 1)) of the output.  The halves exchange
    // their partial row maximum through shared memory once per block (named barrier of the 64 threads that share a
    // quadrant); the partial row sums are combined once per work item.
    const int quad = warp & 3, hf = warp >> 2;
    const int r_local = quad * 32 + lane;       // 0..127 == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale_log2;
    constexpr int HO = HD / 2;                  // output columns per thread
    constexpr int HK = BN / 2;                  // keys per thread and block
    pdl_wait();  // `out` may still be read by the predecessor kernel; also orders our stores after it
    int g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int q0 = (item % nq) * 128, h = (item / nq) % p.Hq, b = item / (nq * p.Hq);
      const int kv_end = kv_end_of(q0);
      const int nblk = (kv_end + BN - 1) / BN;
      const int row = q0 + r_local;
      const int key_lim = p.causal ? min(p.Sk, row + shift + 1) : p.Sk;  // keys [0, key_lim) are visible to this row
      float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float o[HO];
#pragma unroll
      for (int i = 0; i < HO; ++i) o[i] = 0.f;
      // O_run = alpha_prev * O_run + (P V)_{g-1}: the update for block g - 1 (o_full already waited for)
      auto accumulate_prev = [&]() {
        const uint32_t src = tmem_o + hf * HO + lane_base;
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
        const int kblk = j * BN;                               // first key of the block
        const int nkb = min(BN, kv_end - kblk);                // keys of the block
        const int nk16 = (nkb + 15) & ~15;                     // columns the S MMA produced / the P V MMA consumes
        const int c_beg = hf * HK;                             // this thread's columns: [c_beg, c_end) in steps of 16
        const int c_end = min(c_beg + HK, nk16);
        const uint32_t s_addr = tmem_base + (g & 1) * BN + lane_base;
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && g < 64;
        if (stamp) p.dbg[g * 16 + 0] = clock64();
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        if (stamp) p.dbg[g * 16 + 1] = clock64();
        // ---- softmax of the block.  Reference value m of the exponent: the first block of an item takes the exact
        // block maximum (max pass + exp pass); later blocks keep the STALE m and run ONE pass with no rescaling of O,
        // as long as the block's partial row sums stay below kSumLimit (then every p <= kSumLimit: exact in fp32,
        // harmless in bf16) — the usual case.  A block that breaks the bound (or overflows to inf) is redone with its
        // true maximum.  Both threads of a row see the same exchanged flags / maxima, so they take the same decision.
        constexpr float kSumLimit = 1024.f;
        float sum = 0.f, alpha = 1.f, m_new = m;
        auto exchange = [&](float v) {  // -> max(v, partner's v), identical in both threads (values rounded UP to bf16)
          const __nv_bfloat16 vb = __float2bfloat16_ru(v);
          xch[hf * 128 + r_local] = vb;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
          const float r = fmaxf(__bfloat162float(vb), __bfloat162float(xch[(hf ^ 1) * 128 + r_local]));
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // xch is rewritten by the next exchange
          return r;
        };
        auto max_pass = [&]() {
          float mx = -INFINITY;
          for (int cc = c_beg; cc < c_end; cc += 16) {
            uint32_t s0[16];
            tmem_ld_32x32b_x16(s_addr + cc, s0);
            tmem_ld_wait();
            const int kb = kblk + cc;
            if (kb + 16 <= key_lim) {
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(s0[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (kb + i < key_lim) mx = fmaxf(mx, __uint_as_float(s0[i]));
            }
          }
          return mx;
        };
        // one pass: p = exp2(c s - mc) for this thread's keys -> bf16 -> swizzled K-major P boxes; partial row sum.
        // Unmasked 16-key groups run on packed fp32 pairs (FFMA2 / FADD2): ~3 issue slots per element.
        auto exp_pass = [&](float mc) {
          const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
          uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
          float tail = 0.f;
          uint8_t* const prow0 = p_s + r_local * 128;
          const int sw = r_local & 7, pb = g & 1;
          for (int cc = c_beg; cc < c_end; cc += 16) {
            uint32_t sr[16];
            tmem_ld_32x32b_x16(s_addr + cc, sr);
            tmem_ld_wait();
            const int kb = kblk + cc;
            uint32_t pk[8];
            if (kb + 16 <= key_lim) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float x0, x1;
                unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sr[2 * i]), __uint_as_float(sr[2 * i + 1])), c2, nmc2),
                             x0, x1);
                const float p0 = fast_ex2(x0), p1 = fast_ex2(x1);
                if (i & 1) acc1 = add_f32x2(acc1, pack_f32x2(p0, p1));
                else acc0 = add_f32x2(acc0, pack_f32x2(p0, p1));
                pk[i] = pack_bf16x2(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float p0 = (kb + 2 * i < key_lim) ? fast_ex2(fmaf(__uint_as_float(sr[2 * i]), c, -mc)) : 0.f;
                const float p1 = (kb + 2 * i + 1 < key_lim) ? fast_ex2(fmaf(__uint_as_float(sr[2 * i + 1]), c, -mc)) : 0.f;
                tail += p0 + p1;
                pk[i] = pack_bf16x2(p0, p1);
              }
            }
            // P buffer g & 1: keys 0..63 in its own box, keys 64.. in the second box (shared: chunk offset 4 for buffer 1)
            uint8_t* prow = prow0 + ((cc < 64) ? pb * 16384 : 32768 + (Cfg::kShareBox1 ? 0 : pb * 16384));
            const int ch = ((cc & 63) >> 3) + ((cc >= 64 && Cfg::kShareBox1) ? pb * 4 : 0);
            *reinterpret_cast<uint4*>(prow + ((ch ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          float a0, a1, a2, a3;
          unpack_f32x2(acc0, a0, a1);
          unpack_f32x2(acc1, a2, a3);
          sum = ((a0 + a1) + (a2 + a3)) + tail;
        };
        // true maximum of the block + rescale factor for what has been accumulated so far; warp-convergent (tcgen05.ld
        // and the named barrier are .aligned), `take` selects per row whether the new reference is adopted
        auto exact_block = [&](bool take) {
          const float bm = exchange(max_pass());
          if (take) {
            m_new = fmaxf(m, bm);
            if (m_new == -INFINITY) m_new = 0.f;  // no visible key so far (rows past Sq only): avoid NaN
            alpha = fast_ex2((m - m_new) * c);    // m = -inf on the first block -> 0
          }
          exp_pass(m_new * c);  // rows that keep their m recompute identical values
        };
        if (j == 0) {
          if (stamp) p.dbg[g * 16 + 2] = clock64();
          if (stamp) p.dbg[g * 16 + 3] = clock64();
          exact_block(true);  // (the P tile is free: the previous item waited for its last P V)
          if (stamp) p.dbg[g * 16 + 4] = clock64();
        } else {
          if (stamp) p.dbg[g * 16 + 2] = clock64();
          if (stamp) p.dbg[g * 16 + 3] = clock64();
          exp_pass(m * c);  // (P buffer g & 1 is free: its last reader, P V of block g - 2, was waited for one block ago)
          if (stamp) p.dbg[g * 16 + 4] = clock64();
          const float flag = exchange(!(sum <= kSumLimit) ? 1.f : 0.f);  // NaN / inf count as "over the limit"
          // rare: some row's scores outgrew the stale reference.  The two warps that share these rows see the same flags,
          // so both take this branch together (the exchange inside stays matched).
          if (__any_sync(0xffffffffu, flag != 0.f)) exact_block(flag != 0.f);
        }
        if (stamp) p.dbg[g * 16 + 5] = clock64();
        // fold in P V of the PREVIOUS block (it had the whole exponent pass to complete) before handing the O columns
        // back to the tensor core
        if (j > 0) {
          mbar_wait(o_full, (g - 1) & 1);
          tc_fence_after();
          accumulate_prev();
        }
        fence_proxy_async_smem();  // P stores -> visible to the tensor core's (async proxy) reads
        tc_fence_before();         // the TMEM reads of S_g and O_{g-1} are complete before the MMA warp reuses them
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (stamp) p.dbg[g * 16 + 6] = clock64();
        l = l * alpha + sum;
        m = m_new;
        alpha_prev = alpha;
      }
      mbar_wait(o_full, (g - 1) & 1);  // the last block of this work item
      tc_fence_after();
      accumulate_prev();
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
  const uint32_t box_kv[4] = {64, 1, static_cast<uint32_t>(hd == 64 ? AttnTcCfg<64>::kBN : AttnTcCfg<128>::kBN), 1};
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
    if (!make_tmap_4d_bf16(&tk, k, dims, sk, box_kv)) return MB_ERR_CUDA;
    if (!make_tmap_4d_bf16(&tv, v, dims, sv, box_kv)) return MB_ERR_CUDA;
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
