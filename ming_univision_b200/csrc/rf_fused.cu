// Rectified-flow head as ONE persistent weight-streaming kernel per sample (RectifiedFlowLoss.sample + SimpleMLPAdaLN,
// mingunivision/diff_loss_rf_swiglu.py:103-181, :363-385; ResBlock :268-272, FinalLayer :288-292, SwiGLUFFNFused :54-72).
//
// The head is a chain of 16 Euler steps x 12 residual blocks x 2 skinny GEMMs (M = CFG rows <= 3): 29 GB of bf16
// weights per visual token, pure HBM streaming.  As ~430 separate launches the stream stalls at every kernel boundary
// (ramp, tail, dependency gap: 25 us kernels reaching 55 % of the copy bandwidth).  Here the whole sampler is one
// launch of one CTA per SM:
//   * every CTA owns a FIXED slice of the output rows of every layer (rows c*R/G .. (c+1)*R/G of all K columns), so no
//     cross-CTA reduction exists and every sum has a fixed order (bit-reproducible);
//   * the weights are pre-packed once (mb_rf_pack_weights) into the order the CTAs consume them: per CTA, per 16-row MMA
//     tile, per 1024-wide K chunk one CONTIGUOUS block laid out [k-group of 32][row][32 k] — so a pipeline stage is ONE
//     cp.async.bulk (TMA, <= 32 KB) into shared memory and every tensor-core A fragment is one conflict-free 16-byte
//     shared load (the K order inside a fragment is a permutation the activation fragment shares, as in gemv.cu);
//   * a producer warp streams the stages through a ring of shared-memory buffers guarded by mbarriers and runs AHEAD OF
//     THE GRID BARRIERS — weights do not depend on activations — so HBM keeps streaming while the CTAs exchange the
//     tiny activation vectors (h [B, 3072], hid [B, 8192]) through L2 between layers;
//   * 8 consumer warps split the K range of every stage, multiply with mma.sync.m16n8k16 (fp32 accumulate), reduce
//     across warps in shared memory in a fixed order and apply the reference's epilogues with its rounding points:
//     adaLN modulate (fp32 LN, bf16(1 + scale)), bf16(silu(bf16 x1)) * bf16 x2, x + bf16(gate * bf16 h), CFG combine and
//     the fp32 Euler update.
// Algorithmic bytes per launch: steps * depth * (2H*W + W*H) * 2 = 29.0 GB at the default sizes.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

constexpr int kRfConsumerWarps = 8;
constexpr int kRfThreads = (kRfConsumerWarps + 1) * 32;  // + 1 producer warp
#ifndef MB_RF_KC
#define MB_RF_KC 1024  // (512 / 256 measured: 7.90 / 8.73 ms against 7.98 ms per 6-row sample, profiles/r02_rf_experiments.md)
#endif
constexpr int kRfKC = MB_RF_KC;                           // K elements per pipeline stage
constexpr int kRfStageBytes = 16 * kRfKC * 2;             // a full 16-row tile chunk: 32 KB
constexpr int kRfStages = 160 * 1024 / kRfStageBytes;     // ring slots (as many as fit next to the activation rows)
constexpr int kRfRowGroup = 3;                            // rows handled per register round of the row-wise parts
constexpr int kRfMaxRows = 6;                             // rows per launch: CFG rows x images generated together

struct RfFusedParams {
  const void* const* blocks;  // device table [depth][6]: w12 packed, b12, w3 packed, b3, ln weight, ln bias
  const __nv_bfloat16* in_w; const __nv_bfloat16* in_b;    // input_proj [W, C], [W]
  const __nv_bfloat16* fin_w; const __nv_bfloat16* fin_b;  // final linear [C, W], [C]
  const __nv_bfloat16* mod; int64_t ld_mod;                // adaLN modulations [steps * B, depth * 3W + 2W]
  float* x;                                                 // [B, C] fp32, in / out
  __nv_bfloat16* h; __nv_bfloat16* hid; __nv_bfloat16* v;  // global scratch [B, W], [B, H], [B, C]
  uint32_t* bar;                                            // grid barrier counter (zeroed before the launch)
  int B;         // rows of this launch = cfg_rows x independent samples (images); rows of one sample are adjacent
  int cfg_rows;  // rows per sample: 1 (no guidance), 2 (cond, uncond) or 3 (+ text-uncond)
  int nstages;   // ring slots in use (what fits next to the B activation rows)
  int W, H, C, depth, steps;
  float dt, text_cfg, image_cfg;
  unsigned long long* dbg;  // optional [2 CTAs][8] accumulated phase times in ns (mb_rf_set_debug), or null
};

__host__ __device__ inline int rf_unit_begin(int c, int R, int G) { return static_cast<int>((static_cast<int64_t>(c) * R) / G); }

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kRfConsumerWarps * 32) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// (not volatile: a pure function of its operands, so the compiler may schedule it around the shared loads)
__device__ __forceinline__ void mma16816_rf(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                            uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 16-byte shared load from a 32-bit shared-space address (no generic-address conversion in the hot loop); volatile keeps
// it behind the mbarrier wait that guards the stage it reads
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// Grid barrier of the consumer threads of all CTAs (all CTAs are resident: one per SM).  `target` = arrivals expected so
// far.  Bounded: a lost CTA traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target) {
  consumer_sync();  // every consumer thread's global stores are ordered before thread 0's release below
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_gpu(bar) - target) < 0) {
      if (++spins > (1u << 24)) {
        printf("rf_sample_fused: grid barrier timeout (block %d, target %u)\n", static_cast<int>(blockIdx.x), target);
        __trap();
      }
    }
  }
  consumer_sync();
}

__global__ void __launch_bounds__(kRfThreads, 1) rf_sample_fused_kernel(const RfFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // layout: ring [kRfStages][32 KB] | act [B][act_pitch] | red [2][8][4][32] f32 | small
  uint8_t* ring = smem;
  const int act_pitch = p.H * 2 + 64;  // bytes per activation row (H >= W); + 64: consecutive rows in other bank halves
  uint8_t* act = ring + p.nstages * kRfStageBytes;
  float* red = reinterpret_cast<float*>(act + p.B * act_pitch);
  float* small = red + 2 * kRfConsumerWarps * 4 * 32;       // [64]: block-reduction scratch
  float* xs = small + 64;                                    // [B][C] fp32 Euler state (replicated in every CTA)
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(xs + kRfMaxRows * 32);  // [B][C] bf16 copy
  __shared__ uint64_t full_bar[kRfStages], empty_bar[kRfStages];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, c = blockIdx.x;
  const int B = p.B, W = p.W, H = p.H, C = p.C;
  // unit ranges of this CTA: hidden columns of w12 (H units), rows of w3 / columns of h (W units)
  const int hu0 = rf_unit_begin(c, H, G), hu1 = rf_unit_begin(c + 1, H, G);
  const int wu0 = rf_unit_begin(c, W, G), wu1 = rf_unit_begin(c + 1, W, G);

  if (tid == 0) {
    for (int s = 0; s < kRfStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kRfConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kRfConsumerWarps) {
    // ===================== producer: streams every stage of the whole sample, in consumption order =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t round = 0;
      for (int step = 0; step < p.steps; ++step) {
        for (int blk = 0; blk < p.depth; ++blk) {
          const void* const* bp = p.blocks + blk * 6;
          for (int phase = 0; phase < 2; ++phase) {
            const int K = phase == 0 ? W : H;
            const int u0 = phase == 0 ? hu0 : wu0, u1 = phase == 0 ? hu1 : wu1;
            const int ut = phase == 0 ? 8 : 16, rows_per_unit = phase == 0 ? 2 : 1;
            const uint8_t* src = static_cast<const uint8_t*>(bp[phase == 0 ? 0 : 2]) +
                                 static_cast<int64_t>(u0) * rows_per_unit * K * 2;
            for (int t0 = u0; t0 < u1; t0 += ut) {
              const int nst = min(ut, u1 - t0) * rows_per_unit;  // storage rows of this tile
              const uint32_t bytes = static_cast<uint32_t>(nst) * kRfKC * 2;
              for (int kc = 0; kc < K; kc += kRfKC) {
                if (round > 0) mbar_wait(&empty_bar[slot], (round - 1) & 1);
                mbar_arrive_expect_tx(&full_bar[slot], bytes);
                bulk_load(ring + slot * kRfStageBytes, src, bytes, &full_bar[slot]);
                src += bytes;
                if (++slot == p.nstages) {
                  slot = 0;
                  ++round;
                }
              }
            }
          }
        }
      }
    }
    return;
  }

  // ================================================ consumers (256 threads) ================================================
  const int g = lane >> 2, t = lane & 3;
  int cslot = 0;            // ring slot / parity of the next stage, in lockstep with the producer
  uint32_t cpar = 0;
  uint32_t nbar = 0;        // grid barriers passed
  const uint32_t ring_s = smem_u32(ring);
  const uint32_t brow_s = smem_u32(act) + min(g, B - 1) * act_pitch + t * 16;  // this lane's activation row (tokens >= B
                                                                                // alias the last row: never stored)
  // debug aid: thread 0 of the first and the last CTA accumulate wall time per phase kind (ns, %globaltimer)
  const bool dbg_on = p.dbg != nullptr && tid == 0 && (c == 0 || c == G - 1);
  unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long dbg_t = 0;
  auto stamp = [&](int kind) {  // closes the interval since the previous stamp and books it under `kind`
    if (dbg_on) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (kind >= 0) dbg_acc[kind] += now - dbg_t;
      dbg_t = now;
    }
  };
  stamp(-1);
  auto block_sum3 = [&](float (&v)[kRfRowGroup]) {  // sums over the 256 consumer threads (fixed order), a row group at once
#pragma unroll
    for (int b = 0; b < kRfRowGroup; ++b) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[b] += __shfl_xor_sync(0xffffffffu, v[b], o);
    }
    consumer_sync();
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < kRfRowGroup; ++b) small[b * kRfConsumerWarps + warp] = v[b];
    }
    consumer_sync();
#pragma unroll
    for (int b = 0; b < kRfRowGroup; ++b) {
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kRfConsumerWarps; ++w2) s += small[b * kRfConsumerWarps + w2];
      v[b] = s;
    }
  };

  // Euler state
  for (int i = tid; i < B * C; i += kRfConsumerWarps * 32) {
    const float v = p.x[i];
    xs[i] = v;
    xb[i] = __float2bfloat16_rn(v);
  }
  consumer_sync();

  // Streams the stages of one layer phase of this CTA.  Per tile: `epi_preload` requests the per-output parameters of the
  // epilogue threads early, the 8 warps split every stage's 32 k groups, two accumulators halve the MMA dependency chain,
  // the warps' partial sums meet in shared memory (double-buffered by tile parity) and `epilogue` finishes the tile.
  // Rows / tokens a partial tile or B < 8 does not have alias valid ones (clamped addresses): their accumulator elements
  // are garbage that the epilogues never store — so the hot loop has no predicates.
  auto stream_phase = [&](int K, int u0, int u1, bool swiglu, auto&& epi_preload, auto&& epilogue) {
    const int ut = swiglu ? 8 : 16, rpu = swiglu ? 2 : 1;
    int tile_idx = 0;
    constexpr int kJ = kRfKC / 32 / kRfConsumerWarps;  // k groups per warp and chunk
    for (int t0 = u0; t0 < u1; t0 += ut, ++tile_idx) {
      const int nu = min(ut, u1 - t0), nst = nu * rpu;
      const int srow_lo = min(g, nu - 1);
      const int srow_hi = swiglu ? nu + min(g, nu - 1) : min(g + 8, nu - 1);
      const uint32_t off_lo = srow_lo * 64 + t * 16, off_hi = srow_hi * 64 + t * 16;
      const uint32_t kg_stride = nst * 64;  // bytes per k group in a chunk
      epi_preload(t0, nu);
      float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
      // one chunk through the shared-memory ring
      auto ring_chunk = [&](int kc) {
        const uint32_t bb = brow_s + kc * 2 + warp * 64;
        mbar_wait(&full_bar[cslot], cpar);
        const uint32_t sb = ring_s + cslot * kRfStageBytes + warp * kg_stride;
        uint4 alo[kJ], ahi[kJ], xv[kJ];
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
          alo[j] = lds128(sb + j * kRfConsumerWarps * kg_stride + off_lo);
          ahi[j] = lds128(sb + j * kRfConsumerWarps * kg_stride + off_hi);
          xv[j] = lds128(bb + j * kRfConsumerWarps * 64);
        }
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
          mma16816_rf(acc0, alo[j].x, ahi[j].x, alo[j].y, ahi[j].y, xv[j].x, xv[j].y);
          mma16816_rf(acc1, alo[j].z, ahi[j].z, alo[j].w, ahi[j].w, xv[j].z, xv[j].w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[cslot]);
        if (++cslot == p.nstages) {
          cslot = 0;
          cpar ^= 1;
        }
      };
      for (int kc = 0; kc < K; kc += kRfKC) ring_chunk(kc);
      // cross-warp reduction of the K slices (fixed order), double-buffered by tile parity
      float* rb = red + (tile_idx & 1) * kRfConsumerWarps * 4 * 32;
#pragma unroll
      for (int e = 0; e < 4; ++e) rb[(warp * 4 + e) * 32 + lane] = acc0[e] + acc1[e];
      consumer_sync();
      epilogue(t0, nu, rb);
    }
  };
  auto red_sum = [&](const float* rb, int e, int ln) {
    float s = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kRfConsumerWarps; ++w2) s += rb[(w2 * 4 + e) * 32 + ln];
    return s;
  };

  // adaLN parameters of the NEXT prologue (this thread's <= 2 chunks of gamma / beta and of every row's shift / scale).
  // They do not depend on the activations, so they are requested BEFORE the grid barrier that precedes the prologue: their
  // latency (queued behind the bulk weight stream) overlaps the barrier instead of sitting on the critical path after it.
  constexpr int kCh = 2;
  uint4 pre_gm[kCh], pre_bt[kCh], pre_sh[kRfRowGroup][kCh], pre_sc[kRfRowGroup][kCh];
  auto load_mod = [&](int step, int blk, int row0) {  // shift / scale of rows row0 .. row0 + 2 of (step, blk)
    const __nv_bfloat16* mb = p.mod + static_cast<int64_t>(step) * B * p.ld_mod + static_cast<int64_t>(blk) * 3 * W;
#pragma unroll
    for (int cc = 0; cc < kCh; ++cc) {
      const int i = tid + cc * kRfConsumerWarps * 32;
#pragma unroll
      for (int b = 0; b < kRfRowGroup; ++b) {
        const bool okb = i < W / 8 && row0 + b < B;
        const __nv_bfloat16* sh = mb + static_cast<int64_t>(row0 + b) * p.ld_mod;
        pre_sh[b][cc] = okb ? *reinterpret_cast<const uint4*>(sh + i * 8) : make_uint4(0, 0, 0, 0);
        pre_sc[b][cc] = okb ? *reinterpret_cast<const uint4*>(sh + W + i * 8) : make_uint4(0, 0, 0, 0);
      }
    }
  };
  auto preload = [&](int step, int blk) {
    const bool fin = blk == p.depth;
    if (fin && c >= C) return;
    const void* const* bp = p.blocks + (fin ? 0 : blk) * 6;
    const __nv_bfloat16* lnw = fin ? nullptr : static_cast<const __nv_bfloat16*>(bp[4]);
    const __nv_bfloat16* lnb = fin ? nullptr : static_cast<const __nv_bfloat16*>(bp[5]);
#pragma unroll
    for (int cc = 0; cc < kCh; ++cc) {
      const int i = tid + cc * kRfConsumerWarps * 32;
      const bool ok = i < W / 8;
      pre_gm[cc] = (ok && lnw) ? *reinterpret_cast<const uint4*>(lnw + i * 8) : make_uint4(0, 0, 0, 0);
      pre_bt[cc] = (ok && lnb) ? *reinterpret_cast<const uint4*>(lnb + i * 8) : make_uint4(0, 0, 0, 0);
    }
    load_mod(step, blk, 0);
  };

  for (int step = 0; step < p.steps; ++step) {
    // ---- input_proj (:371) on this CTA's columns of h:  h[b][n] = bf16(x_bf16[b] . in_w[n] + in_b[n])
    for (int i = tid; i < (wu1 - wu0) * B; i += kRfConsumerWarps * 32) {
      const int n = wu0 + i / B, b = i % B;
      float s = 0.f;
      for (int cc = 0; cc < C; ++cc)
        s += __bfloat162float(p.in_w[static_cast<int64_t>(n) * C + cc]) * __bfloat162float(xb[b * C + cc]);
      p.h[static_cast<int64_t>(b) * W + n] = __float2bfloat16_rn(s + __bfloat162float(p.in_b[n]));
    }
    preload(step, 0);
    stamp(5);
    grid_barrier(p.bar, ++nbar * G);
    stamp(6);

    const __nv_bfloat16* mod_step = p.mod + static_cast<int64_t>(step) * B * p.ld_mod;
    for (int blk = 0; blk <= p.depth; ++blk) {
      const bool final_layer = blk == p.depth;
      const void* const* bp = p.blocks + (final_layer ? 0 : blk) * 6;
      const __nv_bfloat16* lnw = final_layer ? nullptr : static_cast<const __nv_bfloat16*>(bp[4]);
      const __nv_bfloat16* lnb = final_layer ? nullptr : static_cast<const __nv_bfloat16*>(bp[5]);
      const __nv_bfloat16* mod_blk = mod_step + static_cast<int64_t>(blk) * 3 * W;  // shift | scale | gate (final: shift | scale)
      // ---- adaLN prologue (ResBlock :270 / FinalLayer :290): every CTA normalises the full rows of h into shared memory
      // (the final layer needs them only in the CTAs that compute an output channel)
      if (!final_layer || c < C) {
        // rows in groups of 3: each thread keeps its 16-byte chunks of the group's rows in registers (W / 8 <= 2 * 256
        // chunks per row); the first group's shift / scale were requested before the grid barrier, later groups' here
        auto unpack8 = [](const uint4& q, float (&f)[8]) {
          const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
          f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; f[4] = f2.x; f[5] = f2.y; f[6] = f3.x; f[7] = f3.y;
        };
        for (int row0 = 0; row0 < B; row0 += kRfRowGroup) {
          if (row0 > 0) load_mod(step, blk, row0);
          uint4 raw[kRfRowGroup][kCh];
          float sum[kRfRowGroup], sq[kRfRowGroup];
#pragma unroll
          for (int b = 0; b < kRfRowGroup; ++b) {
            sum[b] = 0.f;
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc) {
              const int i = tid + cc * kRfConsumerWarps * 32;
              raw[b][cc] = (row0 + b < B && i < W / 8)
                               ? __ldcg(reinterpret_cast<const uint4*>(p.h + static_cast<int64_t>(row0 + b) * W) + i)
                               : make_uint4(0, 0, 0, 0);
              float f[8];
              unpack8(raw[b][cc], f);
#pragma unroll
              for (int e = 0; e < 8; ++e) sum[b] += f[e];
            }
          }
          block_sum3(sum);
#pragma unroll
          for (int b = 0; b < kRfRowGroup; ++b) {
            sum[b] /= W;  // mean
            sq[b] = 0.f;
#pragma unroll
            for (int cc = 0; cc < kCh; ++cc) {
              if (tid + cc * kRfConsumerWarps * 32 < W / 8) {
                float f[8];
                unpack8(raw[b][cc], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float d = f[e] - sum[b];
                  sq[b] += d * d;
                }
              }
            }
          }
          block_sum3(sq);
#pragma unroll
          for (int cc = 0; cc < kCh; ++cc) {
            const int i = tid + cc * kRfConsumerWarps * 32;
            if (i >= W / 8) continue;
            __nv_bfloat16 gm[8], bt[8];
            *reinterpret_cast<uint4*>(gm) = pre_gm[cc];
            *reinterpret_cast<uint4*>(bt) = pre_bt[cc];
#pragma unroll
            for (int b = 0; b < kRfRowGroup; ++b) {
              if (row0 + b >= B) continue;
              const float mean = sum[b], rstd = rsqrtf(sq[b] / W + 1e-6f);
              __nv_bfloat16 shv[8], scv[8];
              *reinterpret_cast<uint4*>(shv) = pre_sh[b][cc];
              *reinterpret_cast<uint4*>(scv) = pre_sc[b][cc];
              float f[8], o[8];
              unpack8(raw[b][cc], f);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float xv = (f[e] - mean) * rstd;
                if (lnw) xv = xv * __bfloat162float(gm[e]) + (lnb ? __bfloat162float(bt[e]) : 0.f);
                o[e] = xv * bf16_round(1.f + __bfloat162float(scv[e])) + __bfloat162float(shv[e]);
              }
              uint4 o4;
              o4.x = pack_bf16x2(o[0], o[1]); o4.y = pack_bf16x2(o[2], o[3]);
              o4.z = pack_bf16x2(o[4], o[5]); o4.w = pack_bf16x2(o[6], o[7]);
              *reinterpret_cast<uint4*>(act + (row0 + b) * act_pitch + i * 16) = o4;
            }
          }
        }
      }
      consumer_sync();
      stamp(0);
      if (final_layer) {
        // ---- FinalLayer linear (:291): v[b][ch] = bf16(a[b] . fin_w[ch] + fin_b[ch]); CTA ch < C computes channel ch
        if (c < C) {
          for (int row0 = 0; row0 < B; row0 += kRfRowGroup) {
            float dot[kRfRowGroup] = {0.f, 0.f, 0.f};
            for (int i = tid; i < W / 8; i += kRfConsumerWarps * 32) {
              __nv_bfloat16 w8[8];
              *reinterpret_cast<uint4*>(w8) = *reinterpret_cast<const uint4*>(p.fin_w + static_cast<int64_t>(c) * W + i * 8);
#pragma unroll
              for (int b = 0; b < kRfRowGroup; ++b) {
                if (row0 + b >= B) continue;
                __nv_bfloat16 a8[8];
                *reinterpret_cast<uint4*>(a8) = *reinterpret_cast<const uint4*>(act + (row0 + b) * act_pitch + i * 16);
#pragma unroll
                for (int e = 0; e < 8; ++e) dot[b] += __bfloat162float(a8[e]) * __bfloat162float(w8[e]);
              }
            }
            block_sum3(dot);
            if (tid < kRfRowGroup && row0 + tid < B)
              p.v[(row0 + tid) * C + c] = __float2bfloat16_rn(dot[tid] + __bfloat162float(p.fin_b[c]));
          }
        }
        stamp(5);
        grid_barrier(p.bar, ++nbar * G);
        stamp(6);
        // ---- CFG combine + Euler (:145-179), replicated in every CTA on its shared-memory copy of x
        for (int i = tid; i < (B / p.cfg_rows) * C; i += kRfConsumerWarps * 32) {
          const int ch = i % C, r0 = (i / C) * p.cfg_rows;  // sample i / C owns rows r0 .. r0 + cfg_rows - 1
          auto vld = [&](int b) {  // written by other CTAs before the grid barrier: read through L2
            return __bfloat162float(
                __ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p.v) + (r0 + b) * C + ch)));
          };
          float stepv;  // every row of a sample receives the same update
          if (p.cfg_rows == 3) {
            const float vc = vld(0), vu = vld(1), vt = vld(2);
            const float t3 = bf16_round(vu + bf16_round(p.image_cfg * bf16_round(vt - vu)));
            const float vg = bf16_round(t3 + bf16_round(p.text_cfg * bf16_round(vc - vt)));
            stepv = bf16_round(vg * p.dt);
          } else if (p.cfg_rows == 2) {
            const float vc = vld(0), vu = vld(1);
            const float vg = bf16_round(vu + bf16_round(p.text_cfg * bf16_round(vc - vu)));
            stepv = bf16_round(vg * p.dt);
          } else {
            stepv = bf16_round(vld(0) * p.dt);
          }
          for (int b = 0; b < p.cfg_rows; ++b) {
            const float nx = xs[(r0 + b) * C + ch] + stepv;
            xs[(r0 + b) * C + ch] = nx;
            xb[(r0 + b) * C + ch] = __float2bfloat16_rn(nx);
          }
        }
        consumer_sync();
        stamp(5);
        break;
      }

      // ---- w12 + SwiGLU (:54-72 via :271): hid[b][u] = bf16(silu(bf16(a.Wg[u] + bg[u]))) * bf16(a.Wu[u] + bu[u])
      const __nv_bfloat16* b12 = static_cast<const __nv_bfloat16*>(bp[1]);
      float e_bg = 0.f, e_bu = 0.f;
      stream_phase(W, hu0, hu1, true,
                   [&](int t0, int nu) {
                     if (tid < 64) {
                       const int gg = (tid & 31) >> 2, tok = 2 * (tid & 3) + (tid >> 5);
                       if (gg < nu && tok < B) {
                         e_bg = __bfloat162float(b12[t0 + gg]);
                         e_bu = __bfloat162float(b12[H + t0 + gg]);
                       }
                     }
                   },
                   [&](int t0, int nu, const float* rb) {
                     if (tid < 64) {
                       const int ln = tid & 31, sel = tid >> 5;  // accumulator elements sel (gate), sel + 2 (up) of lane ln
                       const int gg = ln >> 2, tok = 2 * (ln & 3) + sel;
                       if (gg < nu && tok < B) {
                         const float x1 = bf16_round(red_sum(rb, sel, ln) + e_bg);
                         const float x2 = bf16_round(red_sum(rb, sel + 2, ln) + e_bu);
                         p.hid[static_cast<int64_t>(tok) * H + t0 + gg] = __float2bfloat16_rn(bf16_round(silu(x1)) * x2);
                       }
                     }
                   });
      stamp(1);
      grid_barrier(p.bar, ++nbar * G);
      stamp(2);

      // ---- w3 + gated residual (:272): h[b][n] = bf16(h[b][n] + bf16(gate[b][n] * bf16(hid[b] . W3[n] + b3[n])))
      for (int i = tid; i < B * (H / 8); i += kRfConsumerWarps * 32) {
        const int b = i / (H / 8), ch = i % (H / 8);
        *reinterpret_cast<uint4*>(act + b * act_pitch + ch * 16) =
            __ldcg(reinterpret_cast<const uint4*>(p.hid + static_cast<int64_t>(b) * H) + ch);
      }
      consumer_sync();
      const __nv_bfloat16* b3 = static_cast<const __nv_bfloat16*>(bp[3]);
      float e_b3 = 0.f, e_gate = 0.f, e_h = 0.f;
      stream_phase(H, wu0, wu1, false,
                   [&](int t0, int nu) {
                     if (tid < 128) {
                       const int ln = tid & 31, e = tid >> 5;
                       const int r = (ln >> 2) + 8 * (e >> 1), tok = 2 * (ln & 3) + (e & 1);
                       if (r < nu && tok < B) {
                         const int n = t0 + r;
                         e_b3 = __bfloat162float(b3[n]);
                         e_gate = __bfloat162float(mod_blk[static_cast<int64_t>(tok) * p.ld_mod + 2 * W + n]);
                         e_h = __bfloat162float(__ushort_as_bfloat16(
                             __ldcg(reinterpret_cast<const unsigned short*>(p.h) + static_cast<int64_t>(tok) * W + n)));
                       }
                     }
                   },
                   [&](int t0, int nu, const float* rb) {
                     if (tid < 128) {
                       const int ln = tid & 31, e = tid >> 5;
                       const int r = (ln >> 2) + 8 * (e >> 1), tok = 2 * (ln & 3) + (e & 1);
                       if (r < nu && tok < B) {
                         const float vv = bf16_round(red_sum(rb, e, ln) + e_b3);
                         const float gh = bf16_round(e_gate * vv);
                         p.h[static_cast<int64_t>(tok) * W + t0 + r] = __float2bfloat16_rn(e_h + gh);
                       }
                     }
                   });
      preload(step, blk + 1);
      stamp(3);
      grid_barrier(p.bar, ++nbar * G);
      stamp(4);
    }
  }
  if (dbg_on)
    for (int i = 0; i < 8; ++i) p.dbg[(c == 0 ? 0 : 8) + i] = dbg_acc[i];
  if (c == 0)
    for (int i = tid; i < B * C; i += kRfConsumerWarps * 32) p.x[i] = xs[i];
}

// W [N, K] (nn.Linear layout; swiglu: N = 2 * units, gate rows then up rows) -> the per-CTA stage order described above.
__global__ void __launch_bounds__(256) rf_pack_kernel(const __nv_bfloat16* __restrict__ Wsrc, __nv_bfloat16* __restrict__ out,
                                                      int units, int K, int swiglu) {
  const int G = gridDim.x, c = blockIdx.x;
  const int u0 = rf_unit_begin(c, units, G), u1 = rf_unit_begin(c + 1, units, G);
  const int ut = swiglu ? 8 : 16, rpu = swiglu ? 2 : 1;
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(u0) * rpu * K);
  for (int t0 = u0; t0 < u1; t0 += ut) {
    const int nu = min(ut, u1 - t0), nst = nu * rpu;
    const int chunks16 = nst * K / 8;  // 16-byte chunks of this tile
    for (int i = threadIdx.x; i < chunks16; i += blockDim.x) {
      // destination order: [K chunk][k group][storage row][4 x 16 B]
      const int per_kc = nst * kRfKC / 8;
      const int kc = i / per_kc, r1 = i % per_kc;
      const int kg = r1 / (nst * 4), r2 = r1 % (nst * 4);
      const int srow = r2 / 4, tq = r2 % 4;
      const int unit = t0 + (swiglu ? srow % nu : srow);
      const int64_t src_row = swiglu ? (srow < nu ? unit : units + unit) : unit;
      const int k = kc * kRfKC + kg * 32 + tq * 8;
      dst[i] = *reinterpret_cast<const uint4*>(Wsrc + src_row * K + k);
    }
    dst += chunks16;
  }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_rf_pack_weights(const void* W, int N, int K, int swiglu, int n_cta, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_rf_pack_weights: no sm_100 device");
  MB_CHECK_ARG(N >= 1 && K >= kRfKC && K % kRfKC == 0 && n_cta >= 1 && (!swiglu || N % 2 == 0), MB_ERR_SHAPE,
               "mb_rf_pack_weights: K must be a multiple of %d (N=%d K=%d)", kRfKC, N, K);
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, MB_ERR_ALIGN,
               "mb_rf_pack_weights: 16-byte aligned buffers");
  rf_pack_kernel<<<n_cta, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(W), static_cast<__nv_bfloat16*>(out),
                                            swiglu ? N / 2 : N, K, swiglu);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

static unsigned long long* g_rf_dbg = nullptr;
// Debug aid (tools/debug_rf_timing.py): device buffer of 16 u64 — per phase kind accumulated ns of the first and the last
// CTA: 0 adaLN prologue, 1 w12 stream + epilogue, 2 barrier after w12, 3 w3 stream + epilogue (incl. staging hid),
// 4 barrier after w3, 5 input_proj / final layer / Euler, 6 their barriers.  NULL switches it off.
extern "C" int mb_rf_set_debug(void* buf) {
  g_rf_dbg = static_cast<unsigned long long*>(buf);
  return MB_OK;
}

static int rf_stages_for(int B, int H) {  // ring slots that fit next to B activation rows (<= kRfStages)
  const long avail = 226L * 1024 - 10 * 1024 - static_cast<long>(B) * (H * 2 + 64);
  const long n = avail / kRfStageBytes;
  return n > kRfStages ? kRfStages : static_cast<int>(n);
}

extern "C" int mb_rf_fused_supported(int B, int W, int H, int C) {
  return (B >= 1 && B <= kRfMaxRows && W % kRfKC == 0 && H % kRfKC == 0 && H >= W && W / 8 <= 2 * kRfConsumerWarps * 32 &&
          C >= 1 && C <= 32 && rf_stages_for(B, H) >= 2)
             ? 1
             : 0;
}

extern "C" int mb_rf_sample_fused(const void* const* block_ptrs, const void* in_w, const void* in_b, const void* fin_w,
                                  const void* fin_b, const void* mod, int64_t ld_mod, float* x, void* h_scratch,
                                  void* hid_scratch, void* v_scratch, uint32_t* barrier, int B, int cfg_rows, int W,
                                  int H, int C, int depth, int steps, float text_cfg, float image_cfg, int n_cta,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_rf_sample_fused: no sm_100 device");
  MB_CHECK_ARG(mb_rf_fused_supported(B, W, H, C) && depth >= 1 && steps >= 1, MB_ERR_SHAPE,
               "mb_rf_sample_fused: unsupported shape (B=%d W=%d H=%d C=%d)", B, W, H, C);
  MB_CHECK_ARG(cfg_rows >= 1 && cfg_rows <= 3 && B % cfg_rows == 0, MB_ERR_SHAPE,
               "mb_rf_sample_fused: %d rows are not whole samples of %d CFG rows", B, cfg_rows);
  MB_CHECK_ARG(n_cta == num_sms(), MB_ERR_SHAPE, "mb_rf_sample_fused: weights packed for %d CTAs, device has %d SMs",
               n_cta, num_sms());
  MB_CHECK_ARG(ld_mod % 8 == 0 && (reinterpret_cast<uintptr_t>(mod) & 15) == 0, MB_ERR_ALIGN,
               "mb_rf_sample_fused: modulation rows must be 16-byte aligned");
  RfFusedParams p;
  p.blocks = block_ptrs;
  p.in_w = static_cast<const __nv_bfloat16*>(in_w); p.in_b = static_cast<const __nv_bfloat16*>(in_b);
  p.fin_w = static_cast<const __nv_bfloat16*>(fin_w); p.fin_b = static_cast<const __nv_bfloat16*>(fin_b);
  p.mod = static_cast<const __nv_bfloat16*>(mod); p.ld_mod = ld_mod;
  p.x = x;
  p.h = static_cast<__nv_bfloat16*>(h_scratch); p.hid = static_cast<__nv_bfloat16*>(hid_scratch);
  p.v = static_cast<__nv_bfloat16*>(v_scratch);
  p.bar = barrier;
  p.B = B; p.cfg_rows = cfg_rows; p.nstages = rf_stages_for(B, H);
  p.W = W; p.H = H; p.C = C; p.depth = depth; p.steps = steps;
  p.dt = 1.0f / steps; p.text_cfg = text_cfg; p.image_cfg = image_cfg;
  p.dbg = g_rf_dbg;
  const size_t smem = static_cast<size_t>(p.nstages) * kRfStageBytes + static_cast<size_t>(B) * (H * 2 + 64) +
                      2 * kRfConsumerWarps * 4 * 32 * 4 + 64 * 4 + kRfMaxRows * 32 * 4 + kRfMaxRows * 32 * 2 + 64;
  MB_CHECK_ARG(smem <= 226 * 1024, MB_ERR_SHAPE, "mb_rf_sample_fused: %zu bytes of shared memory", smem);
  static size_t attr_smem = 0;  // (the kernel also has ~1 KB of static shared memory: 227 KB is the sum's limit)
  if (attr_smem < smem) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(rf_sample_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    attr_smem = smem;
  }
  MB_CHECK_CUDA(cudaMemsetAsync(barrier, 0, sizeof(uint32_t), stream));
  rf_sample_fused_kernel<<<n_cta, kRfThreads, smem, stream>>>(p);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}
