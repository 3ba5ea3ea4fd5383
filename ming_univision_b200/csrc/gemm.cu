// Persistent warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] @ W[N,K]^T + bias).
//
//   warp 0        TMA producer   : cp.async.bulk.tensor 2-D loads of A (128 x 64) and W (BN x 64) tiles into a
//                                  kStages-deep ring of 128B-swizzled shared-memory buffers, mbarrier complete_tx.
//   warp 1        MMA issuer     : one lane issues tcgen05.mma (M=128, N=BN, K=16) with both operands read from
//                                  shared memory through UMMA descriptors; accumulators live in TMEM, double
//                                  buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of i+1.
//   warps 2..9    epilogue       : BEFORE the accumulator is complete: the tile's per-column vectors (bias, or the
//                                  folded LayerNorm bias and column sums) go to shared memory and every residual box
//                                  of the tile is requested by TMA into its own staging box; then tcgen05.ld 32 lanes
//                                  x 32 columns -> registers -> bias / GELU / SwiGLU / residual (packed FFMA2 math) ->
//                                  bf16 -> 128B-swizzled staging box -> TMA store.  (Direct 16-byte stores remain for
//                                  the remapped outputs of the patch embedding.)
//
// Grouped mode (mb_moe_grouped_gemm): the same kernel walks expert-sorted rows whose tile -> expert map and tile COUNT
// live in device memory (routing plan), offsets the B tile by expert, and can assemble it from two half-height TMA
// boxes (gate rows, up rows) for the SwiGLU epilogue.
//
// CG = 2 runs the same roles on a CTA PAIR (cluster of 2, tcgen05 cta_group::2): one 256 x BN tile per pair, each CTA
// stages its own 128 rows of A and HALF of the W tile (BN/2 rows), the leader CTA issues M = 256 MMAs that read both
// CTAs' shared memory, and each CTA drains its own 128 accumulator rows from its TMEM.  Per CTA this loads
// 32 KB instead of 48 KB per 128x256x64 MACs (1.5x less L2 -> SM traffic, half the B shared-memory reads).
//
// Both operands are K-major (nn.Linear weight layout), so no transposes are needed anywhere.
// Tails in M, N and K are handled by TMA zero fill on loads and by predication on stores.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = (2 + kNumEpiWarps) * 32;

struct GemmParams {
  int M, N, K;
  const __nv_bfloat16* bias;
  __nv_bfloat16* out;
  int64_t ldo;
  const __nv_bfloat16* residual;
  int64_t ldr;
  int res_row_mod;
  int out_row_group;
  int out_row_pad;
  int tma_epi;  // 1: epilogue goes through shared-memory staging + TMA store (and TMA load of the residual)
  // LayerNorm folded into this GEMM (see mb_gemm_bf16_ex): A holds the UN-normalised rows, W already carries gamma,
  //   out = rstd[r] * (acc - mean[r] * csum[n]) + bias_f32[n]
  const float* ln_stats_in;   // [M][ln_slots_in][2] per-row partial (sum, sum of squares) of A, or NULL
  int ln_slots_in;
  const float* ln_csum;       // [N]    row sums of the gamma-scaled weight (fp32)
  const float* ln_bias;       // [N]    bias + W . beta (fp32)
  float ln_inv_dim, ln_eps;
  float* ln_stats_out;        // [M][ceil(N/64)][2] RESIDUAL epilogue: (sum, sumsq) of every 64-column box it writes
  int ln_slots_out;           //   (one slot per box, each written exactly once: deterministic, no atomics)
  // Grouped (per-expert) mode, mb_moe_grouped_gemm: A rows are expert-sorted and every expert's segment is padded to
  // a multiple of 128 rows, so each 128-row M tile belongs to ONE expert; W is [E][N][K] and the tile's B rows start at
  // expert * N.  The number of M tiles is data dependent and is read from device memory (no host sync).
  const int32_t* grp_tile_expert;  // [max M tiles] expert of every M tile, or NULL (dense GEMM)
  const int32_t* grp_num_m_tiles;  // device scalar
  int grp_w_rows;             // rows of W per expert (N, or 2I for the [gate; up] slab)
  int grp_n_out;              // output width (I or N)
  long long* dbg;             // optional clock64() stamps of worker 0 (mb_gemm_set_debug; development only)
  int grp_split;              // > 0: the B tile is two (BN/2)-row boxes, rows n_tile*BN/2 and grp_split + n_tile*BN/2 of
                              //      the expert's [gate; up] slab, so the SwiGLU epilogue needs no interleaved repack
};

template <int BN, int CG, int EPI>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;           // this CTA's 128 rows of A
  static constexpr int kBRows = BN / CG;                  // this CTA's share of the W tile
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue staging: one 32-row x 64-col bf16 box (4 KB) per epilogue warp; the RESIDUAL epilogue keeps one box per
  // 64-column slice of the warp's share so that every residual box of a tile is requested by TMA before the tile's
  // accumulator is even complete (the load latency hides behind the main loop instead of in front of the math)
  static constexpr int kOutTileN = (EPI == MB_EPI_SWIGLU) ? BN / 2 : BN;
  static constexpr int kBoxesPerWarp = kOutTileN / 2 / 64;
  static constexpr int kEpiBufs = (EPI == MB_EPI_RESIDUAL) ? kBoxesPerWarp : 1;
  static constexpr int kEpiStageBytes = kNumEpiWarps * kEpiBufs * 4096;
  static constexpr int kAuxBytes = 2 * BN * 4;            // per-column fp32 vectors of the current tile (bias, csum)
  static constexpr int kBarBytes = 512;
  static constexpr int kStages = ((232448 - kEpiStageBytes - kAuxBytes - kBarBytes) / kStageBytes) > 8
                                     ? 8 : ((232448 - kEpiStageBytes - kAuxBytes - kBarBytes) / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + kAuxBytes + kBarBytes;
  static constexpr int kTileM = kBM * CG;
  static_assert(kSmemBytes <= 232448 && kStages >= 3, "shared memory budget");
  static_assert((2 * kStages + 4 + kNumEpiWarps * 2) * 8 + 8 <= kBarBytes, "barrier area");
};

// Converts 32 fp32 values (one row segment) to bf16 and stores them; handles the N tail.
__device__ __forceinline__ void store_row_segment(__nv_bfloat16* dst, const float (&v)[32], int ncols_valid) {
  if (ncols_valid >= 32) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 q;
      q.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      q.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      q.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      q.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      d4[i] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols_valid) dst[i] = __float2bfloat16_rn(v[i]);
  }
}

// Accumulator chunk (this warp's 32 rows x 32 columns starting at tile column tc) -> bias / activation, fp32 in v[].
// The per-column vectors of the tile live in shared memory (aux_b: bias or the folded LayerNorm bias, aux_c: column
// sums of the gamma-scaled weight), staged before the accumulator wait, zero past the matrix edge; every lane reads the
// same addresses (broadcast LDS.128), so they cost no global-memory latency inside the epilogue.
//   plain:  v = acc + aux_b[col]            LayerNorm fold:  v = rstd * (acc - mean * aux_c[col]) + aux_b[col]
// For MB_EPI_RESIDUAL the residual is added by the caller.
__device__ __forceinline__ void col_affine32(const float* aux_b, const float* aux_c, bool fold, int tcol, float mean,
                                             float rstd, const uint32_t (&a)[32], float (&v)[32]) {
  const float4* bs = reinterpret_cast<const float4*>(aux_b + tcol);
  if (fold) {
    const float4* cs = reinterpret_cast<const float4*>(aux_c + tcol);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c = cs[i], b = bs[i];
      v[4 * i + 0] = fmaf(rstd, __uint_as_float(a[4 * i + 0]) - mean * c.x, b.x);
      v[4 * i + 1] = fmaf(rstd, __uint_as_float(a[4 * i + 1]) - mean * c.y, b.y);
      v[4 * i + 2] = fmaf(rstd, __uint_as_float(a[4 * i + 2]) - mean * c.z, b.z);
      v[4 * i + 3] = fmaf(rstd, __uint_as_float(a[4 * i + 3]) - mean * c.w, b.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = bs[i];
      v[4 * i + 0] = __uint_as_float(a[4 * i + 0]) + b.x;
      v[4 * i + 1] = __uint_as_float(a[4 * i + 1]) + b.y;
      v[4 * i + 2] = __uint_as_float(a[4 * i + 2]) + b.z;
      v[4 * i + 3] = __uint_as_float(a[4 * i + 3]) + b.w;
    }
  }
}

template <int BN, int EPI>
__device__ __forceinline__ void epi_math(uint32_t t_row, int tc, const float* aux_b, const float* aux_c, bool fold,
                                         float ln_mean, float ln_rstd, float (&v)[32]) {
  if constexpr (EPI == MB_EPI_SWIGLU) {
    uint32_t g[32], u[32];
    tmem_ld_32x32b_x32(t_row + tc, g);
    tmem_ld_32x32b_x32(t_row + BN / 2 + tc, u);
    tmem_ld_wait();
    float x1[32];
    col_affine32(aux_b, aux_c, fold, tc, ln_mean, ln_rstd, g, x1);
    col_affine32(aux_b, aux_c, fold, BN / 2 + tc, ln_mean, ln_rstd, u, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(silu(bf16_round(x1[i]))) * bf16_round(v[i]);
  } else {
    uint32_t a[32];
    tmem_ld_32x32b_x32(t_row + tc, a);
    tmem_ld_wait();
    col_affine32(aux_b, aux_c, fold, tc, ln_mean, ln_rstd, a, v);
    if constexpr (EPI == MB_EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {  // bf16(acc + bias) as the reference's Linear output, then GELU on packed pairs
        bf16x2_to_f32(pack_bf16x2(v[2 * i], v[2 * i + 1]), v[2 * i], v[2 * i + 1]);
        gelu_erf_x2(v[2 * i], v[2 * i + 1]);
      }
    }
  }
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, CG, EPI>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileM = Cfg::kTileM;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  // one "worker" = a CTA (CG = 1) or a CTA pair (CG = 2)
  const int worker = blockIdx.x / CG;
  const int num_workers = gridDim.x / CG;

  extern __shared__ __align__(1024) uint8_t smem[];  // (no static shared memory in this kernel: the base is 1 KB aligned)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_epi = smem + kStages * Cfg::kStageBytes;  // [kNumEpiWarps][kEpiBufs][32 rows][128 B], 128B-swizzled
  float* aux_b = reinterpret_cast<float*>(smem_epi + Cfg::kEpiStageBytes);  // [BN] bias / folded LayerNorm bias
  float* aux_c = aux_b + BN;                                                // [BN] column sums (LayerNorm fold)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::kEpiStageBytes + Cfg::kAuxBytes);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;     // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]      epilogue -> MMA
  uint64_t* epi_bar = bars + 2 * kStages + 4;   // [kNumEpiWarps][2] residual box landed (TMA -> epilogue warp)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + 2 * kNumEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const bool grouped = p.grp_tile_expert != nullptr;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_k_blocks = (p.K + kBK - 1) / kBK;
  // dense: from the shape; grouped: produced on the device by the routing plan -> read it after pdl_wait()
  auto tile_count = [&]() -> int {
    const int m_tiles = grouped ? __ldg(p.grp_num_m_tiles) : (p.M + kTileM - 1) / kTileM;
    return m_tiles * num_n_tiles;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kNumEpiWarps * CG);  // CG = 2: the leader's barrier collects both CTAs' epilogues
    }
    for (int w = 0; w < 2 * kNumEpiWarps; ++w) mbar_init(&epi_bar[w], 1);
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("gemm_bf16_kernel: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    if (p.tma_epi) {
      tma_prefetch_desc(&tmap_out);
      if (EPI == MB_EPI_RESIDUAL) tma_prefetch_desc(&tmap_res);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_base_slot, Cfg::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_base_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the predecessor's tail; from
  // here on the roles touch global memory, each after its own pdl_wait()
  pdl_launch_dependents();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      pdl_wait();
      const int num_tiles = tile_count();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int m_tile = tile / num_n_tiles;
        const int n_tile = tile % num_n_tiles;
        const int a_row = m_tile * kTileM + static_cast<int>(cta_rank) * kBM;
        int b_row = n_tile * BN + static_cast<int>(cta_rank) * Cfg::kBRows;
        if (grouped) {
          const int e = __ldg(p.grp_tile_expert + m_tile);
          b_row = e * p.grp_w_rows + (p.grp_split > 0 ? n_tile * (BN / 2) : n_tile * BN);
        }
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (CG == 2) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_cg2(&tmap_a, leader_full, smem_a + stage * Cfg::kABytes, kb * kBK, a_row);
            tma_load_2d_cg2(&tmap_b, leader_full, smem_b + stage * Cfg::kBBytes, kb * kBK, b_row);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::kABytes, kb * kBK, a_row);
            tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes, kb * kBK, b_row);
            if (p.grp_split > 0)  // second half of the B tile: the matching `up` rows (tmap_b box = BN/2 rows)
              tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::kBBytes + Cfg::kBBytes / 2, kb * kBK,
                          b_row + p.grp_split);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0 && is_leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (grouped) pdl_wait();
      const int num_tiles = tile_count();
      // stage 0 descriptors, split into (low word, shared high word): stages and K steps are 32-bit adds on the low word
      const uint64_t da_base = umma_desc_sw128_kmajor(smem_u32(smem_a));
      const uint64_t db_base = umma_desc_sw128_kmajor(smem_u32(smem_b));
      const uint32_t desc_hi = static_cast<uint32_t>(da_base >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(da_base), b_lo0 = static_cast<uint32_t>(db_base);
      int tcount = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++tcount) {
        const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && tcount < 16;
        if (stamp) p.dbg[tcount * 8 + 0] = clock64();
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (stamp) p.dbg[tcount * 8 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (stamp && kb == 0) p.dbg[tcount * 8 + 2] = clock64();
          const uint32_t a_lo = a_lo0 + stage * (Cfg::kABytes >> 4);
          const uint32_t b_lo = b_lo0 + stage * (Cfg::kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            if constexpr (CG == 2) umma_bf16_cg2_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (k != 0) ? 1u : static_cast<uint32_t>(kb));
            else umma_bf16_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (k != 0) ? 1u : static_cast<uint32_t>(kb));
          }
          // frees this smem slot (in both CTAs of a pair) once the MMAs above have read it
          if constexpr (CG == 2) umma_commit_cg2_mc(&empty_bar[stage], 0x3); else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (CG == 2) umma_commit_cg2_mc(&tmem_full[acc], 0x3); else umma_commit(&tmem_full[acc]);
        if (stamp) p.dbg[tcount * 8 + 3] = clock64();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = ew >> 2;           // which half of the tile's columns
    constexpr int kOutTileN = Cfg::kOutTileN;
    constexpr int kColsPerWarp = kOutTileN / 2;
    const bool fold = p.ln_stats_in != nullptr;
    const int n_out_total = grouped ? p.grp_n_out : ((EPI == MB_EPI_SWIGLU) ? p.N / 2 : p.N);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t epi_phase = 0;
    pdl_wait();  // the epilogue reads bias / residual / statistics and overwrites buffers earlier kernels may still read
    const int num_tiles = tile_count();
    int tcount = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++tcount) {
      const int m_tile = tile / num_n_tiles;
      const int n_tile = tile % num_n_tiles;
      const bool stamp = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 64 && tcount < 16;
      if (stamp) p.dbg[tcount * 8 + 4] = clock64();
      // ---- before the accumulator is complete (this overlaps the tile's main loop):
      // (1) the per-column vectors of the tile go to shared memory; (2) RESIDUAL: every residual box of the tile is
      // requested by TMA into its own staging box
      const int row0 = m_tile * kTileM + static_cast<int>(cta_rank) * kBM + quad * 32;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // every epilogue warp is done with the previous tile's vectors
      {
        const int e = static_cast<int>(threadIdx.x) - 64;
        if (e < BN) {
          const int col = n_tile * BN + e;
          float bv = 0.f, cv = 0.f;
          if (col < p.N) {
            if (fold) {
              bv = __ldg(p.ln_bias + col);
              cv = __ldg(p.ln_csum + col);
            } else if (p.bias != nullptr) {
              bv = __bfloat162float(p.bias[col]);
            }
          }
          aux_b[e] = bv;
          aux_c[e] = cv;
        }
      }
      if constexpr (EPI == MB_EPI_RESIDUAL) {
        if (p.tma_epi && lane == 0) {
          tma_store_wait_read<0>();  // the stores of the previous tile have read this warp's staging boxes
#pragma unroll
          for (int bx = 0; bx < Cfg::kBoxesPerWarp; ++bx) {
            const int box_col = n_tile * kOutTileN + half * kColsPerWarp + bx * 64;
            if (box_col < n_out_total && row0 < p.M) {
              mbar_arrive_expect_tx(&epi_bar[ew * 2 + bx], 4096);
              tma_load_2d(&tmap_res, &epi_bar[ew * 2 + bx], smem_epi + (ew * Cfg::kEpiBufs + bx) * 4096, box_col, row0);
            }
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // vectors visible to all epilogue warps
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (stamp) p.dbg[tcount * 8 + 5] = clock64();
      const int row = m_tile * kTileM + static_cast<int>(cta_rank) * kBM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      int64_t out_row = row;
      if (p.out_row_group > 0) out_row += static_cast<int64_t>(row / p.out_row_group) * p.out_row_pad;
      const int64_t res_row = (p.res_row_mod > 0) ? (row % p.res_row_mod) : row;
      const uint32_t t_row = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (p.ln_stats_in != nullptr && row_ok) {
        const float2* sp = reinterpret_cast<const float2*>(p.ln_stats_in) + static_cast<int64_t>(row) * p.ln_slots_in;
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < p.ln_slots_in; ++i) {  // fixed order -> bitwise reproducible
          const float2 st = __ldg(sp + i);
          s1 += st.x;
          s2 += st.y;
        }
        ln_mean = s1 * p.ln_inv_dim;
        ln_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_dim - ln_mean * ln_mean, 0.f) + p.ln_eps);
      }

      if (p.tma_epi) {
        // ---- staged path: registers -> 128B-swizzled smem box (32 rows x 64 cols) -> one TMA store per box; the
        // residual box arrives by TMA as well.  Every global access of the epilogue is a full-line bulk transfer
        // (a thread-per-row register store touches 32 different lines per instruction and was the limiter at K<=1024).
#pragma unroll 1
        for (int bx = 0; bx < kColsPerWarp / 64; ++bx) {
          uint8_t* stage_buf = smem_epi + (ew * Cfg::kEpiBufs + (Cfg::kEpiBufs > 1 ? bx : 0)) * 4096;
          const int box_tc = half * kColsPerWarp + bx * 64;   // first column of the box inside the output tile
          const int box_col = n_tile * kOutTileN + box_tc;    // global output column
          if (box_col >= n_out_total || row0 >= p.M) continue;  // warp-uniform
          float st_sum = 0.f, st_sq = 0.f;
          if constexpr (EPI != MB_EPI_RESIDUAL) {
            if (lane == 0) tma_store_wait_read<0>();          // previous box fully read out of the (single) staging box
            __syncwarp();
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int tc = box_tc + cc * 32;
            float v[32];
            epi_math<BN, EPI>(t_row, tc, aux_b, aux_c, fold, ln_mean, ln_rstd, v);
            uint32_t packed[16];  // the 32 output values of this chunk as bf16 pairs
            if constexpr (EPI == MB_EPI_RESIDUAL) {
              if (cc == 0) mbar_wait(&epi_bar[ew * 2 + bx], epi_phase);
              // out = bf16( bf16(acc + bias) + residual ), statistics of the stored values: all on packed fp32 pairs
              // (6 issue slots per element instead of 11 — this epilogue outlasted the K = 1024 main loop)
              const bool want_stats = p.ln_stats_out != nullptr;
              const int nv = n_out_total - (n_tile * kOutTileN + tc);
              uint64_t s2 = pack_f32x2(0.f, 0.f), q2 = s2;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(stage_buf + lane * 128 + (((cc * 4 + j) ^ (lane & 7)) << 4));
                const uint32_t rr[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int h2 = 0; h2 < 4; ++h2) {
                  const int i = 8 * j + 2 * h2;
                  float a0, a1, r0, r1;
                  bf16x2_to_f32(pack_bf16x2(v[i], v[i + 1]), a0, a1);
                  bf16x2_to_f32(rr[h2], r0, r1);
                  float o0, o1;
                  unpack_f32x2(add_f32x2(pack_f32x2(a0, a1), pack_f32x2(r0, r1)), o0, o1);
                  const uint32_t pk = pack_bf16x2(o0, o1);
                  packed[4 * j + h2] = pk;
                  if (want_stats) {
                    float b0, b1;
                    bf16x2_to_f32(pk, b0, b1);
                    if (i >= nv) b0 = 0.f;
                    if (i + 1 >= nv) b1 = 0.f;
                    const uint64_t b = pack_f32x2(b0, b1);
                    s2 = add_f32x2(s2, b);
                    q2 = fma_f32x2(b, b, q2);
                  }
                }
              }
              if (want_stats) {
                float a, b;
                unpack_f32x2(s2, a, b);
                st_sum += a + b;
                unpack_f32x2(q2, a, b);
                st_sq += a + b;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              *reinterpret_cast<uint4*>(stage_buf + lane * 128 + (((cc * 4 + j) ^ (lane & 7)) << 4)) =
                  make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
          }
          if constexpr (EPI == MB_EPI_RESIDUAL) {
            if (p.ln_stats_out != nullptr && row_ok)
              reinterpret_cast<float2*>(p.ln_stats_out)[static_cast<int64_t>(row) * p.ln_slots_out + (box_col >> 6)] =
                  make_float2(st_sum, st_sq);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, stage_buf, box_col, row0);
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp; c += 32) {
          const int tc = half * kColsPerWarp + c;       // column inside the output tile
          const int out_col = n_tile * kOutTileN + tc;  // global output column
          const int ncols_valid = n_out_total - out_col;
          float v[32];
          epi_math<BN, EPI>(t_row, tc, aux_b, aux_c, fold, ln_mean, ln_rstd, v);
          if constexpr (EPI == MB_EPI_RESIDUAL) {
            if (row_ok && ncols_valid > 0) {
              const __nv_bfloat16* rp = p.residual + res_row * p.ldr + out_col;
              if (ncols_valid >= 32) {
                const uint4* r4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint4 q = __ldg(r4 + i);
                  const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z),
                               f3 = unpack_bf16x2(q.w);
                  v[8 * i + 0] = bf16_round(v[8 * i + 0]) + f0.x; v[8 * i + 1] = bf16_round(v[8 * i + 1]) + f0.y;
                  v[8 * i + 2] = bf16_round(v[8 * i + 2]) + f1.x; v[8 * i + 3] = bf16_round(v[8 * i + 3]) + f1.y;
                  v[8 * i + 4] = bf16_round(v[8 * i + 4]) + f2.x; v[8 * i + 5] = bf16_round(v[8 * i + 5]) + f2.y;
                  v[8 * i + 6] = bf16_round(v[8 * i + 6]) + f3.x; v[8 * i + 7] = bf16_round(v[8 * i + 7]) + f3.y;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < ncols_valid) v[i] = bf16_round(v[i]) + __bfloat162float(rp[i]);
              }
            }
          }
          if (row_ok && ncols_valid > 0) store_row_segment(p.out + out_row * p.ldo + out_col, v, ncols_valid);
        }
      }
      if (p.tma_epi) epi_phase ^= 1;  // residual boxes: one barrier phase per tile
      // all TMEM reads of this warp are complete (tmem_ld_wait above) -> hand the accumulator back
      if (stamp) p.dbg[tcount * 8 + 6] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_epi && lane == 0) tma_store_wait_all();  // bulk stores still read our shared memory until they complete
    (void)epi_phase;
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // the peer may still signal our barriers until here
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int EPI, int CG>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
                       const GemmParams& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::kSmemBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  MB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, EPI, CG>, ta, tb, to, tr, p));
  return MB_OK;
}

// Tile shape choice.  A "worker" is a CTA (CG = 1, 128 x BN tile) or a CTA pair (CG = 2, 256 x BN tile); the model is
// waves x per-tile MMA time x a measured penalty for the L2 -> SM traffic of the narrower shapes (bench_ops.py).
struct TileChoice { int cg, bn; };
static long long* g_gemm_dbg = nullptr;
static int g_force_cg = getenv("MB_GEMM_CG") ? atoi(getenv("MB_GEMM_CG")) : 0;
static int g_force_bn = getenv("MB_GEMM_BN") ? atoi(getenv("MB_GEMM_BN")) : 0;
static int g_no_tma_epi = getenv("MB_GEMM_NO_TMA_EPI") ? atoi(getenv("MB_GEMM_NO_TMA_EPI")) : 0;
static TileChoice choose_tile(int M, int N, int sms, bool swiglu) {
  const int force_cg = g_force_cg, force_bn = g_force_bn;
  // Measured on B200 (tools/bench_ops.py, gpurun_out/bench_ops2.log): every shape is bound by L2 -> SM operand
  // traffic (~12-13 TB/s), so the 256-wide tiles always win per wave; the pair kernel moves 1.5x fewer bytes per
  // MAC than the single-CTA one but quantises M to 256.  The narrow tiles only pay off when they save whole waves.
  const TileChoice cands[4] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}};
  const double penalty[4] = {1.0, 1.7, 1.06, 1.7};
  TileChoice best = {2, 256};
  double best_cost = 1e30;
  for (int i = 0; i < 4; ++i) {
    const TileChoice c = cands[i];
    if (swiglu && c.bn != 256) continue;
    if (force_cg && c.cg != force_cg) continue;
    if (force_bn && c.bn != force_bn) continue;
    const long tiles = static_cast<long>((M + 128 * c.cg - 1) / (128 * c.cg)) * ((N + c.bn - 1) / c.bn);
    const long workers = sms / c.cg;
    const long waves = (tiles + workers - 1) / workers;
    const double cost = static_cast<double>(waves) * c.bn * penalty[i];
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

}  // namespace mb

using namespace mb;

// Development aid: device buffer of >= 16 * 8 int64 receiving clock64() stamps of the first CTA's MMA / epilogue roles.
extern "C" int mb_gemm_set_debug(void* dev_buf) {
  mb::g_gemm_dbg = static_cast<long long*>(dev_buf);
  return MB_OK;
}

extern "C" int mb_gemm_force_tile(int cta_group, int bn) {
  MB_CHECK_ARG(((cta_group & 3) <= 2) && (cta_group & ~0x13) == 0 && (bn == 0 || bn == 128 || bn == 256), MB_ERR_SHAPE,
               "mb_gemm_force_tile: cta_group in {0,1,2}, bn in {0,128,256}");
  mb::g_force_cg = cta_group & 3;
  mb::g_force_bn = bn;
  mb::g_no_tma_epi = (cta_group >> 4) & 1;  // bit 4: force the direct-store epilogue (tests / A-B measurements)
  return MB_OK;
}

extern "C" int mb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                            int64_t ldo, int M, int N, int K, int epi, const void* residual, int64_t ldr,
                            int res_row_mod, int out_row_group, int out_row_pad, void* stream_) {
  return mb_gemm_bf16_ex(A, lda, W, ldw, bias, out, ldo, M, N, K, epi, residual, ldr, res_row_mod, out_row_group,
                         out_row_pad, nullptr, 0, nullptr, nullptr, 0.f, nullptr, stream_);
}

extern "C" int mb_gemm_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* out,
                               int64_t ldo, int M, int N, int K, int epi, const void* residual, int64_t ldr,
                               int res_row_mod, int out_row_group, int out_row_pad, const float* ln_stats_in,
                               int ln_slots_in, const float* ln_csum, const float* ln_bias_f32, float ln_eps,
                               float* ln_stats_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_gemm_bf16: no sm_100 device");
  MB_CHECK_ARG(M >= 0 && N >= 1 && K >= 1, MB_ERR_SHAPE, "mb_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
  if (M == 0) return MB_OK;
  MB_CHECK_ARG(epi >= 0 && epi <= 3, MB_ERR_SHAPE, "mb_gemm_bf16: unknown epilogue %d", epi);
  MB_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0, MB_ERR_ALIGN,
               "mb_gemm_bf16: K/lda/ldw/ldo must be multiples of 8 (K=%d lda=%ld ldw=%ld ldo=%ld)", K, (long)lda,
               (long)ldw, (long)ldo);
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               MB_ERR_ALIGN, "mb_gemm_bf16: A/W/out/bias must be 16-byte aligned");
  if (epi == MB_EPI_SWIGLU)
    MB_CHECK_ARG(N % 256 == 0, MB_ERR_SHAPE, "mb_gemm_bf16: SWIGLU needs packed N %% 256 == 0 (N=%d)", N);
  if (epi == MB_EPI_RESIDUAL)
    MB_CHECK_ARG(residual != nullptr && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
                 MB_ERR_ALIGN, "mb_gemm_bf16: RESIDUAL needs a 16-byte aligned residual with ldr %% 8 == 0");

  const int sms = mb::num_sms();
  const TileChoice tc = choose_tile(M, N, sms, epi == MB_EPI_SWIGLU);
  const int bn = tc.bn, cg = tc.cg;

  CUtensorMap ta, tb;
  if (!make_tmap_2d_bf16(&ta, A, K, M, lda, kBK, kBM)) return MB_ERR_CUDA;
  if (!make_tmap_2d_bf16(&tb, W, K, N, ldw, kBK, bn / cg)) return MB_ERR_CUDA;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.res_row_mod = res_row_mod;
  p.out_row_group = out_row_group;
  p.out_row_pad = out_row_pad;

  // Staged TMA epilogue whenever the output is a plain dense matrix; the remapped / row-modulo cases (patch-embed)
  // and very narrow outputs keep the direct register stores.
  const int n_out = (epi == MB_EPI_SWIGLU) ? N / 2 : N;
  const bool tma_epi = !g_no_tma_epi && out_row_group == 0 && res_row_mod == 0 && n_out >= 64 &&
                       (reinterpret_cast<uintptr_t>(residual) & 15) == 0;
  p.tma_epi = tma_epi ? 1 : 0;
  p.ln_stats_in = ln_stats_in;
  p.ln_slots_in = ln_slots_in;
  p.ln_slots_out = (n_out + 63) / 64;
  p.ln_csum = ln_csum;
  p.ln_bias = ln_bias_f32;
  p.ln_inv_dim = 1.0f / static_cast<float>(K);
  p.ln_eps = ln_eps;
  p.ln_stats_out = ln_stats_out;
  p.grp_tile_expert = nullptr;
  p.grp_num_m_tiles = nullptr;
  p.dbg = g_gemm_dbg;
  p.grp_split = 0;
  p.grp_w_rows = 0;
  p.grp_n_out = 0;
  if (ln_stats_in != nullptr)
    MB_CHECK_ARG(ln_csum != nullptr && ln_bias_f32 != nullptr && epi != MB_EPI_RESIDUAL && ln_slots_in >= 1 &&
                     (reinterpret_cast<uintptr_t>(ln_csum) & 15) == 0 && (reinterpret_cast<uintptr_t>(ln_bias_f32) & 15) == 0,
                 MB_ERR_SHAPE, "mb_gemm_bf16_ex: LayerNorm fold needs 16-byte aligned csum / bias_f32 and a non-residual epilogue");
  if (ln_stats_out != nullptr) {
    MB_CHECK_ARG(epi == MB_EPI_RESIDUAL && tma_epi, MB_ERR_SHAPE,
                 "mb_gemm_bf16_ex: row statistics are produced by the dense RESIDUAL epilogue only");
  }
  CUtensorMap to = ta, tr = ta;  // valid placeholders when unused
  if (tma_epi) {
    if (!make_tmap_2d_bf16(&to, out, n_out, M, ldo, 64, 32)) return MB_ERR_CUDA;
    if (epi == MB_EPI_RESIDUAL && !make_tmap_2d_bf16(&tr, residual, n_out, M, ldr, 64, 32)) return MB_ERR_CUDA;
  }

  const int tiles = ((M + kBM * cg - 1) / (kBM * cg)) * ((N + bn - 1) / bn);
  const int workers = sms / cg;
  const int grid = (tiles < workers ? tiles : workers) * cg;

#define MB_DISPATCH_EPI(BN_, CG_)                                                                  \
  switch (epi) {                                                                                   \
    case MB_EPI_BIAS: return launch_gemm<BN_, MB_EPI_BIAS, CG_>(ta, tb, to, tr, p, grid, stream);          \
    case MB_EPI_GELU: return launch_gemm<BN_, MB_EPI_GELU, CG_>(ta, tb, to, tr, p, grid, stream);          \
    case MB_EPI_RESIDUAL: return launch_gemm<BN_, MB_EPI_RESIDUAL, CG_>(ta, tb, to, tr, p, grid, stream);  \
    default: break;                                                                                \
  }
  if (epi == MB_EPI_SWIGLU) {
    if (cg == 2) return launch_gemm<256, MB_EPI_SWIGLU, 2>(ta, tb, to, tr, p, grid, stream);
    return launch_gemm<256, MB_EPI_SWIGLU, 1>(ta, tb, to, tr, p, grid, stream);
  }
  if (cg == 2 && bn == 256) { MB_DISPATCH_EPI(256, 2) }
  else if (cg == 2) { MB_DISPATCH_EPI(128, 2) }
  else if (bn == 256) { MB_DISPATCH_EPI(256, 1) }
  else { MB_DISPATCH_EPI(128, 1) }
#undef MB_DISPATCH_EPI
  set_error("mb_gemm_bf16: unreachable dispatch");
  return MB_ERR_SHAPE;
}

// ------------------------------------------------------------------------------------------------------------
// Grouped (per-expert) GEMM of the routed MoE experts, prefill regime (modeling_bailing_moe.py:609-639 replaces the
// Python loop over experts + one cuBLAS call per expert and projection):
//   swiglu = 1:  out[r, i] = bf16( bf16(silu(bf16(A[r] . Wg[e(r)][i]))) * bf16(A[r] . Wu[e(r)][i]) ),  W = [E][2I][K]
//   swiglu = 0:  out[r, n] = bf16( A[r] . W[e(r)][n] ),                                               W = [E][N][K]
// Rows are expert-sorted with every expert segment padded to 128 rows (mb_moe_plan), e(r) = tile_expert[r / 128].
// Same tcgen05 / TMA kernel as the dense GEMM (128 x 256 tile per CTA, persistent over the device-side tile count).
// ------------------------------------------------------------------------------------------------------------
extern "C" int mb_moe_grouped_gemm(const void* A, const void* W, void* out, const int32_t* tile_expert,
                                   const int32_t* num_m_tiles, int max_rows, int N, int K, int E, int swiglu,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_grouped_gemm: no sm_100 device");
  MB_CHECK_ARG(max_rows >= 0 && max_rows % 128 == 0 && N >= 8 && K >= 8 && E >= 1, MB_ERR_SHAPE,
               "mb_moe_grouped_gemm: bad shape rows=%d N=%d K=%d E=%d (rows must be a multiple of 128)", max_rows, N, K, E);
  if (max_rows == 0) return MB_OK;
  const int n_out = swiglu ? N / 2 : N;
  MB_CHECK_ARG(K % 8 == 0 && n_out % 8 == 0 && (!swiglu || N % 16 == 0), MB_ERR_ALIGN,
               "mb_moe_grouped_gemm: K and the output width must be multiples of 8 (K=%d N=%d)", K, N);
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0 && tile_expert != nullptr && num_m_tiles != nullptr,
               MB_ERR_ALIGN, "mb_moe_grouped_gemm: A/W/out must be 16-byte aligned, plan pointers non-null");
  constexpr int BN = 256;
  CUtensorMap ta, tb, to;
  if (!make_tmap_2d_bf16(&ta, A, K, max_rows, K, kBK, kBM)) return MB_ERR_CUDA;
  if (!make_tmap_2d_bf16(&tb, W, K, static_cast<uint64_t>(E) * N, K, kBK, swiglu ? BN / 2 : BN)) return MB_ERR_CUDA;
  if (!make_tmap_2d_bf16(&to, out, n_out, max_rows, n_out, 64, 32)) return MB_ERR_CUDA;
  GemmParams p = {};
  p.M = max_rows;
  // swiglu: one "N tile" = 128 gate + 128 up rows -> ceil(I / 128) tiles; the kernel derives that from p.N / BN
  p.N = swiglu ? 2 * (((n_out + 127) / 128) * 128) : N;
  p.K = K;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = n_out;
  p.tma_epi = 1;
  p.ln_inv_dim = 1.f;
  p.ln_slots_out = (n_out + 63) / 64;
  p.grp_tile_expert = tile_expert;
  p.grp_num_m_tiles = num_m_tiles;
  p.grp_split = swiglu ? n_out : 0;
  p.grp_w_rows = N;
  p.grp_n_out = n_out;
  const int sms = mb::num_sms();
  const long tiles = static_cast<long>(max_rows / 128) * ((p.N + BN - 1) / BN);
  const int grid = static_cast<int>(tiles < sms ? tiles : sms);
  if (swiglu) return launch_gemm<256, MB_EPI_SWIGLU, 1>(ta, tb, to, ta, p, grid, stream);
  return launch_gemm<256, MB_EPI_BIAS, 1>(ta, tb, to, ta, p, grid, stream);
}

// ------------------------------------------------------------------------------------------------------------
// SwiGLU weight pre-pack
// ------------------------------------------------------------------------------------------------------------
namespace mb {
__global__ void pack_swiglu_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int H,
                                        int Hp, int K) {
  // one block per destination row
  const int drow = blockIdx.x;
  const int blk = drow / 256;
  const int within = drow % 256;
  const int is_up = within >= 128;
  const int h = blk * 128 + (within & 127);
  __nv_bfloat16* d = dst + static_cast<int64_t>(drow) * K;
  if (h < H) {
    const __nv_bfloat16* s = src + static_cast<int64_t>(is_up ? H + h : h) * K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) d[k] = s[k];
  } else {
    for (int k = threadIdx.x; k < K; k += blockDim.x) d[k] = __float2bfloat16_rn(0.f);
  }
}
}  // namespace mb

extern "C" int mb_pack_swiglu_rows(const void* src, void* dst, int H, int Hp, int K, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(H >= 1 && Hp >= H && Hp % 128 == 0 && K >= 1, MB_ERR_SHAPE,
               "mb_pack_swiglu_rows: need Hp %% 128 == 0 and Hp >= H (H=%d Hp=%d K=%d)", H, Hp, K);
  pack_swiglu_rows_kernel<<<2 * Hp, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(src),
                                                      static_cast<__nv_bfloat16*>(dst), H, Hp, K);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}
