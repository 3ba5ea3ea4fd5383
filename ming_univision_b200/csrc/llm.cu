// Bailing-MoE AR-step operators (mingunivision/modeling_bailing_moe.py): RMSNorm, NeoX RoPE fused with the KV-cache
// append, GQA decode attention with a per-row key mask, softmax-top-k router, and the routed-expert FFN for small
// token counts (pairs grouped by expert so every expert's weights are streamed once per chunk of <= 8 tokens).
// All of these are HBM / latency bound at the CFG-row batch sizes of the path; they use 16-byte coalesced loads,
// warp-shuffle reductions and read the current cache length from DEVICE memory so a whole AR step can be replayed as
// one CUDA graph.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint4 ldg_stream16(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float dot8f(const uint4& w, const uint4& a) {
  const float2 w0 = unpack_bf16x2(w.x), w1 = unpack_bf16x2(w.y), w2 = unpack_bf16x2(w.z), w3 = unpack_bf16x2(w.w);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  float s = w0.x * a0.x;
  s = fmaf(w0.y, a0.y, s); s = fmaf(w1.x, a1.x, s); s = fmaf(w1.y, a1.y, s);
  s = fmaf(w2.x, a2.x, s); s = fmaf(w2.y, a2.y, s); s = fmaf(w3.x, a3.x, s); s = fmaf(w3.y, a3.y, s);
  return s;
}

// ------------------------------------------------------------------------------------------------------------
// RMSNorm: y = bf16(w * (x * rsqrt(mean(x^2) + eps)))   one warp per row
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int dim, float eps) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + static_cast<int64_t>(row) * ldx;
  float sq = 0.f;
  for (int c = lane * 8; c < dim; c += 256) {
    const uint4 q = *reinterpret_cast<const uint4*>(xr + c);
    const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), cc = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
    sq += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
  }
  const float r = rsqrtf(warp_sum_f(sq) / dim + eps);
  __nv_bfloat16* yr = y + static_cast<int64_t>(row) * ldy;
  for (int c = lane * 8; c < dim; c += 256) {
    const uint4 q = *reinterpret_cast<const uint4*>(xr + c);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(w + c));
    const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), cc = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
    const float2 ga = unpack_bf16x2(g.x), gb = unpack_bf16x2(g.y), gc = unpack_bf16x2(g.z), gd = unpack_bf16x2(g.w);
    uint4 o;
    o.x = pack_bf16x2(ga.x * (a.x * r), ga.y * (a.y * r));
    o.y = pack_bf16x2(gb.x * (b.x * r), gb.y * (b.y * r));
    o.z = pack_bf16x2(gc.x * (cc.x * r), gc.y * (cc.y * r));
    o.w = pack_bf16x2(gd.x * (d.x * r), gd.y * (d.y * r));
    *reinterpret_cast<uint4*>(yr + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// RoPE (rotate-half, 1-D positions, bf16 cos/sin tables as the reference's `.to(x.dtype)`) + KV-cache append.
// qkv rows: [H q heads | Hkv k heads | Hkv v heads] x hd.  One CTA per token row, one warp per head.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rope_kv_append_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ position_ids,
                      __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ kcache,
                      __nv_bfloat16* __restrict__ vcache, int B, int S, int H, int Hkv, int hd, int Tmax,
                      const int32_t* __restrict__ t_dev, int t_host, float theta) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  const int row = blockIdx.x;  // b * S + s
  const int b = row / S, s = row % S;
  const int slot = (t_dev ? *t_dev : 0) + t_host + s;
  const int pos = position_ids[row];
  const int nheads = H + 2 * Hkv;
  const int half = hd / 2;
  const __nv_bfloat16* src_row = qkv + static_cast<int64_t>(row) * nheads * hd;
  for (int head = threadIdx.x >> 5; head < nheads; head += blockDim.x >> 5) {
    const __nv_bfloat16* src = src_row + head * hd;
    __nv_bfloat16* dst;
    bool rot = true;
    if (head < H) {
      dst = q_out + (static_cast<int64_t>(row) * H + head) * hd;
    } else if (head < H + Hkv) {
      dst = kcache + ((static_cast<int64_t>(b) * Hkv + (head - H)) * Tmax + slot) * hd;
    } else {
      dst = vcache + ((static_cast<int64_t>(b) * Hkv + (head - H - Hkv)) * Tmax + slot) * hd;
      rot = false;
    }
    for (int i = threadIdx.x & 31; i < half; i += 32) {
      const float x1 = __bfloat162float(src[i]), x2 = __bfloat162float(src[i + half]);
      if (!rot) {
        dst[i] = src[i];
        dst[i + half] = src[i + half];
      } else {
        // inv_freq = theta^(-2i/hd); angle in fp32 exactly as torch.outer(t, inv_freq) (modeling_bailing_moe.py:219-224)
        const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * i) / static_cast<float>(hd));
        const float ang = static_cast<float>(pos) * inv_freq;
        const float c = bf16_round(cosf(ang)), sn = bf16_round(sinf(ang));
        // q*cos + rotate_half(q)*sin with bf16 rounding after every tensor op (:455-461)
        dst[i] = __float2bfloat16_rn(bf16_round(x1 * c) + bf16_round(-x2 * sn));
        dst[i + half] = __float2bfloat16_rn(bf16_round(x2 * c) + bf16_round(x1 * sn));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// GQA decode attention (q_len = 1), head_dim 128: one warp per (batch row, q head); lane owns 4 dims.
// Keys 0..T-1 of the static cache, skipping keys whose mask entry is 0 (the reference's unpad/varlen path).
// ------------------------------------------------------------------------------------------------------------
// Long contexts split the keys over gridDim.y CTAs per head (chunks of `chunk` keys, flash-decoding style): each CTA
// then writes its un-normalised partial (max, sum, output[128]) to `partial`, and attn_decode_merge_kernel combines them.
__global__ void __launch_bounds__(128)
attn_decode_gqa128_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kcache,
                          const __nv_bfloat16* __restrict__ vcache, const int32_t* __restrict__ key_mask,
                          int64_t mask_stride, __nv_bfloat16* __restrict__ out, int B, int H, int Hkv, int Tmax,
                          const int32_t* __restrict__ t_dev, int t_host, float scale, float* __restrict__ partial,
                          int chunk) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  // one CTA (4 warps) per (batch row, q head): the warps take interleaved groups of 4 keys, then merge their partial
  // (max, sum, output) triples through shared memory — 4x shorter dependent chain than one warp per head
  __shared__ float s_m[4], s_l[4], s_o[4][128];
  const int bh = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = bh / H, h = bh % H, hk = h / (H / Hkv);
  const int T_all = (t_dev ? *t_dev : 0) + t_host;
  const int j_begin = partial ? static_cast<int>(blockIdx.y) * chunk : 0;
  const int T = partial ? min(T_all, j_begin + chunk) : T_all;
  const uint2 qq = *reinterpret_cast<const uint2*>(q + (static_cast<int64_t>(b) * H + h) * 128 + lane * 4);
  const float2 q0 = unpack_bf16x2(qq.x), q1 = unpack_bf16x2(qq.y);
  const __nv_bfloat16* kc = kcache + (static_cast<int64_t>(b) * Hkv + hk) * Tmax * 128 + lane * 4;
  const __nv_bfloat16* vc = vcache + (static_cast<int64_t>(b) * Hkv + hk) * Tmax * 128 + lane * 4;
  const int32_t* mk = key_mask ? key_mask + static_cast<int64_t>(b) * mask_stride : nullptr;
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  for (int j0 = j_begin + warp * 4; j0 < T; j0 += 16) {
    uint2 kk[4], vv[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      ok[u] = j < T && (mk == nullptr || mk[j] != 0);
      if (ok[u]) {
        kk[u] = *reinterpret_cast<const uint2*>(kc + static_cast<int64_t>(j) * 128);
        vv[u] = *reinterpret_cast<const uint2*>(vc + static_cast<int64_t>(j) * 128);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;  // warp-uniform
      const float2 k0 = unpack_bf16x2(kk[u].x), k1 = unpack_bf16x2(kk[u].y);
      float s = q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y;
      s = warp_sum_f(s) * scale;
      const float m_new = fmaxf(m, s);
      const float corr = __expf(m - m_new), pj = __expf(s - m_new);
      const float2 v0 = unpack_bf16x2(vv[u].x), v1 = unpack_bf16x2(vv[u].y);
      l = l * corr + pj;
      o0 = o0 * corr + pj * v0.x; o1 = o1 * corr + pj * v0.y;
      o2 = o2 * corr + pj * v1.x; o3 = o3 * corr + pj * v1.y;
      m = m_new;
    }
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  s_o[warp][lane * 4 + 0] = o0; s_o[warp][lane * 4 + 1] = o1;
  s_o[warp][lane * 4 + 2] = o2; s_o[warp][lane * 4 + 3] = o3;
  __syncthreads();
  if (warp == 0) {
    const float mm = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float lt = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - mm);
      lt += s_l[w] * c;
      r0 += s_o[w][lane * 4 + 0] * c; r1 += s_o[w][lane * 4 + 1] * c;
      r2 += s_o[w][lane * 4 + 2] * c; r3 += s_o[w][lane * 4 + 3] * c;
    }
    if (partial != nullptr) {  // [bh][split][2 + 128] floats: max, sum, un-normalised output
      float* dst = partial + (static_cast<int64_t>(bh) * gridDim.y + blockIdx.y) * 130;
      if (lane == 0) { dst[0] = mm; dst[1] = lt; }
      *reinterpret_cast<float2*>(dst + 2 + lane * 4) = make_float2(r0, r1);
      *reinterpret_cast<float2*>(dst + 4 + lane * 4) = make_float2(r2, r3);
      return;
    }
    const float inv = lt > 0.f ? 1.f / lt : 0.f;
    uint2 r;
    r.x = pack_bf16x2(r0 * inv, r1 * inv);
    r.y = pack_bf16x2(r2 * inv, r3 * inv);
    *reinterpret_cast<uint2*>(out + (static_cast<int64_t>(b) * H + h) * 128 + lane * 4) = r;
  }
}

// one warp per (batch row, q head): combines the per-chunk partials of attn_decode_gqa128_kernel
__global__ void __launch_bounds__(128)
attn_decode_merge_kernel(const float* __restrict__ partial, __nv_bfloat16* __restrict__ out, int n_heads, int n_splits) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  const int bh = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (bh >= n_heads) return;
  const float* src = partial + static_cast<int64_t>(bh) * n_splits * 130;
  float mm = -INFINITY;
  for (int sp = 0; sp < n_splits; ++sp) mm = fmaxf(mm, src[sp * 130]);
  float lt = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
  for (int sp = 0; sp < n_splits; ++sp) {
    const float ms = src[sp * 130];
    if (ms == -INFINITY) continue;  // chunk past the end of the context, or fully masked
    const float c = __expf(ms - mm);
    lt += src[sp * 130 + 1] * c;
    const float2 a = *reinterpret_cast<const float2*>(src + sp * 130 + 2 + lane * 4);
    const float2 b2 = *reinterpret_cast<const float2*>(src + sp * 130 + 4 + lane * 4);
    r0 += a.x * c; r1 += a.y * c; r2 += b2.x * c; r3 += b2.y * c;
  }
  const float inv = lt > 0.f ? 1.f / lt : 0.f;
  uint2 r;
  r.x = pack_bf16x2(r0 * inv, r1 * inv);
  r.y = pack_bf16x2(r2 * inv, r3 * inv);
  *reinterpret_cast<uint2*>(out + static_cast<int64_t>(bh) * 128 + lane * 4) = r;
}

// two-stage greedy sampling for large vocabularies: per-chunk (max, first index), then the row winner
__global__ void __launch_bounds__(256)
argmax_partial_kernel(const float* __restrict__ x, float* __restrict__ pval, int32_t* __restrict__ pidx, int V,
                      int chunk) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int row = blockIdx.y, c0 = blockIdx.x * chunk, c1 = min(V, c0 + chunk);
  const float* xr = x + static_cast<int64_t>(row) * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = c0 + threadIdx.x; i < c1; i += 256) {
    const float v = xr[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    pval[row * gridDim.x + blockIdx.x] = best;
    pidx[row * gridDim.x + blockIdx.x] = bi;
  }
}
__global__ void argmax_final_kernel(const float* __restrict__ pval, const int32_t* __restrict__ pidx,
                                    int32_t* __restrict__ out, int n_chunks) {
  const int row = blockIdx.x, lane = threadIdx.x;  // one warp per row
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < n_chunks; c += 32) {
    const float v = pval[row * n_chunks + c];
    const int i = pidx[row * n_chunks + c];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) out[row] = bi;
}

// ------------------------------------------------------------------------------------------------------------
// Router: fp32 softmax over E bf16 logits, top-k (ties -> lowest index), optional renormalisation.   warp per token
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
router_topk_kernel(const __nv_bfloat16* __restrict__ logits, const __nv_bfloat16* __restrict__ logits_img,
                   const uint8_t* __restrict__ image_mask, int32_t* __restrict__ idx, float* __restrict__ wout, int T,
                   int E, int k, int renorm) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= T) return;
  const __nv_bfloat16* lg = (logits_img != nullptr && image_mask != nullptr && image_mask[t]) ? logits_img : logits;
  lg += static_cast<int64_t>(t) * E;
  // up to 8 experts per lane (E <= 256)
  float v[8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int e = lane + 32 * i;
    v[i] = e < E ? __bfloat162float(lg[e]) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max_f(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = (lane + 32 * i < E) ? expf(v[i] - mx) : -1.f;
    if (v[i] > 0.f) sum += v[i];
  }
  sum = warp_sum_f(sum);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = v[i] >= 0.f ? v[i] / sum : -1.f;
  float wsum = 0.f, picked_w = 0.f;
  int picked_e = 0;
  for (int j = 0; j < k; ++j) {
    float best = -1.f;
    int best_e = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = lane + 32 * i;
      if (v[i] > best || (v[i] == best && e < best_e)) { best = v[i]; best_e = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oe = __shfl_xor_sync(0xffffffffu, best_e, o);
      if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (lane + 32 * i == best_e) v[i] = -1.f;  // remove the winner
    wsum += best;
    if (lane == j) { picked_w = best; picked_e = best_e; }
  }
  if (lane < k) {
    idx[static_cast<int64_t>(t) * k + lane] = picked_e;
    wout[static_cast<int64_t>(t) * k + lane] = renorm ? picked_w / wsum : picked_w;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Counting sort of the T*k (token, slot) pairs by expert (single CTA; T*k is small on this path).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
moe_sort_kernel(const int32_t* __restrict__ idx, int32_t* __restrict__ expert_offsets,
                int32_t* __restrict__ sorted_pair, int npairs, int E, int e_begin) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  // E = number of LOCAL experts [e_begin, e_begin + E); pairs routed elsewhere (expert parallelism) are not listed
  extern __shared__ int32_t sm[];  // counts[E], cursor[E]
  int32_t* counts = sm;
  int32_t* cursor = sm + E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) counts[e] = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
    const int e = idx[p] - e_begin;
    if (e >= 0 && e < E) atomicAdd(&counts[e], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int e = 0; e < E; ++e) {
      expert_offsets[e] = acc;
      cursor[e] = acc;
      acc += counts[e];
    }
    expert_offsets[E] = acc;
  }
  __syncthreads();
  // stable order inside an expert is not required by the math (each pair is independent), but keep it deterministic:
  // thread 0..E-1 each fills its own expert's segment in increasing pair order
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    int c = cursor[e];
    if (counts[e] == 0) continue;
    for (int p = 0; p < npairs; ++p)
      if (idx[p] - e_begin == e) sorted_pair[c++] = p;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Routed experts, small-token regime.  grid = (X row-tile blocks, E experts).  For its expert a CTA walks the expert's
// pairs in chunks of <= 8 rows (= the 8 columns of mma.m16n8k16): stages the rows in shared memory, then every warp
// streams 16-row weight tiles straight from global memory into tensor-core A fragments (same permuted-K scheme as
// gemv.cu: lane (g, t) loads 16 bytes of rows g and g + 8 at K offset 32 kg + 8 t; the activation fragment uses the
// same K permutation), so an expert's weights are read once per chunk at ~4 instructions per KB.
//   PHASE 0: hid[p, i]  = bf16(bf16(silu(bf16(x[tok(p)] . Wg[e][i]))) * bf16(x[tok(p)] . Wu[e][i]))     (gate/up + SwiGLU)
//   PHASE 1: out[pair(p), n] = bf16(hid[p] . Wd[e][n])                                                   (down)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816_f(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                           uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NT = 8-row column groups per pass: an expert's pairs are walked in chunks of 8 * NT rows, so its weights stream ONCE
// per 8 * NT pairs (NT = 1: the CFG rows of one request; NT = 4: rows gathered from several requests / ranks, or a short
// prefill — up to 32 pairs per expert and pass).
template <int PHASE, int NT>
__global__ void __launch_bounds__(256)
moe_expert_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W,
                  const int32_t* __restrict__ expert_offsets, const int32_t* __restrict__ sorted_pair,
                  __nv_bfloat16* __restrict__ dst, int topk, int K, int n_cols, int64_t expert_stride) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  constexpr int kTiles = (PHASE == 0) ? 2 : 1;
  constexpr int kUnroll = (PHASE == 0) ? 4 : 8;
  constexpr int kRows = 8 * NT;
  extern __shared__ __align__(16) uint8_t moe_smem[];
  const int e = blockIdx.y;
  const int beg = expert_offsets[e], end = expert_offsets[e + 1];
  if (beg == end) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int kchunks = K >> 3;
  const int kgroups = (kchunks + 3) >> 2;
  const int row_bytes = kgroups * 64 + 64;
  const __nv_bfloat16* We = W + static_cast<int64_t>(e) * expert_stride;
  const int n_items = (n_cols + 15) / 16;
  for (int p0 = beg; p0 < end; p0 += kRows) {
    const int cnt = min(kRows, end - p0);
    __syncthreads();  // previous chunk fully consumed
    for (int i = tid; i < (cnt + 1) * (row_bytes / 16); i += blockDim.x) {
      const int m = i / (row_bytes / 16), c = i % (row_bytes / 16);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (m < cnt && c < kchunks) {
        // PHASE 0 gathers token rows of x; PHASE 1 reads the (already expert-sorted) hidden rows
        const int64_t src_row = (PHASE == 0) ? (sorted_pair[p0 + m] / topk) : (p0 + m);
        v = *reinterpret_cast<const uint4*>(A + src_row * K + c * 8);
      }
      *reinterpret_cast<uint4*>(moe_smem + m * row_bytes + c * 16) = v;
    }
    __syncthreads();
    const uint8_t* xrow[NT];  // this lane's activation row in every column group (rows >= cnt: the shared zero row)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) xrow[nt] = moe_smem + min(g + 8 * nt, cnt) * row_bytes + t * 16;
    const int nwarps = blockDim.x >> 5;
    for (int item = blockIdx.x * nwarps + warp; item < n_items; item += gridDim.x * nwarps) {
      const int n0 = item * 16;
      const __nv_bfloat16* wr[kTiles][2];
#pragma unroll
      for (int tl = 0; tl < kTiles; ++tl) {
        const int base = tl * n_cols;  // PHASE 0: up rows follow the gate rows inside the expert slab
        wr[tl][0] = We + static_cast<int64_t>(base + min(n0 + g, n_cols - 1)) * K + t * 8;
        wr[tl][1] = We + static_cast<int64_t>(base + min(n0 + g + 8, n_cols - 1)) * K + t * 8;
      }
      float acc[kTiles][NT][4];
#pragma unroll
      for (int tl = 0; tl < kTiles; ++tl)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[tl][nt][0] = acc[tl][nt][1] = acc[tl][nt][2] = acc[tl][nt][3] = 0.f;
      for (int kg0 = 0; kg0 < kgroups; kg0 += kUnroll) {
        uint4 w[kTiles][2][kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int kg = kg0 + u;
          const bool ok = kg < kgroups && (kg * 4 + t) < kchunks;
#pragma unroll
          for (int tl = 0; tl < kTiles; ++tl) {
            w[tl][0][u] = ok ? ldg_stream16(reinterpret_cast<const uint4*>(wr[tl][0] + kg * 32)) : make_uint4(0, 0, 0, 0);
            w[tl][1][u] = ok ? ldg_stream16(reinterpret_cast<const uint4*>(wr[tl][1] + kg * 32)) : make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int kg = kg0 + u;
          if (kg < kgroups) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              if (nt * 8 >= cnt) continue;  // (uniform per chunk) column groups without rows
              const uint4 xb = *reinterpret_cast<const uint4*>(xrow[nt] + kg * 64);
#pragma unroll
              for (int tl = 0; tl < kTiles; ++tl) {
                mma16816_f(acc[tl][nt], w[tl][0][u].x, w[tl][1][u].x, w[tl][0][u].y, w[tl][1][u].y, xb.x, xb.y);
                mma16816_f(acc[tl][nt], w[tl][0][u].z, w[tl][1][u].z, w[tl][0][u].w, w[tl][1][u].w, xb.z, xb.w);
              }
            }
          }
        }
      }
      // accumulator element ei: weight row n0 + g + 8 (ei >> 1), chunk row (token) 8 nt + 2 t + (ei & 1)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int ei = 0; ei < 4; ++ei) {
          const int n = n0 + g + 8 * (ei >> 1);
          const int m = 8 * nt + 2 * t + (ei & 1);
          if (n < n_cols && m < cnt) {
            if (PHASE == 0) {
              const float gg = bf16_round(acc[0][nt][ei]), uu = bf16_round(acc[kTiles - 1][nt][ei]);
              dst[static_cast<int64_t>(p0 + m) * n_cols + n] = __float2bfloat16_rn(bf16_round(silu(gg)) * uu);
            } else {
              dst[static_cast<int64_t>(sorted_pair[p0 + m]) * n_cols + n] = __float2bfloat16_rn(acc[0][nt][ei]);
            }
          }
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// Routing plan: bucket-sorted row layout of the T*k (token, slot) pairs.  Bucket of pair p = (idx[p] - e_begin) / div
// (pairs outside [0, E) buckets are "not local"), every bucket's segment padded to `granule` rows.  Two uses:
//   * grouped (prefill) expert GEMMs: div = 1, granule = 128 -> each 128-row GEMM tile belongs to ONE local expert
//   * expert-parallel dispatch:       div = experts per rank, granule = 1 -> send buffer ordered by destination rank
//   pair_row[p]     row of pair p = t*k + j, or -1 if not local
//   row_token[r]    source token (p / k) of row r, or -1 for padding rows
//   tile_expert[m]  bucket of 128-row tile m (granule = 128 only; else NULL)
//   meta[0] = number of 128-row tiles, meta[1] = number of rows incl. padding;  counts[b] = pairs in bucket b (or NULL)
// Single CTA (T*k <= a few 10^4 pairs).  Row order inside a bucket is the arrival order of a shared-memory atomic: it
// may differ from run to run, but every output row of the grouped GEMMs depends on its own input row only and the
// combine addresses rows through pair_row, so the RESULTS are bitwise reproducible.
// ------------------------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;
__global__ void __launch_bounds__(kPlanThreads)
moe_plan_kernel(const int32_t* __restrict__ idx, int32_t* __restrict__ pair_row, int32_t* __restrict__ row_token,
                int32_t* __restrict__ tile_expert, int32_t* __restrict__ meta, int32_t* __restrict__ counts_out,
                int npairs, int topk, int E, int e_begin, int div, int granule, int max_rows) {
  extern __shared__ int32_t psm[];  // counts[E], pad_off[E + 1]
  int32_t* counts = psm;
  int32_t* pad_off = psm + E;
  const int tid = threadIdx.x;
  const int last = e_begin + E * div;  // first global expert id past the local buckets
  for (int i = tid; i < E; i += kPlanThreads) counts[i] = 0;
  for (int r = tid; r < max_rows; r += kPlanThreads) row_token[r] = -1;
  __syncthreads();
  // pass 1: rank of every local pair inside its bucket (smem atomic), parked in pair_row until the offsets exist
  for (int p = tid; p < npairs; p += kPlanThreads) {
    const int g = idx[p];
    if (g >= e_begin && g < last) pair_row[p] = atomicAdd(&counts[(g - e_begin) / div], 1);
  }
  __syncthreads();
  if (tid < 32) {  // one warp: padded offsets by a shuffle scan over the buckets, 32 at a time
    int rows_carry = 0;
    for (int e0 = 0; e0 < E; e0 += 32) {
      const int e = e0 + tid;
      const int c = (e < E) ? counts[e] : 0;
      const int padded = ((c + granule - 1) / granule) * granule;
      int incl = padded;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      const int first = rows_carry + incl - padded;
      if (e < E) {
        pad_off[e] = first;
        if (counts_out != nullptr) counts_out[e] = c;
        if (tile_expert != nullptr)
          for (int t = 0; t < (padded >> 7); ++t) tile_expert[(first >> 7) + t] = e;
      }
      rows_carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) {
      pad_off[E] = rows_carry;
      meta[0] = (rows_carry + 127) >> 7;
      meta[1] = rows_carry;
    }
  }
  __syncthreads();
  for (int p = tid; p < npairs; p += kPlanThreads) {  // same thread wrote pair_row[p] in pass 1
    const int g = idx[p];
    int r = -1;
    if (g >= e_begin && g < last) {
      r = pad_off[(g - e_begin) / div] + pair_row[p];
      row_token[r] = p / topk;
    }
    pair_row[p] = r;
  }
}

// Ag[r] = x[row_token[r]] (zero rows for padding) for r < meta[1]; 16-byte vectors, one warp-row per loop step
__global__ void __launch_bounds__(256)
moe_gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ row_token,
                       const int32_t* __restrict__ meta, __nv_bfloat16* __restrict__ out, int n_rows_host, int D) {
  const int rows = (meta != nullptr) ? meta[1] : n_rows_host;
  const int vec_per_row = D >> 3;
  const int64_t total = static_cast<int64_t>(rows) * vec_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / vec_per_row);
    const int c = static_cast<int>(i % vec_per_row);
    const int t = __ldg(row_token + r);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (t >= 0) v = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(t) * D) + c);
    reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * D)[c] = v;
  }
}

// y[t] = bf16( bf16( bf16(sum_j w[t,j] * out_pairs[t*k+j]) + shared[t] ) + residual[t] )
// (moe_infer's fp32 weighted sum :632-638, `y + shared_experts(identity)` :604-605, layer residual :1226)
__global__ void moe_combine_kernel(const __nv_bfloat16* __restrict__ out_pairs, const float* __restrict__ w,
                                   const __nv_bfloat16* __restrict__ shared, const __nv_bfloat16* __restrict__ residual,
                                   __nv_bfloat16* __restrict__ y, float* __restrict__ y_partial,
                                   const int32_t* __restrict__ pair_row, int T, int k, int D) {
  pdl_launch_dependents();  // (PDL: the next kernel of the stream may start launching; it waits below us)
  pdl_wait();
  const int64_t total = static_cast<int64_t>(T) * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / D;
    const int d = static_cast<int>(i % D);
    float acc = 0.f;
    if (pair_row != nullptr) {  // grouped layout: pair p lives in row pair_row[p] (< 0: expert owned by another rank)
      for (int j = 0; j < k; ++j) {
        const int r = pair_row[t * k + j];
        if (r >= 0) acc += w[t * k + j] * __bfloat162float(out_pairs[static_cast<int64_t>(r) * D + d]);
      }
    } else {
      for (int j = 0; j < k; ++j) acc += w[t * k + j] * __bfloat162float(out_pairs[(t * k + j) * D + d]);
    }
    if (y_partial != nullptr) {  // expert-parallel: this rank's share of the fp32 sum; finalised after the all-reduce
      y_partial[i] = acc;
      continue;
    }
    float v = bf16_round(acc);
    if (shared) v = bf16_round(v + __bfloat162float(shared[i]));
    if (residual) v = v + __bfloat162float(residual[i]);
    y[i] = __float2bfloat16_rn(v);
  }
}

__global__ void moe_finalize_kernel(const float* __restrict__ y_sum, const __nv_bfloat16* __restrict__ shared,
                                    const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ y,
                                    int64_t total) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = bf16_round(y_sum[i]);
    if (shared) v = bf16_round(v + __bfloat162float(shared[i]));
    if (residual) v = v + __bfloat162float(residual[i]);
    y[i] = __float2bfloat16_rn(v);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Expert-parallel combine FUSED with its exchange over NVLink peer memory (decode-sized inputs, tokens replicated on
// every rank): no NCCL call, no host involvement.  Every rank owns an exchange area in symmetric (peer-mapped) memory:
//     slots [2 parities][G source ranks][Tmax * D] fp32,   flags [G] u32,   then local-only: epoch u32, done u32
// and knows the base pointer of every peer's area (`peers`, a device array of G pointers).
//   mb_moe_combine_push   : this rank's fp32 partial sums  sum_j w[t,j] * out_pairs[row(t,j)]  are STORED DIRECTLY INTO
//                           EVERY PEER'S slot [parity][my_rank] (plain st.global on the peer pointers: the stores
//                           travel over NVLink while the kernel is still computing); the last CTA to finish issues a
//                           system-scope fence and raises flag[my_rank] = epoch in every peer's area.
//   mb_moe_reduce_finalize: waits (bounded spin, ld.acquire.sys) until all G flags of the LOCAL area show the epoch,
//                           adds the G slots in rank order (so every rank computes bit-identical sums), applies the
//                           reference's rounding chain  bf16(bf16(bf16(sum) + shared) + residual)  and bumps the epoch.
// Two parities suffice: a rank can be at most one call ahead of its slowest peer (it cannot pass its own reduce of call
// n + 1 before every peer has pushed call n + 1, i.e. finished its reduce of call n).
// ------------------------------------------------------------------------------------------------------------
struct PeerArea {
  static __host__ __device__ size_t slot_floats(int G, int Tmax, int D) { return static_cast<size_t>(2) * G * Tmax * D; }
};
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
moe_combine_push_kernel(const __nv_bfloat16* __restrict__ out_pairs, const float* __restrict__ w,
                        const int32_t* __restrict__ pair_row, float* const* __restrict__ peers, int my_rank, int G,
                        int T, int Tmax, int k, int D) {
  float* local = peers[my_rank];
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(local + PeerArea::slot_floats(G, Tmax, D));  // flags[G], epoch, done
  const uint32_t epoch = ctrl[G] + 1;
  const size_t slot_off = (static_cast<size_t>(epoch & 1) * G + my_rank) * Tmax * D;
  const int64_t total = static_cast<int64_t>(T) * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / D;
    const int d = static_cast<int>(i % D);
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      const int r = pair_row ? pair_row[t * k + j] : static_cast<int>(t * k + j);
      if (r >= 0) acc += w[t * k + j] * __bfloat162float(out_pairs[static_cast<int64_t>(r) * D + d]);
    }
    for (int pr = 0; pr < G; ++pr) peers[pr][slot_off + i] = acc;  // NVLink peer stores (local store for pr == my_rank)
  }
  // last CTA: everything this rank pushed is ordered before the flags it raises
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = atomicAdd(&ctrl[G + 1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    for (int pr = threadIdx.x; pr < G; pr += blockDim.x) {
      uint32_t* pctrl = reinterpret_cast<uint32_t*>(peers[pr] + PeerArea::slot_floats(G, Tmax, D));
      st_release_sys_u32(pctrl + my_rank, epoch);
    }
    if (threadIdx.x == 0) ctrl[G + 1] = 0;  // done counter ready for the next call
  }
}

__global__ void __launch_bounds__(256)
moe_reduce_finalize_kernel(float* const* __restrict__ peers, int my_rank, int G, int T, int Tmax, int D,
                           const __nv_bfloat16* __restrict__ shared, const __nv_bfloat16* __restrict__ residual,
                           __nv_bfloat16* __restrict__ y, uint32_t* __restrict__ fin_done) {
  const float* local = peers[my_rank];
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(peers[my_rank] + PeerArea::slot_floats(G, Tmax, D));
  const uint32_t epoch = ctrl[G] + 1;
  if (threadIdx.x < G) {  // bounded wait for every source rank's flag (a diverged peer must trap, not hang the GPU)
    uint32_t spins = 0;
    // flags only grow; a fast peer may already have raised the NEXT epoch (it writes the other parity)
    while (static_cast<int32_t>(ld_acquire_sys_u32(ctrl + threadIdx.x) - epoch) < 0) {
      __nanosleep(64);
      if (++spins > (1u << 24)) {
        printf("moe_reduce_finalize: rank %d never saw epoch %u from rank %d\n", my_rank, epoch, (int)threadIdx.x);
        __trap();
      }
    }
  }
  __syncthreads();
  const float* slots = local + static_cast<size_t>(epoch & 1) * G * Tmax * D;
  const int64_t total = static_cast<int64_t>(T) * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < G; ++r) sum += __ldcg(slots + static_cast<size_t>(r) * Tmax * D + i);  // rank order: deterministic
    float v = bf16_round(sum);
    if (shared) v = bf16_round(v + __bfloat162float(shared[i]));
    if (residual) v = v + __bfloat162float(residual[i]);
    y[i] = __float2bfloat16_rn(v);
  }
  // the last CTA to finish advances the epoch (all CTAs have read `epoch` and the slots by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(fin_done, 1u) == gridDim.x - 1) {
      *fin_done = 0;
      ctrl[G] = epoch;
    }
  }
}

// greedy sampling: first index of the row maximum (torch.argmax tie rule)   one CTA per row
__global__ void __launch_bounds__(256)
argmax_f32_kernel(const float* __restrict__ x, int32_t* __restrict__ out, int V) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const float* xr = x + static_cast<int64_t>(blockIdx.x) * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 256) {
    const float v = xr[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    out[blockIdx.x] = bi;
  }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_argmax_f32(const float* x, int32_t* out, int rows, int V, float* ws_val, int32_t* ws_idx, int n_chunks,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_argmax_f32: no sm_100 device");
  MB_CHECK_ARG(rows >= 0 && V >= 1 && n_chunks >= 1, MB_ERR_SHAPE, "mb_argmax_f32: bad shape");
  if (rows == 0) return MB_OK;
  if (n_chunks == 1 || ws_val == nullptr || ws_idx == nullptr) {
    argmax_f32_kernel<<<rows, 256, 0, stream>>>(x, out, V);
    MB_CHECK_CUDA(cudaGetLastError());
    return MB_OK;
  }
  const int chunk = (V + n_chunks - 1) / n_chunks;
  argmax_partial_kernel<<<dim3(n_chunks, rows), 256, 0, stream>>>(x, ws_val, ws_idx, V, chunk);
  MB_CHECK_CUDA(cudaGetLastError());
  argmax_final_kernel<<<rows, 32, 0, stream>>>(ws_val, ws_idx, out, n_chunks);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int dim, float eps,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_rmsnorm: no sm_100 device");
  MB_CHECK_ARG(rows >= 0 && dim >= 8 && dim % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, MB_ERR_SHAPE,
               "mb_rmsnorm: dim, ldx, ldy must be multiples of 8");
  if (rows == 0) return MB_OK;
  MB_CHECK_CUDA(launch_pdl(rmsnorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream,
                           static_cast<const __nv_bfloat16*>(x), ldx, static_cast<const __nv_bfloat16*>(w),
                           static_cast<__nv_bfloat16*>(y), ldy, rows, dim, eps));
  return MB_OK;
}

extern "C" int mb_rope_kv_append(const void* qkv, const int32_t* position_ids, void* q_out, void* kcache, void* vcache,
                                 int B, int S, int H, int Hkv, int hd, int Tmax, const int32_t* t_dev, int t_host,
                                 float rope_theta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_rope_kv_append: no sm_100 device");
  MB_CHECK_ARG(B >= 0 && S >= 0 && H >= 1 && Hkv >= 1 && hd % 2 == 0, MB_ERR_SHAPE, "mb_rope_kv_append: bad shape");
  MB_CHECK_ARG(t_dev != nullptr || (t_host >= 0 && t_host + S <= Tmax), MB_ERR_SHAPE,
               "mb_rope_kv_append: cache overflow (t=%d S=%d Tmax=%d)", t_host, S, Tmax);
  if (B * S == 0) return MB_OK;
  MB_CHECK_CUDA(launch_pdl(rope_kv_append_kernel, dim3(B * S), dim3(256), 0, stream,
                           static_cast<const __nv_bfloat16*>(qkv), position_ids, static_cast<__nv_bfloat16*>(q_out),
                           static_cast<__nv_bfloat16*>(kcache), static_cast<__nv_bfloat16*>(vcache), B, S, H, Hkv, hd, Tmax,
                           t_dev, t_host, rope_theta));
  return MB_OK;
}

extern "C" int mb_attn_decode_gqa(const void* q, const void* kcache, const void* vcache, const int32_t* key_mask,
                                  int64_t mask_stride, void* out, int B, int H, int Hkv, int hd, int Tmax,
                                  const int32_t* t_dev, int t_host, float scale, float* workspace, int n_splits,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_attn_decode_gqa: no sm_100 device");
  MB_CHECK_ARG(hd == 128 && H % Hkv == 0, MB_ERR_SHAPE, "mb_attn_decode_gqa: head_dim must be 128 and H %% Hkv == 0");
  MB_CHECK_ARG(t_dev != nullptr || (t_host >= 0 && t_host <= Tmax), MB_ERR_SHAPE, "mb_attn_decode_gqa: bad length");
  MB_CHECK_ARG(n_splits >= 1 && n_splits <= 65535 && (n_splits == 1 || workspace != nullptr), MB_ERR_SHAPE,
               "mb_attn_decode_gqa: n_splits > 1 needs a workspace of B * H * n_splits * 130 floats");
  if (B == 0) return MB_OK;
  if (n_splits == 1) {
    MB_CHECK_CUDA(launch_pdl(attn_decode_gqa128_kernel, dim3(B * H), dim3(128), 0, stream,
                             static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kcache),
                             static_cast<const __nv_bfloat16*>(vcache), key_mask, mask_stride,
                             static_cast<__nv_bfloat16*>(out), B, H, Hkv, Tmax, t_dev, t_host, scale, nullptr, 0));
    return MB_OK;
  }
  // the longest context this call can see: the host length, or the whole cache when the length lives on the device
  const int t_bound = (t_dev != nullptr) ? Tmax : t_host;
  const int chunk = ((t_bound + n_splits - 1) / n_splits + 15) / 16 * 16;
  MB_CHECK_CUDA(launch_pdl(attn_decode_gqa128_kernel, dim3(B * H, n_splits), dim3(128), 0, stream,
                           static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(kcache),
                           static_cast<const __nv_bfloat16*>(vcache), key_mask, mask_stride,
                           static_cast<__nv_bfloat16*>(out), B, H, Hkv, Tmax, t_dev, t_host, scale, workspace,
                           chunk < 16 ? 16 : chunk));
  MB_CHECK_CUDA(launch_pdl(attn_decode_merge_kernel, dim3((B * H + 3) / 4), dim3(128), 0, stream, workspace,
                           static_cast<__nv_bfloat16*>(out), B * H, n_splits));
  return MB_OK;
}

extern "C" int mb_router_topk(const void* logits, const void* logits_img, const uint8_t* image_mask, int32_t* idx,
                              float* weights, int T, int E, int k, int renorm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_router_topk: no sm_100 device");
  MB_CHECK_ARG(E >= 1 && E <= 256 && k >= 1 && k <= 32 && k <= E, MB_ERR_SHAPE, "mb_router_topk: E <= 256, k <= 32");
  if (T == 0) return MB_OK;
  MB_CHECK_CUDA(launch_pdl(router_topk_kernel, dim3((T + 3) / 4), dim3(128), 0, stream,
                           static_cast<const __nv_bfloat16*>(logits), static_cast<const __nv_bfloat16*>(logits_img),
                           image_mask, idx, weights, T, E, k, renorm));
  return MB_OK;
}

extern "C" int mb_moe_sort(const int32_t* idx, int32_t* expert_offsets, int32_t* sorted_pair, int T, int k, int E,
                           int e_begin, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_sort: no sm_100 device");
  MB_CHECK_ARG(E >= 1 && E <= 4096 && e_begin >= 0, MB_ERR_SHAPE, "mb_moe_sort: bad expert range");
  MB_CHECK_CUDA(launch_pdl(moe_sort_kernel, dim3(1), dim3(256), 2 * E * sizeof(int32_t), stream, idx, expert_offsets,
                           sorted_pair, T * k, E, e_begin));
  return MB_OK;
}

template <int PHASE, int NT>
static int launch_moe_phase_nt(const void* A, const void* W, const int32_t* offs, const int32_t* sorted, void* dst,
                               int topk, int E, int K, int n_cols, int64_t expert_stride, int active_max,
                               cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(8 * NT + 1) * (((K / 8 + 3) / 4) * 64 + 64);
  MB_CHECK_ARG(K % 8 == 0 && smem <= 160 * 1024, MB_ERR_SHAPE, "moe: K must be a multiple of 8 and <= %d",
               NT == 1 ? 8192 : 2048);
  static bool attr_set = false;
  if (!attr_set) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(moe_expert_kernel<PHASE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  // One 16-row weight tile per warp.  Only the experts that received pairs do any work (at most `active_max` = the
  // call's pair count of the E CTA columns): with a single decode row (<= 8 pairs) eight-warp CTAs would leave more than
  // half of the SMs idle, so that case runs four-warp CTAs (measured: 3.47 -> 3.33 ms per text token); two or more CFG
  // rows keep the eight-warp shape (smaller CTAs measured slower there: the activation rows are re-staged by every CTA).
  const int n_items = (n_cols + 15) / 16;
  const int warps = (active_max <= 8) ? 4 : 8;
  int xblocks = (n_items + warps - 1) / warps;
  if (xblocks < 1) xblocks = 1;
  dim3 grid(xblocks, E);
  MB_CHECK_CUDA(launch_pdl(moe_expert_kernel<PHASE, NT>, grid, dim3(warps * 32), smem, stream,
                           static_cast<const __nv_bfloat16*>(A), static_cast<const __nv_bfloat16*>(W), offs, sorted,
                           static_cast<__nv_bfloat16*>(dst), topk, K, n_cols, expert_stride));
  return MB_OK;
}

// `mean_pairs` = expected (token, slot) pairs per expert of the call (T k / all experts, rounded up).  Up to ~6 the 8-column
// kernel almost always finishes an expert in one pass (P[Poisson(6) > 8] = 15 %, and a second pass re-reads the expert from
// L2) and keeps 2+ CTAs per SM in flight; beyond that (a short prefill, many requests per step) the 32-row form streams
// every expert once per 32 pairs.
static int launch_moe_phase(int phase, const void* A, const void* W, const int32_t* offs, const int32_t* sorted,
                            void* dst, int topk, int E, int K, int n_cols, int64_t expert_stride, int active_max,
                            int mean_pairs, cudaStream_t stream) {
  const bool wide = mean_pairs > 6 && K <= 2048;
  if (phase == 0)
    return wide ? launch_moe_phase_nt<0, 4>(A, W, offs, sorted, dst, topk, E, K, n_cols, expert_stride, active_max, stream)
                : launch_moe_phase_nt<0, 1>(A, W, offs, sorted, dst, topk, E, K, n_cols, expert_stride, active_max, stream);
  return wide ? launch_moe_phase_nt<1, 4>(A, W, offs, sorted, dst, topk, E, K, n_cols, expert_stride, active_max, stream)
              : launch_moe_phase_nt<1, 1>(A, W, offs, sorted, dst, topk, E, K, n_cols, expert_stride, active_max, stream);
}

extern "C" int mb_moe_gate_up(const void* x, const void* Wgu, const int32_t* expert_offsets, const int32_t* sorted_pair,
                              void* hid, int T, int k, int E, int D, int I, int mean_pairs, void* stream_) {
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_gate_up: no sm_100 device");
  if (T == 0) return MB_OK;
  return launch_moe_phase(0, x, Wgu, expert_offsets, sorted_pair, hid, k, E, D, I, static_cast<int64_t>(2) * I * D,
                          T * k, mean_pairs, static_cast<cudaStream_t>(stream_));
}

extern "C" int mb_moe_down(const void* hid, const void* Wd, const int32_t* expert_offsets, const int32_t* sorted_pair,
                           void* out_pairs, int T, int k, int E, int D, int I, int mean_pairs, void* stream_) {
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_down: no sm_100 device");
  if (T == 0) return MB_OK;
  return launch_moe_phase(1, hid, Wd, expert_offsets, sorted_pair, out_pairs, k, E, I, D, static_cast<int64_t>(D) * I,
                          T * k, mean_pairs, static_cast<cudaStream_t>(stream_));
}



// ------------------------------------------------------------------------------------------------------------
// Whole-stage driver of the routed-expert FFN for decode-sized inputs (moe_infer :608-639 + shared-expert add :604-605 +
// layer residual :1226): sort -> gate/up + SwiGLU -> down -> weighted combine, chained on the caller's stream with a
// caller-provided workspace (no allocation, no host synchronisation) — what a non-Python host calls per MoE layer.
// ------------------------------------------------------------------------------------------------------------
static size_t moe_ffn_ws_layout(int T, int k, int E, int D, int I, size_t (&off)[4]) {
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  off[0] = 0;                                                        // expert_offsets [E + 1] i32
  off[1] = up(off[0] + (static_cast<size_t>(E) + 1) * 4);            // sorted_pair [T k] i32
  off[2] = up(off[1] + static_cast<size_t>(T) * k * 4);              // hid [T k, I] bf16
  off[3] = up(off[2] + static_cast<size_t>(T) * k * I * 2);          // out_pairs [T k, D] bf16
  return up(off[3] + static_cast<size_t>(T) * k * D * 2);
}

extern "C" int mb_moe_ffn_workspace_bytes(int T, int k, int E, int D, int I, int64_t* bytes) {
  MB_CHECK_ARG(T >= 0 && k >= 1 && E >= 1 && D >= 8 && I >= 8 && bytes != nullptr, MB_ERR_SHAPE,
               "mb_moe_ffn_workspace_bytes: bad shape");
  size_t off[4];
  *bytes = static_cast<int64_t>(moe_ffn_ws_layout(T, k, E, D, I, off));
  return MB_OK;
}

extern "C" int mb_moe_ffn(const void* x, const int32_t* idx, const float* weights, const void* Wgu, const void* Wd,
                          const void* shared, const void* residual, void* y, void* workspace, int64_t workspace_bytes,
                          int T, int k, int E, int e_begin, int n_experts_total, int D, int I, void* stream_) {
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_ffn: no sm_100 device");
  MB_CHECK_ARG(T >= 0 && k >= 1 && E >= 1 && n_experts_total >= E && D % 8 == 0 && I % 8 == 0, MB_ERR_SHAPE,
               "mb_moe_ffn: bad shape (T=%d k=%d E=%d D=%d I=%d)", T, k, E, D, I);
  if (T == 0) return MB_OK;
  size_t off[4];
  const size_t need = moe_ffn_ws_layout(T, k, E, D, I, off);
  MB_CHECK_ARG(workspace != nullptr && static_cast<size_t>(workspace_bytes) >= need &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               MB_ERR_SHAPE, "mb_moe_ffn: workspace of %lld bytes, %zu needed (256-byte aligned)",
               static_cast<long long>(workspace_bytes), need);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int32_t* offs = reinterpret_cast<int32_t*>(ws + off[0]);
  int32_t* sorted = reinterpret_cast<int32_t*>(ws + off[1]);
  void* hid = ws + off[2];
  void* out_pairs = ws + off[3];
  const int mean_pairs = (T * k + n_experts_total - 1) / n_experts_total;
  int rc = mb_moe_sort(idx, offs, sorted, T, k, E, e_begin, stream_);
  if (rc != MB_OK) return rc;
  rc = mb_moe_gate_up(x, Wgu, offs, sorted, hid, T, k, E, D, I, mean_pairs, stream_);
  if (rc != MB_OK) return rc;
  rc = mb_moe_down(hid, Wd, offs, sorted, out_pairs, T, k, E, D, I, mean_pairs, stream_);
  if (rc != MB_OK) return rc;
  // (pairs of experts outside [e_begin, e_begin + E) were never written: only the unsharded call may combine directly)
  MB_CHECK_ARG(E == n_experts_total && e_begin == 0, MB_ERR_SHAPE,
               "mb_moe_ffn: expert-parallel slabs combine through mb_ep_combine_push / mb_moe_combine(y_partial)");
  return mb_moe_combine(out_pairs, weights, shared, residual, y, nullptr, nullptr, T, k, D, stream_);
}

extern "C" int mb_moe_peer_area_bytes(int G, int Tmax, int D, int64_t* bytes) {
  MB_CHECK_ARG(G >= 1 && Tmax >= 1 && D >= 1 && bytes != nullptr, MB_ERR_SHAPE, "mb_moe_peer_area_bytes: bad shape");
  *bytes = static_cast<int64_t>(PeerArea::slot_floats(G, Tmax, D) * 4 + (static_cast<size_t>(G) + 3) * 4);
  return MB_OK;
}

extern "C" int mb_moe_combine_push(const void* out_pairs, const float* weights, const int32_t* pair_row,
                                   float* const* peers, int my_rank, int G, int T, int Tmax, int k, int D,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_combine_push: no sm_100 device");
  MB_CHECK_ARG(G >= 1 && my_rank >= 0 && my_rank < G && T >= 1 && T <= Tmax && peers != nullptr, MB_ERR_SHAPE,
               "mb_moe_combine_push: bad ranks / rows (T=%d Tmax=%d G=%d)", T, Tmax, G);
  const int64_t total = static_cast<int64_t>(T) * D;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms()) grid = num_sms();
  moe_combine_push_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(out_pairs), weights, pair_row,
                                                    peers, my_rank, G, T, Tmax, k, D);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_moe_reduce_finalize(float* const* peers, int my_rank, int G, int T, int Tmax, int D,
                                      const void* shared, const void* residual, void* y, uint32_t* fin_done,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_reduce_finalize: no sm_100 device");
  MB_CHECK_ARG(G >= 1 && G <= 256 && my_rank >= 0 && my_rank < G && T >= 1 && T <= Tmax && peers != nullptr &&
                   fin_done != nullptr,
               MB_ERR_SHAPE, "mb_moe_reduce_finalize: bad ranks / rows");
  const int64_t total = static_cast<int64_t>(T) * D;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms()) grid = num_sms();
  moe_reduce_finalize_kernel<<<grid, 256, 0, stream>>>(peers, my_rank, G, T, Tmax, D,
                                                       static_cast<const __nv_bfloat16*>(shared),
                                                       static_cast<const __nv_bfloat16*>(residual),
                                                       static_cast<__nv_bfloat16*>(y), fin_done);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_moe_plan(const int32_t* idx, int32_t* pair_row, int32_t* row_token, int32_t* tile_expert,
                           int32_t* meta, int32_t* counts, int T, int k, int E, int e_begin, int div, int granule,
                           int max_rows, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_plan: no sm_100 device");
  MB_CHECK_ARG(E >= 1 && E <= 4096 && e_begin >= 0 && k >= 1 && T >= 0 && div >= 1, MB_ERR_SHAPE,
               "mb_moe_plan: bad bucket range");
  MB_CHECK_ARG(granule == 1 || granule == 128, MB_ERR_SHAPE, "mb_moe_plan: granule must be 1 (dispatch) or 128 (GEMM tiles)");
  MB_CHECK_ARG(granule == 128 || tile_expert == nullptr, MB_ERR_SHAPE, "mb_moe_plan: tile_expert needs granule 128");
  // worst case: every bucket wastes granule - 1 padding rows
  int64_t need = static_cast<int64_t>(T) * k + static_cast<int64_t>(E) * (granule - 1);
  need = ((need + granule - 1) / granule) * granule;
  MB_CHECK_ARG(max_rows % granule == 0 && max_rows >= need, MB_ERR_SHAPE,
               "mb_moe_plan: max_rows=%d must be a multiple of %d and >= %ld", max_rows, granule, (long)need);
  const size_t smem = (2 * static_cast<size_t>(E) + 1) * sizeof(int32_t);
  moe_plan_kernel<<<1, kPlanThreads, smem, stream>>>(idx, pair_row, row_token, tile_expert, meta, counts, T * k, k, E,
                                                     e_begin, div, granule, max_rows);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_moe_gather_rows(const void* x, const int32_t* row_token, const int32_t* meta, void* out, int max_rows,
                                  int D, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_gather_rows: no sm_100 device");
  MB_CHECK_ARG(D % 8 == 0 && max_rows >= 0, MB_ERR_SHAPE, "mb_moe_gather_rows: D must be a multiple of 8");
  if (max_rows == 0) return MB_OK;
  int64_t blocks = (static_cast<int64_t>(max_rows) * (D / 8) + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  moe_gather_rows_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), row_token, meta, static_cast<__nv_bfloat16*>(out), max_rows, D);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_moe_finalize(const float* y_sum, const void* shared, const void* residual, void* y, int T, int D,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_finalize: no sm_100 device");
  const int64_t total = static_cast<int64_t>(T) * D;
  if (total == 0) return MB_OK;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  moe_finalize_kernel<<<grid, 256, 0, stream>>>(y_sum, static_cast<const __nv_bfloat16*>(shared),
                                                static_cast<const __nv_bfloat16*>(residual),
                                                static_cast<__nv_bfloat16*>(y), total);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_moe_combine(const void* out_pairs, const float* weights, const void* shared, const void* residual,
                              void* y, float* y_partial, const int32_t* pair_row, int T, int k, int D, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_moe_combine: no sm_100 device");
  const int64_t total = static_cast<int64_t>(T) * D;
  if (total == 0) return MB_OK;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  MB_CHECK_CUDA(launch_pdl(moe_combine_kernel, dim3(grid), dim3(256), 0, stream,
                           static_cast<const __nv_bfloat16*>(out_pairs), weights, static_cast<const __nv_bfloat16*>(shared),
                           static_cast<const __nv_bfloat16*>(residual), static_cast<__nv_bfloat16*>(y), y_partial, pair_row,
                           T, k, D));
  return MB_OK;
}
