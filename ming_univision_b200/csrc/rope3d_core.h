// 3-D multimodal RoPE (M-RoPE) + KV-cache append: the config-gated `rope_scaling.type == "3D"` variant of SURVEY.md row
// a14 — BailingMoe3DRotaryEmbedding.forward (mingunivision/modeling_bailing_moe.py:413-425) and
// apply_multimodal_rotary_pos_emb (:463-469).  Per-thread body as __host__ __device__ code, shared by the kernel
// (rope3d.cu) and by the CPU emulation in tests/native/rope3d_emu.cpp.
//
// Numerics of the reference for this variant (they differ from the 1-D legacy path): cos / sin stay fp32 (the rotary
// module runs with autocast disabled and never casts down), so q * cos promotes to fp32 and the rotated q / k are
// rounded to bf16 ONCE, when the attention casts them to its compute dtype (:946-975).  Every product and the sum are
// separately rounded fp32 operations (no FMA contraction): __fmul_rn / __fadd_rn on the device.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MBR_HD __host__ __device__ __forceinline__
#else
#define MBR_HD static inline
#endif

namespace mbrope {

MBR_HD float bf16_bits_to_float(uint16_t bits) {
  union {
    uint32_t u;
    float f;
  } c;
  c.u = static_cast<uint32_t>(bits) << 16;
  return c.f;
}

// Round to nearest even, NaN kept quiet (the layout of __float2bfloat16_rn).
MBR_HD uint16_t float_to_bf16_bits(float f) {
  union {
    uint32_t u;
    float f;
  } c;
  c.f = f;
  if ((c.u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((c.u >> 16) | 0x0040u);
  const uint32_t lsb = (c.u >> 16) & 1u;
  return static_cast<uint16_t>((c.u + 0x7fffu + lsb) >> 16);
}

MBR_HD float r_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
MBR_HD float r_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}

// Position component (0 temporal, 1 height, 2 width) that frequency index i in [0, hd/2) takes its angle from: the
// head dimension is cut into sections [s0, s1, s2, s0, s1, s2] and section j uses component j % 3; dims i and
// i + hd/2 fall into sections with the same component.
MBR_HD int mrope_component(int i, int s0, int s1) { return i < s0 ? 0 : (i < s0 + s1 ? 1 : 2); }

// One (row, head, frequency) item.  qkv row layout: [H q heads | Hkv k heads | Hkv v heads] x hd, as mb_rope_kv_append.
// position_ids3: int32 [3, B*S].  Caches: [B, Hkv, Tmax, hd]; the row's slot is `slot`.
MBR_HD void rope3d_item(const uint16_t* qkv, const int32_t* position_ids3, uint16_t* q_out, uint16_t* kcache,
                        uint16_t* vcache, int64_t rows, int S, int H, int Hkv, int hd, int Tmax, int slot, float theta,
                        int s0, int s1, int64_t row, int head, int i) {
  const int nheads = H + 2 * Hkv, half = hd / 2;
  const int64_t b = row / S;
  const uint16_t* src = qkv + (row * nheads + head) * hd;
  uint16_t* dst;
  if (head < H) {
    dst = q_out + (row * H + head) * hd;
  } else if (head < H + Hkv) {
    dst = kcache + ((b * Hkv + (head - H)) * Tmax + slot) * hd;
  } else {
    dst = vcache + ((b * Hkv + (head - H - Hkv)) * Tmax + slot) * hd;
    dst[i] = src[i];
    dst[i + half] = src[i + half];
    return;
  }
  const int pos = position_ids3[mrope_component(i, s0, s1) * rows + row];
  // inv_freq = 1 / base^(2i/hd) in fp32, angle = one fp32 product (:414-419)
  const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * i) / static_cast<float>(hd));
  const float ang = r_mul(inv_freq, static_cast<float>(pos));
  const float c = cosf(ang), sn = sinf(ang);
  const float x1 = bf16_bits_to_float(src[i]), x2 = bf16_bits_to_float(src[i + half]);
  // q * cos + rotate_half(q) * sin, rotate_half = cat(-x2, x1) (:428-433, :467-468)
  dst[i] = float_to_bf16_bits(r_add(r_mul(x1, c), r_mul(-x2, sn)));
  dst[i + half] = float_to_bf16_bits(r_add(r_mul(x2, c), r_mul(x1, sn)));
}

}  // namespace mbrope
