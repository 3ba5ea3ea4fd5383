// CPU emulation of rope3d_kv_append_kernel (ming_univision_b200/csrc/rope3d.cu): the same per-thread body
// (rope3d_core.h) walked over the kernel's grid (one CTA per token row, 8 warps over the heads, lanes over the
// frequencies).  Test infrastructure only; built by tests/test_rope3d_cpu.py with g++ -ffp-contract=off.
#include "../../ming_univision_b200/csrc/rope3d_core.h"

extern "C" void emu_rope3d_kv_append(const uint16_t* qkv, const int32_t* position_ids3, uint16_t* q_out,
                                     uint16_t* kcache, uint16_t* vcache, int B, int S, int H, int Hkv, int hd, int Tmax,
                                     int t_host, float theta, int s0, int s1) {
  const int nheads = H + 2 * Hkv, half = hd / 2;
  for (long long row = 0; row < static_cast<long long>(B) * S; ++row) {  // blockIdx.x
    const int slot = t_host + static_cast<int>(row % S);
    for (int tid = 0; tid < 256; ++tid)                                   // threadIdx.x
      for (int head = tid >> 5; head < nheads; head += 8)
        for (int i = tid & 31; i < half; i += 32)
          mbrope::rope3d_item(qkv, position_ids3, q_out, kcache, vcache, static_cast<long long>(B) * S, S, H, Hkv, hd,
                              Tmax, slot, theta, s0, s1, row, head, i);
  }
}

extern "C" unsigned emu_float_to_bf16_bits(float f) { return mbrope::float_to_bf16_bits(f); }
