// mb_rope3d_kv_append: 3-D multimodal RoPE + KV-cache append (config-gated M-RoPE variant of SURVEY.md row a14).
// One CTA per token row, one warp per head, lanes over the hd/2 frequencies — the thread mapping of the 1-D kernel
// (rope_kv_append_kernel in llm.cu); the arithmetic is rope3d_core.h, shared with the CPU emulation.
#include "common.h"
#include "rope3d_core.h"

namespace {

__global__ void __launch_bounds__(256)
rope3d_kv_append_kernel(const uint16_t* __restrict__ qkv, const int32_t* __restrict__ position_ids3,
                        uint16_t* __restrict__ q_out, uint16_t* __restrict__ kcache, uint16_t* __restrict__ vcache,
                        int B, int S, int H, int Hkv, int hd, int Tmax, const int32_t* __restrict__ t_dev, int t_host,
                        float theta, int s0, int s1) {
  const int64_t row = blockIdx.x;  // b * S + s
  const int slot = (t_dev ? *t_dev : 0) + t_host + static_cast<int>(row % S);
  const int nheads = H + 2 * Hkv, half = hd / 2;
  for (int head = threadIdx.x >> 5; head < nheads; head += blockDim.x >> 5)
    for (int i = threadIdx.x & 31; i < half; i += 32)
      mbrope::rope3d_item(qkv, position_ids3, q_out, kcache, vcache, static_cast<int64_t>(B) * S, S, H, Hkv, hd, Tmax,
                          slot, theta, s0, s1, row, head, i);
}

}  // namespace

using namespace mb;

extern "C" int mb_rope3d_kv_append(const void* qkv, const int32_t* position_ids3, void* q_out, void* kcache,
                                   void* vcache, int B, int S, int H, int Hkv, int hd, int Tmax, const int32_t* t_dev,
                                   int t_host, float rope_theta, int sec0, int sec1, int sec2, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, "mb_rope3d_kv_append: no sm_100 device");
  MB_CHECK_ARG(B >= 0 && S >= 0 && H >= 1 && Hkv >= 1 && hd % 2 == 0, MB_ERR_SHAPE, "mb_rope3d_kv_append: bad shape");
  MB_CHECK_ARG(sec0 >= 0 && sec1 >= 0 && sec2 >= 0 && 2 * (sec0 + sec1 + sec2) == hd, MB_ERR_SHAPE,
               "mb_rope3d_kv_append: mrope sections (%d, %d, %d) must sum to head_dim / 2 = %d", sec0, sec1, sec2,
               hd / 2);
  MB_CHECK_ARG(t_dev != nullptr || (t_host >= 0 && t_host + S <= Tmax), MB_ERR_SHAPE,
               "mb_rope3d_kv_append: cache overflow (t=%d S=%d Tmax=%d)", t_host, S, Tmax);
  if (static_cast<int64_t>(B) * S == 0) return MB_OK;
  rope3d_kv_append_kernel<<<B * S, 256, 0, stream>>>(
      static_cast<const uint16_t*>(qkv), position_ids3, static_cast<uint16_t*>(q_out),
      static_cast<uint16_t*>(kcache), static_cast<uint16_t*>(vcache), B, S, H, Hkv, hd, Tmax, t_dev, t_host, rope_theta,
      sec0, sec1);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}
