// Expert-parallel MoE exchange over NVLink peer memory — dispatch AND combine, no NCCL, no host involvement, CUDA-graph
// capturable (the reference keeps all 64 experts on one device, modeling_bailing_moe.py:543-549, and reads
// the per-expert counts on the host in every layer, :616 — there is nothing to mirror).
//
// Layout of the job ("data parallel x expert parallel"): every rank r of G works on ITS OWN T token rows (its own
// image / request) with replicated attention, gates, shared experts and norms, and owns the routed experts
// [r E/G, (r+1) E/G).  Per MoE layer:
//
//   mb_ep_dispatch_push     rank r stores its T rows  x [T, D] bf16, idx [T, k] i32 (global expert ids), w [T, k] f32
//                           into EVERY peer's exchange area at row block r (plain st.global on peer-mapped pointers: the
//                           stores cross NVLink while the kernel runs), then the last CTA issues fence.sys and raises
//                           flag_disp[r] = epoch in every peer's area.  The dispatch is an all-gather, not a routed
//                           all-to-all, on purpose: T is the number of CFG rows (<= 8) at decode and a prompt at prefill,
//                           so the rows are 12 KB .. a few 100 KB — one hop of latency either way, and no count exchange.
//   mb_ep_dispatch_wait     spins (ld.acquire.sys, wall-clock bounded) until the G flags of the LOCAL area show the epoch.
//   ... the expert kernels (mb_moe_sort / gate_up / down, or mb_moe_plan / grouped GEMMs) run on the gathered
//       [G*T, D] rows, restricted to the local experts ...
//   mb_ep_combine_push      fp32 partial sums  sum_{j local} w[t,j] * out_pairs[row(t,j)]  of ALL G*T rows; the rows of
//                           source rank q are stored straight into q's area at slot [r]; flag_comb[r] = epoch on all peers.
//   mb_ep_reduce_finalize   waits for the G combine flags, adds the G slots in rank order (deterministic), applies the
//                           reference's rounding chain bf16(bf16(bf16(sum) + shared) + residual) (:632-638, :604-605,
//                           :1226) and advances the epoch.
//
// One exchange area per rank, SINGLE-buffered.  Why that is safe: rank A starts call n+1 (overwriting its row block in
// B's area) only after its own reduce_finalize(n), which waited for B's combine flag of call n — raised by the last CTA
// of B's combine_push(n), i.e. after every read B makes of the dispatched rows of call n.  And A overwrites its combine
// slot in B's area (combine_push(n+1)) only after its dispatch_wait(n+1) saw B's dispatch flag n+1, which B raises after
// its reduce_finalize(n) in stream order.  None of these kernels uses programmatic dependent launch, so stream order is
// completion order.  Flags carry the monotonically increasing epoch (a device-side counter), so nothing is ever reset
// and a CUDA graph of a whole token step replays correctly.
//
// A peer that never arrives (crashed / diverged process) must not hang the GPU: the waits are bounded by a wall-clock
// budget (%globaltimer; MB_EP_TIMEOUT_MS, default 20 s).  On expiry the kernel records an error code in the area's
// control block and carries on (its output is garbage); the host reads the code from the control block (ep.PeerDispatch.check,
// called once per request by the model's entry points) and raises — the CUDA context survives, unlike with __trap().
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mb {

struct EpLayout {
  // byte offsets inside one rank's area
  size_t x, idx, w, part, ctrl, total;
  __host__ __device__ EpLayout(int G, int Tmax, int D, int k) {
    auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
    x = 0;
    idx = up(x + static_cast<size_t>(G) * Tmax * D * 2);
    w = up(idx + static_cast<size_t>(G) * Tmax * k * 4);
    part = up(w + static_cast<size_t>(G) * Tmax * k * 4);
    ctrl = up(part + static_cast<size_t>(G) * Tmax * D * 4);
    total = up(ctrl + (2 * static_cast<size_t>(G) + 8) * 4);
  }
};
// control block (u32): flag_disp[G], flag_comb[G], epoch, done_a, done_b, done_c, err, err_peer
__device__ __forceinline__ uint32_t* ep_ctrl(uint8_t* area, const EpLayout& L) {
  return reinterpret_cast<uint32_t*>(area + L.ctrl);
}
enum { kEpEpoch = 0, kEpDoneA = 1, kEpDoneB = 2, kEpDoneC = 3, kEpErr = 4, kEpErrPeer = 5 };

__device__ __forceinline__ void ep_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ep_ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ep_globaltimer() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Bounded wait of thread `src` (< G) for flags[src] >= epoch.  Returns false on timeout (and records it).
__device__ __forceinline__ bool ep_wait_flag(const uint32_t* flags, int src, uint32_t epoch, uint32_t* ctrl_tail,
                                             uint64_t timeout_ns, int code) {
  const uint64_t t0 = ep_globaltimer();
  uint32_t spins = 0;
  while (static_cast<int32_t>(ep_ld_acquire_sys(flags + src) - epoch) < 0) {
    __nanosleep(32);
    if ((++spins & 1023u) == 0 && ep_globaltimer() - t0 > timeout_ns) {
      atomicExch(ctrl_tail + kEpErr, static_cast<uint32_t>(code));
      atomicExch(ctrl_tail + kEpErrPeer, static_cast<uint32_t>(src));
      return false;
    }
  }
  return true;
}

// grid = (blocks per peer, G peers)
__global__ void __launch_bounds__(256)
ep_dispatch_push_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ idx,
                        const float* __restrict__ w, uint8_t* const* __restrict__ peers, int my_rank, int G, int T,
                        int Tmax, int D, int k) {
  const EpLayout L(G, Tmax, D, k);
  uint8_t* local = peers[my_rank];
  uint32_t* ctl = ep_ctrl(local, L) + 2 * G;
  const uint32_t epoch = ctl[kEpEpoch] + 1;
  uint8_t* dst = peers[blockIdx.y];
  // rows of source rank r sit at row block r of a PACKED [G*T, .] array (T is the same on every rank for one call)
  const int vec_x = T * D / 8;
  const uint4* sx = reinterpret_cast<const uint4*>(x);
  uint4* dx = reinterpret_cast<uint4*>(dst + L.x) + static_cast<size_t>(my_rank) * vec_x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < vec_x; i += gridDim.x * blockDim.x) dx[i] = sx[i];
  if (blockIdx.x == 0) {
    int32_t* di = reinterpret_cast<int32_t*>(dst + L.idx) + static_cast<size_t>(my_rank) * T * k;
    float* dw = reinterpret_cast<float*>(dst + L.w) + static_cast<size_t>(my_rank) * T * k;
    for (int i = threadIdx.x; i < T * k; i += blockDim.x) {
      di[i] = idx[i];
      dw[i] = w[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = atomicAdd(&ctl[kEpDoneA], 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    for (int pr = threadIdx.x; pr < G; pr += blockDim.x)
      ep_st_release_sys(ep_ctrl(peers[pr], L) + my_rank, epoch);  // flag_disp[my_rank] in peer pr's area
    if (threadIdx.x == 0) ctl[kEpDoneA] = 0;
  }
}

__global__ void __launch_bounds__(32)
ep_dispatch_wait_kernel(uint8_t* const* __restrict__ peers, int my_rank, int G, int Tmax, int D, int k,
                        uint64_t timeout_ns) {
  const EpLayout L(G, Tmax, D, k);
  uint32_t* flags = ep_ctrl(peers[my_rank], L);
  uint32_t* ctl = flags + 2 * G;
  const uint32_t epoch = ctl[kEpEpoch] + 1;
  for (int s = threadIdx.x; s < G; s += 32) ep_wait_flag(flags, s, epoch, ctl, timeout_ns, 1);
}

// mb_ep_dispatch_wait fused with the counting sort of the gathered pairs by LOCAL expert (mb_moe_sort's job on the
// streaming path): one launch less per layer.  Single CTA; pairs routed to other ranks' experts are not listed.
__global__ void __launch_bounds__(256)
ep_wait_sort_kernel(uint8_t* const* __restrict__ peers, int my_rank, int G, int T, int Tmax, int D, int k,
                    int32_t* __restrict__ expert_offsets, int32_t* __restrict__ sorted_pair, int E, int e_begin,
                    uint64_t timeout_ns) {
  extern __shared__ int32_t sm[];  // counts[E], cursor[E], idx[G*T*k] (local expert index of every pair, or -1)
  const EpLayout L(G, Tmax, D, k);
  uint32_t* flags = ep_ctrl(peers[my_rank], L);
  uint32_t* ctl = flags + 2 * G;
  const uint32_t epoch = ctl[kEpEpoch] + 1;
  if (threadIdx.x < G) ep_wait_flag(flags, threadIdx.x, epoch, ctl, timeout_ns, 1);
  __syncthreads();
  const int32_t* idx = reinterpret_cast<const int32_t*>(peers[my_rank] + L.idx);
  const int npairs = G * T * k;
  int32_t* counts = sm;
  int32_t* cursor = sm + E;
  int32_t* loc = sm + 2 * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) counts[e] = 0;
  __syncthreads();
  // one coalesced pass over the gathered ids (written by the peers: read through L2), everything else in shared memory
  for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
    int e = __ldcg(idx + p) - e_begin;
    if (e < 0 || e >= E) e = -1;
    loc[p] = e;
    if (e >= 0) atomicAdd(&counts[e], 1);
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
  for (int e = threadIdx.x; e < E; e += blockDim.x) {  // deterministic order inside an expert: increasing pair index
    int c = cursor[e];
    if (counts[e] == 0) continue;
    for (int p = 0; p < npairs; ++p)
      if (loc[p] == e) sorted_pair[c++] = p;
  }
}

// grid-stride over the G*T*D/4 float4 elements of the partial sums
__global__ void __launch_bounds__(256)
ep_combine_push_kernel(const __nv_bfloat16* __restrict__ out_pairs, const int32_t* __restrict__ pair_row,
                       uint8_t* const* __restrict__ peers, int my_rank, int G, int T, int Tmax, int D, int k,
                       int e_begin, int e_local) {
  const EpLayout L(G, Tmax, D, k);
  uint8_t* local = peers[my_rank];
  uint32_t* ctl = ep_ctrl(local, L) + 2 * G;
  const uint32_t epoch = ctl[kEpEpoch] + 1;
  const int32_t* idx_all = reinterpret_cast<const int32_t*>(local + L.idx);
  const float* w_all = reinterpret_cast<const float*>(local + L.w);
  const int d4 = D / 4;
  const int64_t total = static_cast<int64_t>(G) * T * d4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / d4);  // gathered row: source rank r / T, its token r % T
    const int c = static_cast<int>(i % d4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < k; ++j) {
      const int p = r * k + j;
      int row;
      if (pair_row != nullptr) {
        row = pair_row[p];  // grouped layout: < 0 for pairs of other ranks' experts
      } else {
        const int e = idx_all[p] - e_begin;
        row = (e >= 0 && e < e_local) ? p : -1;
      }
      if (row >= 0) {
        const float wj = w_all[p];
        const uint2 v = *reinterpret_cast<const uint2*>(out_pairs + static_cast<int64_t>(row) * D + c * 4);
        const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
        acc.x += wj * a.x; acc.y += wj * a.y; acc.z += wj * b.x; acc.w += wj * b.y;
      }
    }
    const int q = r / T, t = r % T;
    float4* dst = reinterpret_cast<float4*>(peers[q] + L.part) + (static_cast<size_t>(my_rank) * T + t) * d4 + c;
    *dst = acc;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = atomicAdd(&ctl[kEpDoneB], 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    for (int pr = threadIdx.x; pr < G; pr += blockDim.x)
      ep_st_release_sys(ep_ctrl(peers[pr], L) + G + my_rank, epoch);  // flag_comb[my_rank] in peer pr's area
    if (threadIdx.x == 0) ctl[kEpDoneB] = 0;
  }
}

__global__ void __launch_bounds__(256)
ep_reduce_finalize_kernel(uint8_t* const* __restrict__ peers, int my_rank, int G, int T, int Tmax, int D, int k,
                          const __nv_bfloat16* __restrict__ shared, const __nv_bfloat16* __restrict__ residual,
                          __nv_bfloat16* __restrict__ y, uint64_t timeout_ns) {
  const EpLayout L(G, Tmax, D, k);
  uint8_t* local = peers[my_rank];
  uint32_t* flags = ep_ctrl(local, L);
  uint32_t* ctl = flags + 2 * G;
  const uint32_t epoch = ctl[kEpEpoch] + 1;
  if (threadIdx.x < G) ep_wait_flag(flags + G, threadIdx.x, epoch, ctl, timeout_ns, 2);
  __syncthreads();
  const float4* part = reinterpret_cast<const float4*>(local + L.part);
  const int d4 = D / 4;
  const int64_t total = static_cast<int64_t>(T) * d4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < G; ++r) {  // rank order: the sum is reproducible
      const float4 v = __ldcg(part + static_cast<size_t>(r) * T * d4 + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float o[4] = {bf16_round(s.x), bf16_round(s.y), bf16_round(s.z), bf16_round(s.w)};
    if (shared != nullptr) {
      const uint2 v = *reinterpret_cast<const uint2*>(shared + i * 4);
      const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
      o[0] = bf16_round(o[0] + a.x); o[1] = bf16_round(o[1] + a.y);
      o[2] = bf16_round(o[2] + b.x); o[3] = bf16_round(o[3] + b.y);
    }
    if (residual != nullptr) {
      const uint2 v = *reinterpret_cast<const uint2*>(residual + i * 4);
      const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
      o[0] += a.x; o[1] += a.y; o[2] += b.x; o[3] += b.y;
    }
    uint2 out;
    out.x = pack_bf16x2(o[0], o[1]);
    out.y = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(y + i * 4) = out;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl[kEpDoneC], 1u) == gridDim.x - 1) {
      ctl[kEpDoneC] = 0;
      ctl[kEpEpoch] = epoch;  // every CTA has read the epoch and its slots by now
    }
  }
}

static uint64_t ep_timeout_ns() {
  static uint64_t v = 0;
  if (v == 0) {
    const char* e = getenv("MB_EP_TIMEOUT_MS");
    const double ms = (e != nullptr && atof(e) > 0.0) ? atof(e) : 20000.0;
    v = static_cast<uint64_t>(ms * 1e6);
  }
  return v;
}

}  // namespace mb

using namespace mb;

#define MB_EP_COMMON_CHECKS(name)                                                                                      \
  MB_CHECK_ARG(mb_device_ok(), MB_ERR_ARCH, name ": no sm_100 device");                                                \
  MB_CHECK_ARG(G >= 1 && G <= 32 && my_rank >= 0 && my_rank < G && T >= 1 && T <= Tmax && k >= 1 && D >= 8 &&          \
                   D % 8 == 0 && peers != nullptr,                                                                     \
               MB_ERR_SHAPE, name ": bad shape (G=%d rank=%d T=%d Tmax=%d D=%d k=%d)", G, my_rank, T, Tmax, D, k)

extern "C" int mb_ep_area_layout(int G, int Tmax, int D, int k, int64_t* offsets6) {
  MB_CHECK_ARG(G >= 1 && Tmax >= 1 && D >= 8 && D % 8 == 0 && k >= 1 && offsets6 != nullptr, MB_ERR_SHAPE,
               "mb_ep_area_layout: bad shape");
  const EpLayout L(G, Tmax, D, k);
  offsets6[0] = static_cast<int64_t>(L.x);
  offsets6[1] = static_cast<int64_t>(L.idx);
  offsets6[2] = static_cast<int64_t>(L.w);
  offsets6[3] = static_cast<int64_t>(L.part);
  offsets6[4] = static_cast<int64_t>(L.ctrl);
  offsets6[5] = static_cast<int64_t>(L.total);
  return MB_OK;
}

extern "C" int mb_ep_dispatch_push(const void* x, const int32_t* idx, const float* w, void* const* peers, int my_rank,
                                   int G, int T, int Tmax, int D, int k, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_EP_COMMON_CHECKS("mb_ep_dispatch_push");
  MB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, MB_ERR_ALIGN, "mb_ep_dispatch_push: x must be 16-byte aligned");
  const int vec = T * D / 8;
  int nblk = (vec + 256 * 4 - 1) / (256 * 4);
  nblk = nblk < 1 ? 1 : (nblk > 32 ? 32 : nblk);
  ep_dispatch_push_kernel<<<dim3(nblk, G), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), idx, w,
                                                             reinterpret_cast<uint8_t* const*>(peers), my_rank, G, T,
                                                             Tmax, D, k);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_ep_dispatch_wait(void* const* peers, int my_rank, int G, int Tmax, int D, int k, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = 1;
  MB_EP_COMMON_CHECKS("mb_ep_dispatch_wait");
  ep_dispatch_wait_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<uint8_t* const*>(peers), my_rank, G, Tmax, D, k,
                                                ep_timeout_ns());
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_ep_wait_sort(void* const* peers, int my_rank, int G, int T, int Tmax, int D, int k,
                               int32_t* expert_offsets, int32_t* sorted_pair, int E, int e_begin, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_EP_COMMON_CHECKS("mb_ep_wait_sort");
  MB_CHECK_ARG(E >= 1 && E <= 4096 && expert_offsets != nullptr && sorted_pair != nullptr, MB_ERR_SHAPE,
               "mb_ep_wait_sort: bad expert range");
  const size_t smem = (2 * static_cast<size_t>(E) + static_cast<size_t>(G) * T * k) * sizeof(int32_t);
  MB_CHECK_ARG(smem <= 48 * 1024, MB_ERR_SHAPE, "mb_ep_wait_sort: %d gathered pairs do not fit shared memory", G * T * k);
  ep_wait_sort_kernel<<<1, 256, smem, stream>>>(reinterpret_cast<uint8_t* const*>(peers), my_rank, G, T,
                                                                    Tmax, D, k, expert_offsets, sorted_pair, E, e_begin,
                                                                    ep_timeout_ns());
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_ep_combine_push(const void* out_pairs, const int32_t* pair_row, void* const* peers, int my_rank,
                                  int G, int T, int Tmax, int D, int k, int e_begin, int e_local, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_EP_COMMON_CHECKS("mb_ep_combine_push");
  const int64_t total = static_cast<int64_t>(G) * T * (D / 4);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 4) grid = num_sms() * 4;
  ep_combine_push_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(out_pairs), pair_row,
                                                   reinterpret_cast<uint8_t* const*>(peers), my_rank, G, T, Tmax, D, k,
                                                   e_begin, e_local);
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}

extern "C" int mb_ep_reduce_finalize(void* const* peers, int my_rank, int G, int T, int Tmax, int D, int k,
                                     const void* shared, const void* residual, void* y, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MB_EP_COMMON_CHECKS("mb_ep_reduce_finalize");
  MB_CHECK_ARG(y != nullptr, MB_ERR_SHAPE, "mb_ep_reduce_finalize: null output");
  const int64_t total = static_cast<int64_t>(T) * (D / 4);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 2) grid = num_sms() * 2;
  ep_reduce_finalize_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<uint8_t* const*>(peers), my_rank, G, T, Tmax, D,
                                                      k, static_cast<const __nv_bfloat16*>(shared),
                                                      static_cast<const __nv_bfloat16*>(residual),
                                                      static_cast<__nv_bfloat16*>(y), ep_timeout_ns());
  MB_CHECK_CUDA(cudaGetLastError());
  return MB_OK;
}
