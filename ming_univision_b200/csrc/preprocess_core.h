// Arithmetic and launch geometry of the image pre- / post-processing kernels (preprocess.cu), written once as
// __host__ __device__ code: the kernels call these functions per thread, and tests/native/preprocess_emu.cpp compiles
// the SAME header with g++ and walks the same grids on the CPU, so every index computation and every rounding step of
// the device path is checked against Pillow / torchvision without a GPU (the GPU tests then check the launch itself).
//
// What is restated (third-party code the reference calls; see oracle/preprocess_oracle.py for versions and call sites):
//   Pillow  src/libImaging/Resample.c : bicubic_filter, precompute_coeffs, normalize_coeffs_8bpc,
//                                        ImagingResampleHorizontal_8bpc / Vertical_8bpc (RGB, 8 bits per channel)
//   torchvision  to_tensor (u8 / 255), normalize ((x - mean) / std), to_pil_image (mul(255).byte())
// Reference call sites: mingtok/utils/processor.py:17-27, mingunivision/processing_bailingmm.py:80-123,
// mingunivision/modeling_bailing_moe.py:84-90.
//
// Bit-exactness rules: the weights are computed in IEEE double with every product and sum rounded separately (Pillow is
// compiled for baseline x86-64: no FMA), so the device side uses the __d*_rn / __f*_rn intrinsics, which nvcc never
// contracts; the host side must be compiled with -ffp-contract=off.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MBP_HD __host__ __device__ __forceinline__
#else
#define MBP_HD static inline
#endif

namespace mbpre {

constexpr int kPrecisionBits = 22;  // Resample.c PRECISION_BITS = 32 - 8 - 2
constexpr int kMaxSmemBytes = 48 * 1024;

// ---- separately rounded IEEE steps
MBP_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
MBP_HD double d_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
MBP_HD double d_sub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
MBP_HD double d_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
MBP_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
MBP_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
MBP_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
MBP_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

// ---- Pillow's filter and coefficient tables

// Keys cubic, a = -0.5, in Resample.c's operation order.
MBP_HD double bicubic_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return d_add(d_mul(d_mul(d_sub(d_mul(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0) return d_mul(d_sub(d_mul(d_add(d_mul(d_sub(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

struct AxisGeom {
  double scale;    // input pixels per output pixel
  double support;  // filter half-width in input pixels (2 * max(scale, 1): antialiasing when shrinking)
  double ss;       // 1 / max(scale, 1)
  int ksize;       // row length of the coefficient table
};

MBP_HD AxisGeom axis_geom(int in_size, int out_size) {
  AxisGeom g;
  g.scale = d_div(static_cast<double>(static_cast<float>(in_size)), static_cast<double>(out_size));
  const double filterscale = g.scale < 1.0 ? 1.0 : g.scale;
  g.support = d_mul(2.0, filterscale);
  g.ksize = static_cast<int>(ceil(g.support)) * 2 + 1;
  g.ss = d_div(1.0, filterscale);
  return g;
}

// Window of output index xx: first input index, tap count, and the (fractional) centre.
MBP_HD void axis_window(const AxisGeom& g, int in_size, int xx, int* xmin_out, int* count_out, double* center_out) {
  const double center = d_mul(d_add(static_cast<double>(xx), 0.5), g.scale);
  int xmin = static_cast<int>(d_add(d_sub(center, g.support), 0.5));  // C cast: truncation, as in Pillow
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(d_add(d_add(center, g.support), 0.5));
  if (xmax > in_size) xmax = in_size;
  *xmin_out = xmin;
  *count_out = xmax - xmin;
  *center_out = center;
}

MBP_HD double axis_tap(const AxisGeom& g, int xmin, double center, int x) {
  return bicubic_filter(d_mul(d_add(d_sub(static_cast<double>(x + xmin), center), 0.5), g.ss));
}

// One row of the tables: bounds2 = (xmin, count), k[0 .. ksize) = normalised weights in 22-bit fixed point.
MBP_HD void axis_coeffs(const AxisGeom& g, int in_size, int xx, int32_t* bounds2, int32_t* k) {
  int xmin, count;
  double center;
  axis_window(g, in_size, xx, &xmin, &count, &center);
  double ww = 0.0;
  for (int x = 0; x < count; ++x) ww = d_add(ww, axis_tap(g, xmin, center, x));
  for (int x = 0; x < count; ++x) {
    double w = axis_tap(g, xmin, center, x);
    if (ww != 0.0) w = d_div(w, ww);
    const double scaled = d_mul(w, static_cast<double>(1 << kPrecisionBits));
    k[x] = w < 0.0 ? static_cast<int32_t>(d_add(-0.5, scaled)) : static_cast<int32_t>(d_add(0.5, scaled));
  }
  for (int x = count; x < g.ksize; ++x) k[x] = 0;
  bounds2[0] = xmin;
  bounds2[1] = count;
}

MBP_HD uint8_t clip8(int32_t acc) {
  const int32_t v = acc >> kPrecisionBits;  // arithmetic shift
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ToTensor + Normalize of one channel value, and its inverse with ToPILImage's truncation.
MBP_HD float normalize_u8(uint8_t v, float mean, float stdv) {
  return f_div(f_sub(f_div(static_cast<float>(v), 255.0f), mean), stdv);
}
MBP_HD uint8_t denormalize_to_u8(float x, float mean, float stdv) {
  const float y = f_mul(f_add(f_mul(x, stdv), mean), 255.0f);
  const int v = static_cast<int>(y);  // truncation toward zero
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---- launch geometry and workspace layout (host side; shared by the C entry point and the CPU emulation)

struct Plan {
  int n, in_h, in_w;          // source images [n, in_h, in_w, 3] u8
  int res_h, res_w;           // size after Resize
  int crop_top, crop_left;    // window of the resized image that is kept
  int out_h, out_w;           // = crop size
  int do_h, do_v;             // passes Pillow runs: only along an axis whose size changes
  int ksize_h, ksize_v;
  int row0, rows;             // input rows the horizontal pass produces: [row0, row0 + rows)
  int tile_w, tile_rows;      // output tile of one horizontal-pass CTA
  int smem_row_bytes;         // staged input row segment (plus <= 3 bytes of alignment slack), multiple of 16
  int64_t off_bounds_h, off_kk_h, off_bounds_v, off_kk_v, off_temp, total_bytes;
};

constexpr int kHThreads = 256;
constexpr int kVThreads = 128;

static inline int64_t align16(int64_t v) { return (v + 15) & ~static_cast<int64_t>(15); }

// Returns 0, or the number of the first violated precondition (the entry point turns it into MB_ERR_SHAPE).
static inline int make_plan(int n, int in_h, int in_w, int res_h, int res_w, int crop_top, int crop_left, int out_h,
                            int out_w, Plan* p) {
  if (n < 0 || in_h < 1 || in_w < 1 || res_h < 1 || res_w < 1 || out_h < 1 || out_w < 1) return 1;
  if (crop_top < 0 || crop_left < 0 || crop_top + out_h > res_h || crop_left + out_w > res_w) return 2;
  if (in_h > (1 << 24) || in_w > (1 << 24)) return 3;  // Pillow's box is float
  // Pillow >= 11 runs the VERTICAL pass first for images more than 100 times taller than wide whose height shrinks
  // (PIL/Image.py, Image.resize); the u8 rounding between the passes makes the order visible.  Not a photograph's
  // geometry: refused rather than answered differently from the library.
  if (static_cast<int64_t>(in_h) > static_cast<int64_t>(in_w) * 100 && res_h < in_h) return 5;
  p->n = n, p->in_h = in_h, p->in_w = in_w, p->res_h = res_h, p->res_w = res_w;
  p->crop_top = crop_top, p->crop_left = crop_left, p->out_h = out_h, p->out_w = out_w;
  p->do_h = res_w != in_w, p->do_v = res_h != in_h;
  const AxisGeom gh = axis_geom(in_w, res_w), gv = axis_geom(in_h, res_h);
  p->ksize_h = gh.ksize, p->ksize_v = gv.ksize;
  if (p->do_v) {
    int xmin, count;
    double c;
    axis_window(gv, in_h, crop_top, &xmin, &count, &c);
    p->row0 = xmin;
    axis_window(gv, in_h, crop_top + out_h - 1, &xmin, &count, &c);
    p->rows = xmin + count - p->row0;
  } else {
    p->row0 = crop_top, p->rows = out_h;
  }
  // horizontal tile: kHThreads output pixels per CTA, narrowed until the staged input segment fits in shared memory
  p->tile_rows = 4, p->tile_w = kHThreads / p->tile_rows;
  for (;;) {
    const int64_t span = static_cast<int64_t>(ceil((p->tile_w - 1) * gh.scale + 2.0 * gh.support)) + 3;
    const int64_t row_bytes = align16(span * 3 + 3);
    if (row_bytes * p->tile_rows <= kMaxSmemBytes) {
      p->smem_row_bytes = static_cast<int>(row_bytes);
      break;
    }
    if (p->tile_w > 1) p->tile_w /= 2;
    else if (p->tile_rows > 1) p->tile_rows /= 2;
    else return 4;  // one output pixel needs more than 48 KB of input: a > 2000-fold reduction
  }
  int64_t off = 0;
  p->off_bounds_h = off, off = align16(off + static_cast<int64_t>(out_w) * 2 * 4);
  p->off_kk_h = off, off = align16(off + static_cast<int64_t>(out_w) * p->ksize_h * 4);
  p->off_bounds_v = off, off = align16(off + static_cast<int64_t>(out_h) * 2 * 4);
  p->off_kk_v = off, off = align16(off + static_cast<int64_t>(out_h) * p->ksize_v * 4);
  p->off_temp = off;
  if (p->do_h) off = align16(off + static_cast<int64_t>(n) * p->rows * out_w * 3);
  p->total_bytes = off;
  return 0;
}

// ---- per-thread bodies

// Coefficient tables of the kept window: entry idx < out_w is column crop_left + idx, the others row crop_top + (idx - out_w).
MBP_HD void coeff_entry(const Plan& p, int idx, int32_t* bounds_h, int32_t* kk_h, int32_t* bounds_v, int32_t* kk_v) {
  if (idx < p.out_w) {
    const AxisGeom g = axis_geom(p.in_w, p.res_w);
    axis_coeffs(g, p.in_w, p.crop_left + idx, bounds_h + 2 * idx, kk_h + static_cast<int64_t>(idx) * g.ksize);
  } else if (idx < p.out_w + p.out_h) {
    const int j = idx - p.out_w;
    const AxisGeom g = axis_geom(p.in_h, p.res_h);
    axis_coeffs(g, p.in_h, p.crop_top + j, bounds_v + 2 * j, kk_v + static_cast<int64_t>(j) * g.ksize);
  }
}

// Copies nbytes from src to dst, where dst was placed so that (dst & 3) == (src & 3): aligned 32-bit words in the
// middle, single bytes at both ends; never touches a byte outside [src, src + nbytes).
MBP_HD void stage_bytes(uint8_t* dst, const uint8_t* src, int nbytes, int tid, int nthreads) {
  const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 3);
  int head = mis ? 4 - mis : 0;
  if (head > nbytes) head = nbytes;
  const int nwords = (nbytes - head) >> 2;
  const int tail0 = head + (nwords << 2);
  for (int i = tid; i < head; i += nthreads) dst[i] = src[i];
  const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src + head);
  uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
  for (int i = tid; i < nwords; i += nthreads) {
#if defined(__CUDA_ARCH__)
    d4[i] = __ldg(s4 + i);
#else
    d4[i] = s4[i];
#endif
  }
  for (int i = tail0 + tid; i < nbytes; i += nthreads) dst[i] = src[i];
}

struct HTile {
  int x0, x1;      // output columns (within the kept window) of this CTA
  int first;       // first input column staged
  int nbytes;      // bytes staged per row
};

MBP_HD HTile h_tile(const Plan& p, const int32_t* bounds_h, int bx) {
  HTile t;
  t.x0 = bx * p.tile_w;
  t.x1 = t.x0 + p.tile_w < p.out_w ? t.x0 + p.tile_w : p.out_w;
  t.first = bounds_h[2 * t.x0];
  t.nbytes = (bounds_h[2 * (t.x1 - 1)] + bounds_h[2 * (t.x1 - 1) + 1] - t.first) * 3;
  return t;
}

MBP_HD const uint8_t* src_px(const Plan& p, const uint8_t* src, int img, int row, int col) {
  return src + ((static_cast<int64_t>(img) * p.in_h + row) * p.in_w + col) * 3;
}

// Horizontal pass, phase 1: stage the input segments of this tile's rows in shared memory.
// Returns false when the segment does not fit (a violated bound of make_plan: the kernel traps).
MBP_HD bool h_phase_load(const Plan& p, const uint8_t* src, const int32_t* bounds_h, int bx, int by, int bz, int tid,
                         int nthreads, uint8_t* smem) {
  const HTile t = h_tile(p, bounds_h, bx);
  if (t.nbytes + 3 > p.smem_row_bytes) return false;
  for (int r = 0; r < p.tile_rows; ++r) {
    const int row = by * p.tile_rows + r;
    if (row >= p.rows) break;
    const uint8_t* g = src_px(p, src, bz, p.row0 + row, t.first);
    uint8_t* dst = smem + r * p.smem_row_bytes + (reinterpret_cast<uintptr_t>(g) & 3);
    stage_bytes(dst, g, t.nbytes, tid, nthreads);
  }
  return true;
}

// Horizontal pass, phase 2 (after the barrier): one output pixel (3 channels) per thread and step.
MBP_HD void h_phase_compute(const Plan& p, const uint8_t* src, const int32_t* bounds_h, const int32_t* kk_h, int bx,
                            int by, int bz, int tid, int nthreads, const uint8_t* smem, uint8_t* temp) {
  const HTile t = h_tile(p, bounds_h, bx);
  const int tw = t.x1 - t.x0;
  for (int item = tid; item < p.tile_rows * tw; item += nthreads) {
    const int r = item / tw, xl = t.x0 + item % tw;
    const int row = by * p.tile_rows + r;
    if (row >= p.rows) continue;
    const int xmin = bounds_h[2 * xl], count = bounds_h[2 * xl + 1];
    const uint8_t* g = src_px(p, src, bz, p.row0 + row, t.first);
    const uint8_t* px = smem + r * p.smem_row_bytes + (reinterpret_cast<uintptr_t>(g) & 3) + (xmin - t.first) * 3;
    const int32_t* k = kk_h + static_cast<int64_t>(xl) * p.ksize_h;
    int32_t a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < count; ++x) {
      const int32_t w = k[x];
      a0 += px[3 * x + 0] * w;
      a1 += px[3 * x + 1] * w;
      a2 += px[3 * x + 2] * w;
    }
    uint8_t* o = temp + ((static_cast<int64_t>(bz) * p.rows + row) * p.out_w + xl) * 3;
    o[0] = clip8(a0), o[1] = clip8(a1), o[2] = clip8(a2);
  }
}

// Vertical pass + crop for output pixel (img, yy, xl): the u8 channel values (v_pixel_u8: what Pillow's resize + the
// crop hold) and their ToTensor + Normalize (v_pixel).
// `in` is the horizontal pass's output (row pitch out_w, rows [row0, row0 + rows)) or, when Pillow skips that pass,
// the source itself (row pitch in_w, column offset crop_left).
MBP_HD void v_pixel_u8(const Plan& p, const uint8_t* src, const uint8_t* temp, const int32_t* bounds_v,
                       const int32_t* kk_v, int img, int yy, int xl, uint8_t u[3]) {
  const uint8_t* base;
  int64_t pitch;  // bytes per row
  int row_origin;
  if (p.do_h) {
    base = temp + (static_cast<int64_t>(img) * p.rows * p.out_w + xl) * 3;
    pitch = static_cast<int64_t>(p.out_w) * 3;
    row_origin = p.row0;
  } else {
    base = src + (static_cast<int64_t>(img) * p.in_h * p.in_w + p.crop_left + xl) * 3;
    pitch = static_cast<int64_t>(p.in_w) * 3;
    row_origin = 0;
  }
  uint8_t u0, u1, u2;
  if (p.do_v) {
    const int ymin = bounds_v[2 * yy], count = bounds_v[2 * yy + 1];
    const int32_t* k = kk_v + static_cast<int64_t>(yy) * p.ksize_v;
    const uint8_t* px = base + (ymin - row_origin) * pitch;
    int32_t a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < count; ++y, px += pitch) {
      const int32_t w = k[y];
      a0 += px[0] * w;
      a1 += px[1] * w;
      a2 += px[2] * w;
    }
    u0 = clip8(a0), u1 = clip8(a1), u2 = clip8(a2);
  } else {
    const uint8_t* px = base + (p.crop_top + yy - row_origin) * pitch;
    u0 = px[0], u1 = px[1], u2 = px[2];
  }
  u[0] = u0, u[1] = u1, u[2] = u2;
}

MBP_HD void v_pixel(const Plan& p, const uint8_t* src, const uint8_t* temp, const int32_t* bounds_v,
                    const int32_t* kk_v, int img, int yy, int xl, const float mean[3], const float stdv[3],
                    float out3[3]) {
  uint8_t u[3];
  v_pixel_u8(p, src, temp, bounds_v, kk_v, img, yy, xl, u);
  out3[0] = normalize_u8(u[0], mean[0], stdv[0]);
  out3[1] = normalize_u8(u[1], mean[1], stdv[1]);
  out3[2] = normalize_u8(u[2], mean[2], stdv[2]);
}

// ---- pixel-decoder tail fused with the u8 conversion (unpatchify -> clamp(-1, 1) -> tensor_to_pil)

MBP_HD float bf16_bits_to_float(uint16_t bits) {
  union {
    uint32_t u;
    float f;
  } cvt;
  cvt.u = static_cast<uint32_t>(bits) << 16;
  return cvt.f;
}

// Output pixel idx over [B, HW, HW] of the head GEMM's rows x [B, g*g, p*p*3] (bf16 bit patterns; channel-last inside a
// patch, vision_transformer.py:515-527): clamp as modeling_mingtok.py:194, then trunc((v*std + mean) * 255).
MBP_HD void unpatchify_u8_pixel(const uint16_t* x, int g, int p, int64_t idx, const float mean[3], const float stdv[3],
                                uint8_t* out) {
  const int HW = g * p;
  const int xw = static_cast<int>(idx % HW);
  const int yh = static_cast<int>((idx / HW) % HW);
  const int64_t b = idx / (static_cast<int64_t>(HW) * HW);
  const int h = yh / p, pp = yh % p, w = xw / p, q = xw % p;
  const int64_t src = ((b * g * g + static_cast<int64_t>(h) * g + w) * (p * p) + pp * p + q) * 3;
  for (int c = 0; c < 3; ++c) {
    float v = bf16_bits_to_float(x[src + c]);
    v = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);  // NaN passes through, as torch.clamp does; the cast below gives 0
    out[idx * 3 + c] = denormalize_to_u8(v, mean[c], stdv[c]);
  }
}

}  // namespace mbpre
