// Image-side kernels (HBM-bound byte/float work):
//   * view_preprocess_kernel: u8 HWC source -> [flip] -> [cutout] -> /255 -> normalise -> bilinear resize
//     (ATen upsample_bilinear2d, align_corners=False) -> zero-pad to /32, fp32 NHWC3.  One fused pass.
//     Reference: torchvision F.to_tensor (cald_train.py:107), cald/cald_helper.py:23-30 (flip), 88-132
//     (cutout fill), tv:models/detection/transform.py:119-255.
//   * Pillow-exact u8 resampling (libImaging/Resample.c two-pass fixed point) and nearest affine rotate
//     (libImaging/Geometry.c affine_fixed) for cald_helper.resize / rotate (cald_helper.py:47-53,135-223).
#pragma once
#include <cmath>
#include <vector>
#include "common.cuh"

namespace cald {

constexpr int MAX_CUT = 4;

struct ViewDesc {
  const uint8_t* src;   // u8 [sh][sw][3]
  int sh, sw;           // source size
  int rh, rw;           // size after the detector's resize
  int flip;             // horizontal flip of the source
  int n_cut_slot;       // index into the device cutout-rect table, or -1
  // noise views (cald_helper.py:72-85): planes [3][sh][sw] drawn by the caller from torch's CPU generator
  const float* noise;   // null for the other views
  int noise_mode;       // 1: x + noise * std / 255     2: salt (noise < lo) / pepper (noise > hi)
  float n0, n1;         // mode 1: std            mode 2: lo, hi thresholds
  const int* mm;        // mode 2: device [2] = min, max of the u8 source; salt = max / 255, pepper = min / 255
  int perm;             // ColorSwap (cald_helper.py:56-62): source channel of output channel c = (perm >> 2c) & 3
};
constexpr int PERM_IDENTITY = 0 | (1 << 2) | (2 << 4);

struct CutRects {       // per image
  int n;
  int rect[MAX_CUT][4]; // l, t, r, b (half-open), in source pixels
};

__device__ __forceinline__ float src_pixel(const ViewDesc& d, const CutRects* cut, int c, int y, int x,
                                           float mean, float stdv) {
  // pixel of the augmented [0,1] image, then (x - mean) / std
  float v;
  bool zero = false;
  if (cut) {
    for (int k = 0; k < cut->n; ++k)
      zero |= (x >= cut->rect[k][0] && x < cut->rect[k][2] && y >= cut->rect[k][1] && y < cut->rect[k][3]);
  }
  int sx = d.flip ? (d.sw - 1 - x) : x;
  const int sc = (d.perm >> (2 * c)) & 3;
  v = zero ? 0.f : ((float)d.src[((long long)y * d.sw + sx) * 3 + sc] / 255.f);
  if (d.noise) {
    const float nz = d.noise[((long long)c * d.sh + y) * d.sw + x];
    if (d.noise_mode == 1) {
      v = v + (nz * d.n0) / 255.0f;
    } else {
      if (nz < d.n0) v = (float)d.mm[1] / 255.0f;   // salt = max(image)   (cald_helper.py:81-84)
      if (nz > d.n1) v = (float)d.mm[0] / 255.0f;   // pepper = min(image)
    }
  }
  return (v - mean) / stdv;
}

// grid: (ceil(Wp/64), Hp, V), block 64*3? -> one thread per (x, c) pair, 192 threads.
__global__ void view_preprocess_kernel(const ViewDesc* __restrict__ views, const CutRects* __restrict__ cuts,
                                       int Hp, int Wp, float* __restrict__ out /*[V][Hp][Wp][3]*/) {
  const int v = blockIdx.z, oy = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int ox = t / 3, c = t % 3;
  if (ox >= Wp) return;
  const ViewDesc d = views[v];
  const CutRects* cut = d.n_cut_slot >= 0 ? &cuts[d.n_cut_slot] : nullptr;
  const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
  const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
  float val = 0.f;
  if (oy < d.rh && ox < d.rw) {
    // ATen area_pixel_compute_source_index, scale = in / out in float (recompute_scale_factor=True)
    const float sh = (float)d.sh / (float)d.rh, sw = (float)d.sw / (float)d.rw;
    float fy = sh * ((float)oy + 0.5f) - 0.5f;
    if (fy < 0.f) fy = 0.f;
    float fx = sw * ((float)ox + 0.5f) - 0.5f;
    if (fx < 0.f) fx = 0.f;
    int y0 = (int)fy, x0 = (int)fx;
    int y1 = y0 + ((y0 < d.sh - 1) ? 1 : 0), x1 = x0 + ((x0 < d.sw - 1) ? 1 : 0);
    float ly = fy - (float)y0, lx = fx - (float)x0;
    float hy = 1.f - ly, hx = 1.f - lx;
    float p00 = src_pixel(d, cut, c, y0, x0, mean, stdv), p01 = src_pixel(d, cut, c, y0, x1, mean, stdv);
    float p10 = src_pixel(d, cut, c, y1, x0, mean, stdv), p11 = src_pixel(d, cut, c, y1, x1, mean, stdv);
    val = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
  }
  out[(((long long)v * Hp + oy) * Wp + ox) * 3 + c] = val;
}

// Fused transform + space-to-depth for the stem: writes the padded phase image
//   S[v][pr][pc][(py*2+px)*3 + c] = input(c, 2*(pr-2)+py, 2*(pc-2)+px)   (channels 12..15 = 0), split pl16,
// with pr in [0, Ho+3), pc in [0, Wo+3).  One thread per (pr, pc): 2 x 32-byte stores.
// The per-pixel work (resize coordinates, cutout test, flip index) is done once for the three channels, and the two
// fp32 divisions of F.to_tensor + normalize, (p / 255 - mean) / std, come from a per-block table built with exactly
// those operations (bit-identical results; noise views, whose values are not u8, take the arithmetic path).
__global__ void view_stem_input_kernel(const ViewDesc* __restrict__ views, const CutRects* __restrict__ cuts, int Hs,
                                       int Ws, pl16* __restrict__ ohi, pl16* __restrict__ olo) {
  __shared__ float s_unit[256];      // p / 255
  __shared__ float s_norm[3][256];   // (p / 255 - mean_c) / std_c
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const float u = (float)i / 255.f;
    s_unit[i] = u;
    s_norm[0][i] = (u - 0.485f) / 0.229f;
    s_norm[1][i] = (u - 0.456f) / 0.224f;
    s_norm[2][i] = (u - 0.406f) / 0.225f;
  }
  __syncthreads();
  const int v = blockIdx.z, pr = blockIdx.y;
  const int pc = blockIdx.x * blockDim.x + threadIdx.x;
  if (pc >= Ws) return;
  const ViewDesc d = views[v];
  const CutRects* cut = d.n_cut_slot >= 0 ? &cuts[d.n_cut_slot] : nullptr;
  const float sh = (float)d.sh / (float)d.rh, sw = (float)d.sw / (float)d.rw;
  // normalised value of the three channels of source pixel (y, x)
  auto fetch3 = [&](int y, int x, float* out) {
    bool zero = false;
    if (cut) {
      for (int k = 0; k < cut->n; ++k)
        zero |= (x >= cut->rect[k][0] && x < cut->rect[k][2] && y >= cut->rect[k][1] && y < cut->rect[k][3]);
    }
    const int sx = d.flip ? (d.sw - 1 - x) : x;
    const uint8_t* px = d.src + ((long long)y * d.sw + sx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int p = zero ? 0 : (int)px[(d.perm >> (2 * c)) & 3];
      if (!d.noise) {
        out[c] = s_norm[c][p];
      } else {
        const float mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
        const float stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
        float val = s_unit[p];
        const float nz = d.noise[((long long)c * d.sh + y) * d.sw + x];
        if (d.noise_mode == 1) {
          val = val + (nz * d.n0) / 255.0f;
        } else {
          if (nz < d.n0) val = (float)d.mm[1] / 255.0f;
          if (nz > d.n1) val = (float)d.mm[0] / 255.0f;
        }
        out[c] = (val - mean) / stdv;
      }
    }
  };
  float val[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oy = 2 * (pr - 2) + (q >> 1), ox = 2 * (pc - 2) + (q & 1);
    float* o = &val[q * 3];
    if (oy < 0 || ox < 0 || oy >= d.rh || ox >= d.rw) {  // zero padding of the batched image
      o[0] = o[1] = o[2] = 0.f;
      continue;
    }
    // ATen area_pixel_compute_source_index, scale = in / out in float (recompute_scale_factor=True)
    float fy = sh * ((float)oy + 0.5f) - 0.5f;
    if (fy < 0.f) fy = 0.f;
    float fx = sw * ((float)ox + 0.5f) - 0.5f;
    if (fx < 0.f) fx = 0.f;
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    if (ly == 0.f && lx == 0.f) {  // identity resize: exact copy
      fetch3(y0, x0, o);
      continue;
    }
    const int y1 = y0 + ((y0 < d.sh - 1) ? 1 : 0), x1 = x0 + ((x0 < d.sw - 1) ? 1 : 0);
    const float hy = 1.f - ly, hx = 1.f - lx;
    float p00[3], p01[3], p10[3], p11[3];
    fetch3(y0, x0, p00); fetch3(y0, x1, p01); fetch3(y1, x0, p10); fetch3(y1, x1, p11);
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = hy * (hx * p00[c] + lx * p01[c]) + ly * (hx * p10[c] + lx * p11[c]);
  }
  val[12] = val[13] = val[14] = val[15] = 0.f;
  uint32_t ph[8], pl[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_pack2(val[2 * i], val[2 * i + 1], ph[i], pl[i]);
  const long long off = (((long long)v * Hs + pr) * Ws + pc) * 16;
  uint4* dh = reinterpret_cast<uint4*>(ohi + off);
  dh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  dh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
  if (olo) {
    uint4* dl = reinterpret_cast<uint4*>(olo + off);
    dl[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    dl[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
  }
}

// ---------------------------------------------------------------- Pillow-exact colour enhancement
// cald_helper.ColorAdjust (cald_helper.py:65-69) = torchvision F.adjust_brightness -> adjust_contrast ->
// adjust_saturation on the PIL image = PIL.ImageEnhance {Brightness, Contrast, Color}: Image.blend(degenerate, image,
// factor), libImaging/Blend.c:  temp = (float)((int)in1 + alpha * ((int)in2 - (int)in1)); alpha in [0,1] -> (UINT8)temp,
// otherwise clip to [0,255] first.  Luma = libImaging/Convert.c rgb2l: (R*19595 + G*38470 + B*7471 + 0x8000) >> 16.
__device__ __forceinline__ int pil_blend(int in1, int in2, float alpha, bool interp) {
  const float temp = (float)in1 + alpha * (float)(in2 - in1);
  if (interp) return (int)(uint8_t)(int)temp;
  if (temp <= 0.f) return 0;
  if (temp >= 255.f) return 255;
  return (int)temp;
}
__device__ __forceinline__ int pil_luma(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// pass 1: brightness (degenerate = black) -> out; sum of the brightened image's luma -> luma_sum (for Contrast's mean)
__global__ void pil_brightness_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long npix,
                                      float alpha, unsigned long long* __restrict__ luma_sum) {
  const bool interp = alpha >= 0.f && alpha <= 1.f;
  unsigned long long local = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const int r = pil_blend(0, in[i * 3], alpha, interp), g = pil_blend(0, in[i * 3 + 1], alpha, interp),
              b = pil_blend(0, in[i * 3 + 2], alpha, interp);
    out[i * 3] = (uint8_t)r; out[i * 3 + 1] = (uint8_t)g; out[i * 3 + 2] = (uint8_t)b;
    local += (unsigned long long)pil_luma(r, g, b);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(luma_sum, local);
}
// pass 2 (in place): contrast against the rounded mean luma, then saturation against the per-pixel luma
__global__ void pil_contrast_saturation_kernel(uint8_t* __restrict__ img, long long npix, float alpha,
                                               const unsigned long long* __restrict__ luma_sum) {
  const bool interp = alpha >= 0.f && alpha <= 1.f;
  // ImageEnhance.Contrast: int(ImageStat.Stat(L).mean[0] + 0.5), python floats (double)
  const int mean = (int)((double)luma_sum[0] / (double)npix + 0.5);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const int r = pil_blend(mean, img[i * 3], alpha, interp), g = pil_blend(mean, img[i * 3 + 1], alpha, interp),
              b = pil_blend(mean, img[i * 3 + 2], alpha, interp);
    const int l = pil_luma(r, g, b);
    img[i * 3] = (uint8_t)pil_blend(l, r, alpha, interp);
    img[i * 3 + 1] = (uint8_t)pil_blend(l, g, alpha, interp);
    img[i * 3 + 2] = (uint8_t)pil_blend(l, b, alpha, interp);
  }
}

// ---------------------------------------------------------------- Pillow-exact resampling
constexpr int PIL_PRECISION_BITS = 32 - 8 - 2;

struct PilCoeffs {          // host-built (double arithmetic as in Resample.c), uploaded once per (in,out,filter)
  int in_size, out_size, ksize;
  std::vector<int> bounds;  // [out][2] xmin, xcnt
  std::vector<int> kk;      // [out][ksize]
};

inline double pil_bilinear(double x) { if (x < 0.0) x = -x; return x < 1.0 ? 1.0 - x : 0.0; }
inline double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
// filter: 0 bilinear (support 1), 1 bicubic (support 2)
inline PilCoeffs pil_precompute(int in_size, int out_size, int filter) {
  PilCoeffs pc;
  pc.in_size = in_size; pc.out_size = out_size;
  double support = filter == 0 ? 1.0 : 2.0;
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  support = support * filterscale;
  int ksize = (int)ceil(support) * 2 + 1;
  pc.ksize = ksize;
  pc.bounds.assign((size_t)out_size * 2, 0);
  pc.kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double w = (filter == 0 ? pil_bilinear : pil_bicubic)((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < xmax; ++x) {
      double s = k[x] * (1 << PIL_PRECISION_BITS);
      pc.kk[(size_t)xx * ksize + x] = (int)(k[x] < 0 ? -0.5 + s : 0.5 + s);
    }
    pc.bounds[xx * 2] = xmin;
    pc.bounds[xx * 2 + 1] = xmax;
  }
  return pc;
}

__device__ __forceinline__ uint8_t pil_clip8(int v) {
  v >>= PIL_PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// horizontal pass: in [h][iw][3] -> out [h][ow][3]
__global__ void pil_resample_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h, int iw, int ow,
                                      const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (t >= ow * 3) return;
  int xx = t / 3, c = t % 3;
  int xmin = bounds[xx * 2], n = bounds[xx * 2 + 1];
  int acc = 1 << (PIL_PRECISION_BITS - 1);
  const uint8_t* row = in + (long long)y * iw * 3;
  for (int x = 0; x < n; ++x) acc += (int)row[(xmin + x) * 3 + c] * kk[xx * ksize + x];
  out[((long long)y * ow + xx) * 3 + c] = pil_clip8(acc);
}
// vertical pass: in [ih][w][3] -> out [oh][w][3]
__global__ void pil_resample_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int ih, int oh, int w,
                                      const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int yy = blockIdx.y;
  if (t >= w * 3) return;
  int ymin = bounds[yy * 2], n = bounds[yy * 2 + 1];
  int acc = 1 << (PIL_PRECISION_BITS - 1);
  for (int y = 0; y < n; ++y) acc += (int)in[(long long)(ymin + y) * w * 3 + t] * kk[yy * ksize + y];
  out[(long long)yy * w * 3 + t] = pil_clip8(acc);
}

// Batched forms: one launch resamples / rotates a whole chunk's images (blockIdx.z = job); every job carries its own
// sizes and coefficient tables, threads outside a job's extent exit.
struct PilJob {
  const uint8_t* src;
  uint8_t* dst;
  int in_h, in_w, out_h, out_w;   // the pass changes ONE of the two dimensions
  const int* bounds;
  const int* kk;
  int ksize;
};
// one thread per output PIXEL: the three channels share the coefficient loads and the loop (same integer arithmetic)
__global__ void pil_resample_h_batched_kernel(const PilJob* __restrict__ jobs) {
  const PilJob j = jobs[blockIdx.z];
  const int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (y >= j.in_h || xx >= j.out_w) return;
  const int xmin = j.bounds[xx * 2], n = j.bounds[xx * 2 + 1];
  int a0 = 1 << (PIL_PRECISION_BITS - 1), a1 = a0, a2 = a0;
  const uint8_t* px = j.src + ((long long)y * j.in_w + xmin) * 3;
  const int* k = j.kk + (long long)xx * j.ksize;
  for (int x = 0; x < n; ++x) {
    const int c = k[x];
    a0 += (int)px[0] * c; a1 += (int)px[1] * c; a2 += (int)px[2] * c;
    px += 3;
  }
  uint8_t* o = j.dst + ((long long)y * j.out_w + xx) * 3;
  o[0] = pil_clip8(a0); o[1] = pil_clip8(a1); o[2] = pil_clip8(a2);
}
// one thread per four consecutive bytes of an output row (the coefficient of a tap is the same for the whole row)
__global__ void pil_resample_v_batched_kernel(const PilJob* __restrict__ jobs) {
  const PilJob j = jobs[blockIdx.z];
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4, yy = blockIdx.y;
  const int rowb = j.in_w * 3;
  if (yy >= j.out_h || t >= rowb) return;
  const int ymin = j.bounds[yy * 2], n = j.bounds[yy * 2 + 1];
  const int m = min(4, rowb - t);
  int acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = 1 << (PIL_PRECISION_BITS - 1);
  const uint8_t* p = j.src + (long long)ymin * rowb + t;
  const int* k = j.kk + (long long)yy * j.ksize;
  for (int y = 0; y < n; ++y) {
    const int c = k[y];
#pragma unroll
    for (int q = 0; q < 4; ++q) if (q < m) acc[q] += (int)p[q] * c;
    p += rowb;
  }
  uint8_t* o = j.dst + (long long)yy * rowb + t;
#pragma unroll
  for (int q = 0; q < 4; ++q) if (q < m) o[q] = pil_clip8(acc[q]);
}

// PIL Image.rotate(angle, expand=True) geometry (PIL/Image.py) + Geometry.c affine_fixed coefficients
struct RotateGeom { int nw, nh; int a0, a1, a2, a3, a4, a5; };
inline double py_round15(double v) { return std::round(v * 1e15) / 1e15; }
inline RotateGeom pil_rotate_geom(int w, int h, double angle_deg) {
  // note: python's round(x, 15) is correctly-rounded decimal; for cos/sin of 5 degrees the value
  // has no representable neighbour within 1e-15 that changes the 16.16 fixed-point result.
  const double PI = 3.14159265358979323846;
  double ang = -(angle_deg * PI / 180.0);
  double m[6] = {py_round15(cos(ang)), py_round15(sin(ang)), 0.0, py_round15(-sin(ang)), py_round15(cos(ang)), 0.0};
  double cx = w / 2.0, cy = h / 2.0;
  auto tf = [&](double x, double y, double& ox, double& oy) {
    ox = m[0] * x + m[1] * y + m[2];
    oy = m[3] * x + m[4] * y + m[5];
  };
  double t2, t5;
  tf(-cx, -cy, t2, t5);
  m[2] = t2 + cx;
  m[5] = t5 + cy;
  double xs[4], ys[4];
  double px[4] = {0, (double)w, (double)w, 0}, py[4] = {0, 0, (double)h, (double)h};
  for (int i = 0; i < 4; ++i) tf(px[i], py[i], xs[i], ys[i]);
  double xmin = xs[0], xmax = xs[0], ymin = ys[0], ymax = ys[0];
  for (int i = 1; i < 4; ++i) {
    xmin = std::min(xmin, xs[i]); xmax = std::max(xmax, xs[i]);
    ymin = std::min(ymin, ys[i]); ymax = std::max(ymax, ys[i]);
  }
  RotateGeom g;
  g.nw = (int)(ceil(xmax) - floor(xmin));
  g.nh = (int)(ceil(ymax) - floor(ymin));
  tf(-(g.nw - w) / 2.0, -(g.nh - h) / 2.0, t2, t5);
  m[2] = t2;
  m[5] = t5;
  auto fix = [](double v) { return (int)floor(v * 65536.0 + 0.5); };
  g.a0 = fix(m[0]); g.a1 = fix(m[1]); g.a3 = fix(m[3]); g.a4 = fix(m[4]);
  g.a2 = fix(m[2] + m[0] * 0.5 + m[1] * 0.5);
  g.a5 = fix(m[5] + m[3] * 0.5 + m[4] * 0.5);
  return g;
}
__global__ void pil_rotate_nearest_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int w, int h,
                                          RotateGeom g) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= g.nw) return;
  long long xx = (long long)g.a2 + (long long)g.a1 * y + (long long)g.a0 * x;
  long long yy = (long long)g.a5 + (long long)g.a4 * y + (long long)g.a3 * x;
  int xin = (int)(xx >> 16), yin = (int)(yy >> 16);
  uint8_t r = 0, gg = 0, b = 0;
  if (xin >= 0 && xin < w && yin >= 0 && yin < h) {
    const uint8_t* p = in + ((long long)yin * w + xin) * 3;
    r = p[0]; gg = p[1]; b = p[2];
  }
  uint8_t* o = out + ((long long)y * g.nw + x) * 3;
  o[0] = r; o[1] = gg; o[2] = b;
}

struct RotJob {
  const uint8_t* src;
  uint8_t* dst;
  int w, h;
  RotateGeom g;
};
__global__ void pil_rotate_nearest_batched_kernel(const RotJob* __restrict__ jobs) {
  const RotJob j = jobs[blockIdx.z];
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= j.g.nw || y >= j.g.nh) return;
  const long long xx = (long long)j.g.a2 + (long long)j.g.a1 * y + (long long)j.g.a0 * x;
  const long long yy = (long long)j.g.a5 + (long long)j.g.a4 * y + (long long)j.g.a3 * x;
  const int xin = (int)(xx >> 16), yin = (int)(yy >> 16);
  uint8_t r = 0, gg = 0, b = 0;
  if (xin >= 0 && xin < j.w && yin >= 0 && yin < j.h) {
    const uint8_t* p = j.src + ((long long)yin * j.w + xin) * 3;
    r = p[0]; gg = p[1]; b = p[2];
  }
  uint8_t* o = j.dst + ((long long)y * j.g.nw + x) * 3;
  o[0] = r; o[1] = gg; o[2] = b;
}

}  // namespace cald
