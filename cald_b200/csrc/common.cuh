// Shared device/host helpers for the CALD B200 scoring engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

#define CALD_CUDA_CHECK(expr)                                                            \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      char _b[512];                                                                      \
      snprintf(_b, sizeof(_b), "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,      \
               cudaGetErrorString(_e));                                                  \
      throw std::runtime_error(_b);                                                      \
    }                                                                                    \
  } while (0)

namespace cald {

// ---------------------------------------------------------------------------------
// Split-half activation / weight format.  A real value v is stored as two IEEE half planes
//   hi = rn16(v),  lo = rn16((v - hi) * 2^11)            =>  v ~= hi + lo * 2^-11
// 11 + 11 significand bits: the operands carry fp32's own precision (measured: fp32 FMAs on the split operands land
// on the same error as torch's CPU fp32 conv, tools/conv_accuracy.py) at the tcgen05 kind::f16 rate and 4 B / element.
// The tensor-core path multiplies A_hi*B_hi + (A_hi*B_lo + A_lo*B_hi) * 2^-11 with fp32 accumulation in TMEM (the
// cross terms have their own accumulator columns, igemm.cuh XSEP), which is what keeps the detector's discrete
// stages (top-k / NMS / arg-max) on the same side of their thresholds as the reference's fp32 CPU run.
// The lo plane is pre-scaled so that it stays in half's normal range; conversions saturate at +-65504 instead of
// producing inf.  (Round 1 used bfloat16 planes: 8 + 8 bits, 64x coarser.)
// ---------------------------------------------------------------------------------
typedef __half pl16;
#define CALD_LO_SCALE 2048.0f
#define CALD_LO_INV (1.0f / 2048.0f)
// tcgen05 instruction-descriptor operand format field (a_format bits [7,10), b_format bits [10,13)): 0 = F16, 1 = BF16
constexpr uint32_t PL16_MMA_FMT = 0u;

__host__ __device__ inline float pl16_to_float(pl16 v) { return __half2float(v); }
__host__ __device__ inline pl16 float_to_pl16(float v) {
  // saturate instead of producing inf (an inf operand would turn a whole accumulator row into NaN)
  v = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
  return __float2half_rn(v);
}

__host__ __device__ inline void split_pl(float v, pl16& hi, pl16& lo) {
  hi = float_to_pl16(v);
  lo = float_to_pl16((v - pl16_to_float(hi)) * CALD_LO_SCALE);
}

__device__ __forceinline__ float join_pl(pl16 hi, pl16 lo) {
  return fmaf(pl16_to_float(lo), CALD_LO_INV, pl16_to_float(hi));  // the scale is a power of two: exact product
}

// two floats -> packed half2 word (first argument in the low half), saturating
__device__ __forceinline__ uint32_t cvt_pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  a = f.x; b = f.y;
}
// x * 2^11 by adding 11 to the exponent field: an integer add on the ALU pipe instead of one more FP32 multiply in the
// (FP32-pipe bound) epilogue.  Exact for normal x; +-0 becomes +-2^-116, which the half conversion rounds back to +-0.
__device__ __forceinline__ float scale_2p11(float x) { return __int_as_float(__float_as_int(x) + (11 << 23)); }

// Two floats -> packed (hi plane, lo plane) words
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = cvt_pack2(a, b);
  float ha, hb;
  unpack2(hi, ha, hb);
  lo = cvt_pack2(scale_2p11(a - ha), scale_2p11(b - hb));
}
// packed (hi, lo plane) words -> two floats
__device__ __forceinline__ void join_pack2(uint32_t hi, uint32_t lo, float& a, float& b) {
  float ha, hb, la, lb;
  unpack2(hi, ha, hb);
  unpack2(lo, la, lb);
  a = fmaf(la, CALD_LO_INV, ha);
  b = fmaf(lb, CALD_LO_INV, hb);
}

__device__ __forceinline__ uint32_t pack_pl16x2(pl16 a, pl16 b) {
  return (uint32_t)(*reinterpret_cast<unsigned short*>(&a)) | ((uint32_t)(*reinterpret_cast<unsigned short*>(&b)) << 16);
}

// ---------------------------------------------------------------------------------
// Conv / GEMM problem description shared by the tcgen05 kernel and the SIMT checker.
// ---------------------------------------------------------------------------------
enum { RES_NONE = 0, RES_SAME = 1, RES_NEAREST = 2 };

struct ConvParams {
  // ---- A operand (activations, split planes [planes][n_img][H_in][W_in][Cin])
  int n_img;           // images (views) in the batch; 1 in linear mode
  int H, W;            // OUTPUT spatial size (== input size for stride 1); linear: H = 1, W = M
  int Cin;             // multiple of 64
  int taps;            // 1 (1x1 / linear) or 9 (3x3)
  int tap_dy[9], tap_dx[9], tap_img[9];  // per tap: input pixel offset (and an image-index offset, unused = 0)
  int a_lo_img;        // image-index offset of the lo plane in the A tensor map
  // ---- tiling
  int tw, th;          // spatial tile, tw * th <= 128 (rows beyond tw*th of the MMA tile are dead)
  int a_bytes;         // bytes one A-tile TMA box delivers per plane = th * tw * 128
  int tiles_x, tiles_y;
  int n_blocks;        // ceil(Cout / BLOCK_N)
  int num_tiles;
  // ---- B operand (weights [2][Cout][taps*Cin])
  int Cout;
  // ---- epilogue
  const float* bias;   // [Cout] or null
  int relu;
  int res_mode;        // RES_*
  const pl16* res_hi;
  const pl16* res_lo;
  int res_H, res_W;    // for RES_NEAREST: source size
  pl16* out_hi;        // null when only fp32 output is wanted
  pl16* out_lo;        // null in pl16 (non-split) mode
  float* out_f32;      // optional fp32 NHWC copy (heads)
  int ldc;             // channel stride of the output row (>= Cout)
  int res_ld;
  int kc;              // chunked accumulation: k-blocks per TMEM accumulation chunk
  int tma_store;       // 1: epilogue stages 64-channel groups in smem and stores them with TMA (NHWC pl16 outputs)
  int c_lo_img;        // image-index offset of the lo plane in the output tensor map
  int res_kb;          // residual-as-MMA: extra identity k-blocks per tile (BLOCK_N / 64), else 0
  int r_lo_img;        // image-index offset of the lo plane in the residual tensor map
  int a_scale;         // 1, or 2 for a stride-2 1x1 conv: the A tensor map walks the input with element stride 2
  int res_conv;        // 1: the extra k-blocks are a second 1x1 contraction (aux input x aux weights, e.g. the ResNet
                       //    downsample shortcut) accumulated into the same tile, not an identity-routed residual
  int r_scale;         // element stride of the aux input's tensor map (2 for a stride-2 shortcut), else 1
  int res_narrow;      // 1: identity-routed residual k-blocks issue N = 64 instructions on their own 64 columns
};

// ATen nearest-neighbour source index (UpSampleKernel: nearest_idx), float scale.
__host__ __device__ inline int nearest_src(int dst, int in_size, int out_size) {
  if (out_size == in_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  float scale = (float)in_size / (float)out_size;
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

// Row (pixel) -> output element offset (without channel).  Returns -1 if masked.
__device__ __forceinline__ long long out_row_offset(const ConvParams& p, int n, int y, int x) {
  return (((long long)n * p.H + y) * p.W + x) * (long long)p.ldc;
}

__device__ __forceinline__ long long res_row_offset(const ConvParams& p, int n, int y, int x) {
  if (p.res_mode == RES_NEAREST) {
    int sy = nearest_src(y, p.res_H, p.H);
    int sx = nearest_src(x, p.res_W, p.W);
    return (((long long)n * p.res_H + sy) * p.res_W + sx) * (long long)p.res_ld;
  }
  return (((long long)n * p.H + y) * p.W + x) * (long long)p.res_ld;
}

// Apply bias / residual / ReLU to 8 consecutive channels and store them.
__device__ __forceinline__ void epilogue_store8(const ConvParams& p, long long orow, long long rrow,
                                                int c, float* v) {
  if (p.bias) {
    const float4 b0 = *reinterpret_cast<const float4*>(p.bias + c);
    const float4 b1 = *reinterpret_cast<const float4*>(p.bias + c + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (p.res_mode != RES_NONE) {
    uint4 rh = *reinterpret_cast<const uint4*>(p.res_hi + rrow + c);
    const pl16* h = reinterpret_cast<const pl16*>(&rh);
    if (p.res_lo) {
      uint4 rl = *reinterpret_cast<const uint4*>(p.res_lo + rrow + c);
      const pl16* l = reinterpret_cast<const pl16*>(&rl);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += join_pl(h[i], l[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += pl16_to_float(h[i]);
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.out_f32) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + orow + c);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (p.out_hi) {
    pl16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_pl(v[i], h[i], l[i]);
    uint4 ph = make_uint4(pack_pl16x2(h[0], h[1]), pack_pl16x2(h[2], h[3]), pack_pl16x2(h[4], h[5]),
                          pack_pl16x2(h[6], h[7]));
    *reinterpret_cast<uint4*>(p.out_hi + orow + c) = ph;
    if (p.out_lo) {
      uint4 pl = make_uint4(pack_pl16x2(l[0], l[1]), pack_pl16x2(l[2], l[3]), pack_pl16x2(l[4], l[5]),
                            pack_pl16x2(l[6], l[7]));
      *reinterpret_cast<uint4*>(p.out_lo + orow + c) = pl;
    }
  }
}

}  // namespace cald
