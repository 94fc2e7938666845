// Transposed-role implicit-GEMM convolution for the 3x3 convs with 64 output channels (ResNet layer1), sm_100a.
//
// Why a third kernel.  With pixels on M and channels on N (igemm.cuh / igemm2.cuh) a 64-channel layer in the split
// arithmetic can only issue N = 128 ([B_hi | B_lo]) and N = 64 instructions, and a tcgen05.mma with both operands in
// shared memory has a fixed cost besides the part that scales with N (fitted from the measurements below: about
// 33 ns + 0.36 ns per column at boost clocks -- N = 64: 56 ns, N = 128: 79 ns, N = 256: 126 ns).  Here the roles are
// swapped:
//
//   D^T[128 x 256] = W_stack[128 x K] * Act^T[K x 256]
//
//   * M = 128 rows = the 64 output channels TWICE: the hi and the lo plane of the weights, interleaved in groups of 16
//     rows ([hi 0..15 | lo 0..15 | hi 16..31 | lo 16..31 | ...]) so that a channel's main and cross-term accumulators
//     sit in the same 32-lane TMEM quadrant, 16 lanes apart (one epilogue warp sees both);
//   * N = 256 = the pixels of a 16 x 16 output patch (one TMA box per plane and tap);
//   * per k-step two FULL-WIDTH instructions:  [W_hi ; W_lo] x A_hi^T   (main rows += W_hi*A_hi, cross rows += W_lo*A_hi)
//                                              [ 0   ; W_hi] x A_lo^T   (cross rows += W_hi*A_lo)
//     The main rows see one accumulate per k-step and the cross rows two, exactly like the XSEP columns of igemm.cuh,
//     so the truncation pre-compensation in the weights (RzPlan) is unchanged and the sums are the same products in
//     the same k order: results are BIT-IDENTICAL to the pixel-major kernels (tests/test_gpu_conv.py).
//
// Measured (profiles/r02_tform.md): the layer1 3x3 conv alone, 16 views, boost clocks: one CTA 0.2895 ms, CTA pair
// 0.2737 ms, this kernel 0.2686 ms; inside the power-capped cfg-2 step 708 -> 633 us per 32-view launch (-10.6 %).
// The quarter of the second instruction that multiplies zeros is what keeps the gain small.  The stem (4 k-blocks per
// patch) is 4 % slower on this kernel -- with so few MMAs per patch the per-slab epilogue below (two named barriers, a
// proxy fence and two TMA stores per 32 pixels) is exposed -- and stays on igemm.cuh.
//
// Epilogue: a thread owns one (channel, plane) row of the accumulator and walks the tile in slabs of 32 pixels (two
// patch rows): tcgen05.ld of 32 columns, one shuffle per value to bring main and cross together (the lower half-warp
// finishes the slab's first patch row, the upper half-warp the second), bias, ReLU, split to half planes, 2-byte stores
// into a 128B-swizzled [32 pixels][64 channels] staging slab, one TMA store per plane and slab (double-buffered).
#pragma once
#include "igemm.cuh"

namespace cald {

constexpr int IGT_TW = 16, IGT_TH = 16;                 // output patch = 256 pixels = N of the MMA
constexpr int IGT_ACT_BYTES = 256 * 128;                // one plane of one k-block of the patch: 32 KB
constexpr int IGT_W_BYTES = 128 * 128;                  // one stacked weight operand of one k-block: 16 KB
constexpr int IGT_STAGE_BYTES = 2 * IGT_ACT_BYTES + 2 * IGT_W_BYTES;   // 96 KB
constexpr int IGT_STAGES = 2;
constexpr int IGT_OFF_A_LO = IGT_ACT_BYTES;
constexpr int IGT_OFF_W1 = 2 * IGT_ACT_BYTES;
constexpr int IGT_OFF_W2 = 2 * IGT_ACT_BYTES + IGT_W_BYTES;
constexpr int IGT_SLAB_PIX = 32;                        // pixels per output slab (two patch rows)
constexpr int IGT_SLAB_BYTES = IGT_SLAB_PIX * 128;      // one plane of one slab: 4 KB
constexpr int IGT_OUT_BYTES = 2 /*slots*/ * 2 /*planes*/ * IGT_SLAB_BYTES;
constexpr int IGT_BAR_BYTES = 1024;
constexpr int IGT_SMEM_BYTES = IGT_STAGES * IGT_STAGE_BYTES + IGT_BAR_BYTES + IGT_OUT_BYTES + 1024;

// Stacked weight operands of a 64-output-channel layer: wt = [2][128][K]
//   operand 0, row r: group g = r / 16, channel c = (g / 2) * 16 + r % 16:  g even -> W_hi[c], g odd -> W_lo[c]
//   operand 1, row r:                                                      g even -> 0,       g odd -> W_hi[c]
// w = [2][64][K] (hi plane, lo plane), the layout every other kernel reads.
__global__ void igemm_t_stack_weights_kernel(const pl16* __restrict__ w, pl16* __restrict__ wt, int K) {
  const long long total = 2LL * 128 * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int r = (int)((i / K) % 128);
    const int op = (int)(i / ((long long)128 * K));
    const int g = r >> 4, c = (g >> 1) * 16 + (r & 15);
    const pl16 hi = w[(long long)c * K + k], lo = w[(long long)(64 + c) * K + k];
    pl16 v;
    if (op == 0) v = (g & 1) ? lo : hi;
    else v = (g & 1) ? hi : float_to_pl16(0.f);
    wt[i] = v;
  }
}

__device__ __forceinline__ void igt_tile_coords(const ConvParams& p, int tile, int& img, int& y0, int& x0) {
  const int tx = tile % p.tiles_x;
  const int m = tile / p.tiles_x;
  const int ty = m % p.tiles_y;
  img = m / p.tiles_y;
  y0 = ty * IGT_TH;
  x0 = tx * IGT_TW;
}

__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_t_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + IGT_STAGES * IGT_STAGE_BYTES;
  // barrier layout: full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  const uint32_t full_bar = bars, empty_bar = bars + 8 * IGT_STAGES;
  const uint32_t tfull_bar = empty_bar + 8 * IGT_STAGES, tempty_bar = tfull_bar + 16;
  volatile uint32_t* tmem_holder =
      reinterpret_cast<volatile uint32_t*>(smem_al + IGT_STAGES * IGT_STAGE_BYTES + 16 * IGT_STAGES + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < IGT_STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_holder)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int k_chunks = p.Cin / IG_BLOCK_K;
  const int num_kb = p.taps * k_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int img, y0, x0;
        igt_tile_coords(p, tile, img, y0, x0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = smem_base + stage * IGT_STAGE_BYTES;
          const int tap = kb / k_chunks;
          const int c0 = (kb - tap * k_chunks) * IG_BLOCK_K;
          const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap];
          const int bk = tap * p.Cin + c0;
          mbar_expect_tx(fb, IGT_STAGE_BYTES);
          tma_load_4d(sa, &tmA, fb, c0, ax, ay, img);
          tma_load_4d(sa + IGT_OFF_A_LO, &tmA, fb, c0, ax, ay, img + p.a_lo_img);
          tma_load_4d(sa + IGT_OFF_W1, &tmW, fb, bk, 0, 0, 0);
          tma_load_4d(sa + IGT_OFF_W2, &tmW, fb, bk, 0, 0, 1);
          if (++stage == IGT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // M = 128 (stacked weight rows), N = 256 (pixels), both operands K-major halves, fp32 accumulate
      const uint32_t idesc = (1u << 4) | (PL16_MMA_FMT << 7) | (PL16_MMA_FMT << 10) | ((uint32_t)(256 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * IGT_STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(sa);
          const uint64_t a_lo = umma_desc_sw128(sa + IGT_OFF_A_LO);
          const uint64_t w1 = umma_desc_sw128(sa + IGT_OFF_W1);
          const uint64_t w2 = umma_desc_sw128(sa + IGT_OFF_W2);
#pragma unroll
          for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
            const uint64_t ko = (uint64_t)((k * IG_UMMA_K * 2) >> 4);
            tcgen05_mma_bf16(d, w1 + ko, a_hi + ko, idesc, (kb | k) != 0);   // main += W_hi*A_hi, cross += W_lo*A_hi
            tcgen05_mma_bf16(d, w2 + ko, a_lo + ko, idesc, 1);               // cross += W_hi*A_lo (main += 0)
          }
          tcgen05_commit(empty_bar + 8 * stage);
          if (++stage == IGT_STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(tfull_bar + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;            // TMEM lane quadrant of this warp: stacked rows [32 * quad, 32 * quad + 32)
    const bool upper = lane >= 16;        // lower half-warp holds the main rows, upper the cross-term rows
    const int ch = quad * 16 + (lane & 15);
    const bool leader = (threadIdx.x == 64);
    const uint32_t out_base = smem_base + IGT_STAGES * IGT_STAGE_BYTES + IGT_BAR_BYTES;  // 1024-aligned
    const float bias = p.bias ? p.bias[ch] : 0.f;
    // this thread's 2-byte position inside a 128-byte staging row (before the per-row swizzle of the 16-byte chunk)
    const uint32_t ch_chunk = (uint32_t)(ch >> 3), ch_in = (uint32_t)(ch & 7) * 2u;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab = 0;   // running slab counter: slot = slab & 1
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int img, y0, x0;
      igt_tile_coords(p, tile, img, y0, x0);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tcgen05_fence_after();
      const uint32_t t0 = tmem_base + acc * 256 + ((uint32_t)(quad * 32) << 16);
      // patch rows below the image would be clipped by the TMA store anyway: their slabs are not computed
      const int rows_valid = p.H - y0 < IGT_TH ? p.H - y0 : IGT_TH;
      const int n_slabs = (rows_valid + 1) / 2;
      for (int s = 0; s < n_slabs; ++s, ++slab) {
        uint32_t r[32];
        tmem_ld32(t0 + s * IGT_SLAB_PIX, r);
        if (s == n_slabs - 1) {   // accumulator drained: hand the TMEM stage back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
        }
        const uint32_t sb = out_base + (slab & 1u) * (2u * IGT_SLAB_BYTES);
        // the TMA store that last read this slot (two slabs ago) must have finished reading it
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // lower half-warp finishes pixel j of the slab (first patch row), upper half-warp pixel 16 + j
          const uint32_t send = upper ? r[j] : r[16 + j];
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 16);
          const float mainv = upper ? __uint_as_float(recv) : __uint_as_float(r[j]);
          const float cross = upper ? __uint_as_float(r[16 + j]) : __uint_as_float(recv);
          float v = fmaf(cross, CALD_LO_INV, mainv) + bias;
          if (p.relu) v = fmaxf(v, 0.f);
          // split to the two half planes (saturating conversions, lo = (v - hi) * 2^11 as an exponent add)
          const uint32_t hw = cvt_pack2(v, 0.f);
          float hf, unused;
          unpack2(hw, hf, unused);
          const uint32_t lw = cvt_pack2(scale_2p11(v - hf), 0.f);
          const uint32_t prow = (upper ? 16u : 0u) + (uint32_t)j;
          const uint32_t off = prow * 128u + ((ch_chunk ^ (prow & 7u)) << 4) + ch_in;
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(sb + off), "h"((unsigned short)(hw & 0xffffu)) : "memory");
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(sb + IGT_SLAB_BYTES + off), "h"((unsigned short)(lw & 0xffffu))
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          const int ys = y0 + 2 * s;
          tma_store_4d(&tmC, sb, 0, x0, ys, img);
          tma_store_4d(&tmC, sb + IGT_SLAB_BYTES, 0, x0, ys, img + p.c_lo_img);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace cald
