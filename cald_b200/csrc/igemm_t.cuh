// Transposed-role implicit-GEMM convolution for the spatial convs with 64 output channels (stem 7x7/2, ResNet layer1
// 3x3), sm_100a.
//
// Why a third kernel.  With pixels on M and channels on N (igemm.cuh / igemm2.cuh) a 64-channel layer in the split
// arithmetic can only issue N = 128 ([B_hi | B_lo]) and N = 64 instructions, and a tcgen05.mma with both operands in
// shared memory has a fixed cost besides the part that scales with N (fitted from the measurements in
// profiles/r02_tform.md: about 33 ns + 0.36 ns per column at boost clocks -- N = 64: 56 ns, N = 128: 79 ns,
// N = 256: 126 ns).  Here the roles are swapped:
//
//   D^T[128 x 256] = W_stack[128 x K] * Act^T[K x 256]
//
//   * M = 128 rows = the 64 output channels TWICE: the hi and the lo plane of the weights, stacked in halves of 16 rows
//     per TMEM quadrant ([hi x16 | lo x16] x 4) so that a channel's main and cross-term accumulators sit in the same
//     32-lane quadrant, 16 lanes apart (one epilogue warp sees both);
//   * N = 256 = the pixels of a 16 x 16 output patch (one TMA box per plane and tap);
//   * per k-step two FULL-WIDTH instructions:  [W_hi ; W_lo] x A_hi^T   (main rows += W_hi*A_hi, cross rows += W_lo*A_hi)
//                                              [ 0   ; W_hi] x A_lo^T   (cross rows += W_hi*A_lo)
//     The main rows see one accumulate per k-step and the cross rows two, exactly like the XSEP columns of igemm.cuh,
//     so the truncation pre-compensation in the weights (RzPlan) is unchanged.  The main sums are bit-identical to the
//     pixel-major kernels; the cross-term sums add the same products with the two partial sums of a k-block in the
//     other order (all four W_lo*A_hi k-steps, then all four W_hi*A_lo ones: the halves of a k-block are separate ring
//     slots), which moves results by <= 1e-7 of the output scale (tests/test_gpu_conv.py).
//
// Ring of four HALF stages (48 KB: one activation plane + one stacked weight operand of a k-block): three loads are in
// flight while one slot is being multiplied.  With two whole stages ncu showed 70 % tensor-pipe activity at 64 % L2
// throughput -- the load latency was exposed.
//
// Measured inside the power-capped cfg-2 step (profiles/r02_tform.md), per 32-view launch: layer1 3x3 708 us (CTA-pair
// kernel) -> 573 us (-19 %), stem 1469 us (one-CTA kernel) -> 1315 us (-10.5 %).  The quarter of the second instruction
// that multiplies zeros is what keeps the gain from being larger.
//
// Epilogue: 16x256b TMEM loads hand a thread the main and the cross-term value of the same (channel pair, pixel pair)
// elements (the weight rows are stacked so that the two rows of such a fragment are adjacent channels): one FMA joins
// them, bias, ReLU, split to half planes, one 4-byte bank-conflict-free store per pixel and plane into a
// 128B-swizzled [64 pixels][64 channels] staging slab, one TMA store per plane and slab (double-buffered).
#pragma once
#include "igemm.cuh"

namespace cald {

constexpr int IGT_TW = 16, IGT_TH = 16;                 // output patch = 256 pixels = N of the MMA
constexpr int IGT_ACT_BYTES = 256 * 128;                // one plane of one k-block of the patch: 32 KB
constexpr int IGT_W_BYTES = 128 * 128;                  // one stacked weight operand of one k-block: 16 KB
// The ring is made of HALF stages: (A_hi, [W_hi ; W_lo]) and (A_lo, [0 ; W_hi]) of a k-block travel and are consumed
// separately -- four 48 KB slots keep three loads in flight while one is being multiplied (two 96 KB stages kept one:
// ncu showed 70 % tensor-pipe activity at 64 % L2 throughput, the load latency was exposed)
constexpr int IGT_STAGE_BYTES = IGT_ACT_BYTES + IGT_W_BYTES;   // 48 KB
constexpr int IGT_STAGES = 4;
constexpr int IGT_OFF_W = IGT_ACT_BYTES;
constexpr int IGT_SLAB_ROWS = 4;                        // patch rows per output slab
constexpr int IGT_SLAB_PIX = IGT_SLAB_ROWS * IGT_TW;    // 64 pixels
constexpr int IGT_SLAB_BYTES = IGT_SLAB_PIX * 128;      // one plane of one slab: 8 KB
constexpr int IGT_OUT_BYTES = 2 /*slots*/ * 2 /*planes*/ * IGT_SLAB_BYTES;
constexpr int IGT_BAR_BYTES = 1024;
constexpr int IGT_SMEM_BYTES = IGT_STAGES * IGT_STAGE_BYTES + IGT_BAR_BYTES + IGT_OUT_BYTES + 1024;

// Stacked weight operands of a 64-output-channel layer: wt = [2][128][K].  Row r = 32 q + 16 half + i (TMEM lane r of
// the accumulator) belongs to channel c = 16 q + (i < 8 ? 2 i : 2 (i - 8) + 1): inside a 16-lane half, lanes i and
// i + 8 hold ADJACENT channels, which is the row pair one thread receives from a 16x256b TMEM load.
//   operand 0:  half 0 -> W_hi[c] (main rows),  half 1 -> W_lo[c] (cross-term rows)
//   operand 1:  half 0 -> 0,                    half 1 -> W_hi[c]
// w = [2][64][K] (hi plane, lo plane), the layout every other kernel reads.
__global__ void igemm_t_stack_weights_kernel(const pl16* __restrict__ w, pl16* __restrict__ wt, int K) {
  const long long total = 2LL * 128 * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int r = (int)((i / K) % 128);
    const int op = (int)(i / ((long long)128 * K));
    const int half = (r >> 4) & 1, li = r & 15;
    const int c = (r >> 5) * 16 + (li < 8 ? 2 * li : 2 * (li - 8) + 1);
    const pl16 hi = w[(long long)c * K + k], lo = w[(long long)(64 + c) * K + k];
    pl16 v;
    if (op == 0) v = half ? lo : hi;
    else v = half ? hi : float_to_pl16(0.f);
    wt[i] = v;
  }
}

// 16 TMEM lanes x 32 columns -> 16 registers: r[4 n + 2 i1 + i0] = (lane base + lane / 4 + 8 i1, column 8 n + 2 (lane % 4) + i0)
// (the m16n8 accumulator-fragment pattern, cute::SM100_TMEM_LOAD_16dp256b4x); no wait
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
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
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int img, y0, x0;
        igt_tile_coords(p, tile, img, y0, x0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / k_chunks;
          const int c0 = (kb - tap * k_chunks) * IG_BLOCK_K;
          const int ax = x0 + p.tap_dx[tap], ay = y0 + p.tap_dy[tap];
          const int bk = tap * p.Cin + c0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {   // hi half stage, then lo half stage
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            const uint32_t fb = full_bar + 8 * stage;
            const uint32_t sa = smem_base + stage * IGT_STAGE_BYTES;
            mbar_expect_tx(fb, IGT_STAGE_BYTES);
            tma_load_4d(sa, &tmA, fb, c0, ax, ay, img + h * p.a_lo_img);
            tma_load_4d(sa + IGT_OFF_W, &tmW, fb, bk, 0, 0, h);
            if (++stage == IGT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
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
          // first half stage: main += W_hi*A_hi, cross += W_lo*A_hi; second: cross += W_hi*A_lo (main += 0)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(full_bar + 8 * stage, phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_base + stage * IGT_STAGE_BYTES;
            const uint64_t act = umma_desc_sw128(sa);
            const uint64_t wst = umma_desc_sw128(sa + IGT_OFF_W);
#pragma unroll
            for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
              const uint64_t ko = (uint64_t)((k * IG_UMMA_K * 2) >> 4);
              tcgen05_mma_bf16(d, wst + ko, act + ko, idesc, (kb | k | h) != 0);
            }
            tcgen05_commit(empty_bar + 8 * stage);
            if (++stage == IGT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        tcgen05_commit(tfull_bar + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM quadrant of this warp = stacked rows [32 q, 32 q + 32): lanes 0..15 main, 16..31 cross terms of channels
    // [16 q, 16 q + 16).  A 16x256b load hands thread t the rows t / 4 and t / 4 + 8 of a 16-lane half = the ADJACENT
    // channels ch0, ch0 + 1 (stacking order above) at the columns (pixels) 8 n + 2 (t % 4) + {0, 1}: main and cross
    // values of the same element arrive in the same thread, and a pixel's channel pair is one 4-byte staging store.
    const int quad = warp & 3;
    const int rp = lane >> 2, cp = lane & 3;
    const int ch0 = quad * 16 + 2 * rp;
    const bool leader = (threadIdx.x == 64);
    const uint32_t out_base = smem_base + IGT_STAGES * IGT_STAGE_BYTES + IGT_BAR_BYTES;  // 1024-aligned
    const float bias0 = p.bias ? p.bias[ch0] : 0.f, bias1 = p.bias ? p.bias[ch0 + 1] : 0.f;
    const uint32_t ch_chunk = (uint32_t)(ch0 >> 3), ch_in = (uint32_t)(ch0 & 7) * 2u;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t slab = 0;   // running slab counter: slot = slab & 1
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int img, y0, x0;
      igt_tile_coords(p, tile, img, y0, x0);
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tcgen05_fence_after();
      const uint32_t t_main = tmem_base + acc * 256 + ((uint32_t)(quad * 32) << 16);
      const uint32_t t_cross = t_main + (16u << 16);
      // patch rows below the image would be clipped by the TMA store anyway: their slabs are not computed
      const int rows_valid = p.H - y0 < IGT_TH ? p.H - y0 : IGT_TH;
      const int n_slabs = (rows_valid + IGT_SLAB_ROWS - 1) / IGT_SLAB_ROWS;
      for (int s = 0; s < n_slabs; ++s, ++slab) {
        const uint32_t sb = out_base + (slab & 1u) * (2u * IGT_SLAB_BYTES);
        // the TMA store that last read this slot (two slabs ago) must have finished reading it
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t m[16], c[16];
          tmem_ld_16x256b_x4(t_main + s * IGT_SLAB_PIX + hf * 32, m);
          tmem_ld_16x256b_x4(t_cross + s * IGT_SLAB_PIX + hf * 32, c);
          tmem_ld_wait();
          if (hf == 1 && s == n_slabs - 1) {   // accumulator drained: hand the TMEM stage back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
          }
#pragma unroll
          for (int n = 0; n < 4; ++n) {
#pragma unroll
            for (int i0 = 0; i0 < 2; ++i0) {
              float v0 = fmaf(__uint_as_float(c[4 * n + i0]), CALD_LO_INV, __uint_as_float(m[4 * n + i0])) + bias0;
              float v1 = fmaf(__uint_as_float(c[4 * n + 2 + i0]), CALD_LO_INV, __uint_as_float(m[4 * n + 2 + i0])) + bias1;
              if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              uint32_t hw, lw;
              split_pack2(v0, v1, hw, lw);
              const uint32_t px = (uint32_t)(hf * 32 + 8 * n + i0) + 2u * (uint32_t)cp;   // pixel inside the slab
              const uint32_t off = px * 128u + ((ch_chunk ^ (px & 7u)) << 4) + ch_in;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + off), "r"(hw) : "memory");
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + IGT_SLAB_BYTES + off), "r"(lw) : "memory");
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          const int ys = y0 + IGT_SLAB_ROWS * s;
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
