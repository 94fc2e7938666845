// Implicit-GEMM convolution / GEMM on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves every dense contraction of the
// detectors the CALD scoring path runs (a11 ResNet body, a12 FPN,
// a13 RPN head, a15 box head, a18 RetinaNet heads):
//   warp 0   : TMA producer  (cp.async.bulk.tensor 4-D boxes -> 128B-swizzled smem)
//   warp 1   : MMA issuer    (tcgen05.mma kind::f16, pl16 x pl16 -> fp32 in TMEM)
//   warps 2-5: epilogue      (tcgen05.ld -> bias / residual / ReLU -> split-pl16 NHWC)
// A tile = 128 output pixels (th x tw patch of one image, or 128 rows in linear
// mode) x 64 input channels; the 3x3 taps are realised as shifted TMA boxes with
// hardware zero fill at the borders, so no im2col buffer ever exists in HBM.
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// Split-pl16 ("bf16x3") arithmetic: A = A_hi + A_lo, B = B_hi + B_lo, product = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi.
// Launches with BLOCK_N <= 128 keep the small cross terms in their OWN accumulator columns (XSEP): one N = 2*BLOCK_N MMA
// multiplies A_hi by the concatenation [B_hi | B_lo] (B_lo sits right behind B_hi in the stage), a second N = BLOCK_N
// MMA adds A_lo*B_hi into the cross columns, and the epilogue sums the two halves in fp32.  That is two wide
// instructions instead of three narrow ones per k-step (the tensor pipe is ~1.6x more efficient at N = 256 than at
// N = 128, measured), and the big accumulator sees one truncating add per k-step instead of three.
#pragma once
#include "common.cuh"

namespace cald {

constexpr int IG_BLOCK_M = 128;
constexpr int IG_BLOCK_K = 64;   // pl16 elements = one 128-byte swizzle row
constexpr int IG_UMMA_K = 16;
constexpr int IG_THREADS = 192;  // 6 warps

template <int BLOCK_N, bool SPLIT>
struct IgemmCfg {
  static constexpr int A_BYTES = IG_BLOCK_M * IG_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * IG_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int OUT_STAGE_BYTES = (SPLIT ? 2 : 1) * A_BYTES;  // epilogue staging: 128 rows x 64 ch per plane
  static constexpr int BAR_BYTES = 1024;                              // barriers + tmem pointer (keeps staging aligned)
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - BAR_BYTES - OUT_STAGE_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + OUT_STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;  // 2 accumulator stages of <= 256 fp32 columns
  static constexpr int ACC_STRIDE = 256;
  // stage layout: A_hi | A_lo | B_hi | B_lo  (B_lo directly behind B_hi: together they are one 2*BLOCK_N-row operand)
  static constexpr int OFF_A_LO = A_BYTES;
  static constexpr int OFF_B_HI = (SPLIT ? 2 : 1) * A_BYTES;
  static constexpr int OFF_B_LO = OFF_B_HI + B_BYTES;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// one lane of the (converged) warp, chosen by the hardware: ptxas knows the predicate of elect.sync has exactly one
// active lane and issues single-thread instructions (tcgen05.mma / commit, TMA) under it without a per-lane loop
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024>>4 | version [46,48) = 1 | layout [61,64) = 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 accumulator columns = main + cross-term columns (XSEP): both TMEM loads are in flight before the one wait, so a
// 64-channel group costs two TMEM round trips instead of four (the epilogue of the short-K layers is latency-bound:
// ncu stall samples sat on the first FADD after every LDTM)
// The cross-term accumulator holds LO_SCALE * (A_hi*B_lo + A_lo*B_hi) (common.cuh).
__device__ __forceinline__ void tmem_ld32_sum2(uint32_t t_main, uint32_t t_cross, float* v) {
  uint32_t a[32], b[32];
  tmem_ld32_nowait(t_main, a);
  tmem_ld32_nowait(t_cross, b);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(b[i]), CALD_LO_INV, __uint_as_float(a[i]));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// 64 residual channels of one output row (8 x 16 B per plane) into registers
__device__ __forceinline__ void load_res64(const ConvParams& p, long long rrow, int c, uint4* rh, uint4* rl) {
#pragma unroll
  for (int q = 0; q < 8; ++q) rh[q] = *reinterpret_cast<const uint4*>(p.res_hi + rrow + c + q * 8);
  if (p.res_lo) {
#pragma unroll
    for (int q = 0; q < 8; ++q) rl[q] = *reinterpret_cast<const uint4*>(p.res_lo + rrow + c + q * 8);
  }
}

// tile index -> (n block, image, y0, x0).  N blocks vary fastest so that CTAs running
// concurrently share the same activation tile in L2.
__device__ __forceinline__ void tile_coords(const ConvParams& p, int tile, int& nb, int& img, int& y0, int& x0) {
  nb = tile % p.n_blocks;
  int m = tile / p.n_blocks;
  int tx = m % p.tiles_x;
  m /= p.tiles_x;
  int ty = m % p.tiles_y;
  img = m / p.tiles_y;
  y0 = ty * p.th;
  x0 = tx * p.tw;
}

// CHUNKED: the tensor core adds into the TMEM accumulator with truncation (measured: relative bias ~6e-8 per
// accumulate, toward zero), which at K = 12544 (fc6) costs 6e-5.  In chunked mode the MMA warp restarts the
// accumulator every p.kc k-blocks and the epilogue warps sum the partial tiles in fp32 registers (round to nearest),
// using the two TMEM stages as a ring, so the bias is bounded by the chunk length instead of K.
template <int BLOCK_N, bool SPLIT, bool CHUNKED>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmI, const __grid_constant__ ConvParams p) {
  using Cfg = IgemmCfg<BLOCK_N, SPLIT>;
  constexpr bool XSEP = SPLIT && BLOCK_N <= 128;  // cross terms in their own accumulator columns
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barrier layout: full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  const uint32_t full_bar = bars, empty_bar = bars + 8 * Cfg::STAGES;
  const uint32_t tfull_bar = empty_bar + 8 * Cfg::STAGES, tempty_bar = tfull_bar + 16;
  volatile uint32_t* tmem_holder =
      reinterpret_cast<volatile uint32_t*>(smem_al + Cfg::STAGES * Cfg::STAGE_BYTES + 16 * Cfg::STAGES + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmI) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
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
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int k_chunks = p.Cin / IG_BLOCK_K;
  const int conv_kb = p.taps * k_chunks;
  // Residual add on the tensor core: the shortcut tensor is streamed through the same TMA pipeline as extra
  // k-blocks and multiplied by a 64-wide identity (R_hi*I + R_lo*I is exact), so the epilogue never waits on
  // strided residual loads.  res_kb = BLOCK_N / 64 for such layers, else 0.  With p.res_conv the extra k-blocks are a
  // second 1x1 contraction instead (the projection shortcut of a stage's first bottleneck, aux_cin / 64 of them):
  // relu(W3*y + Wd*x + b3 + bd) is ONE accumulation, so the 4x-wide shortcut tensor is never written or re-read.
  const int num_kb = conv_kb + p.res_kb;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int nb, img, y0, x0;
        tile_coords(p, tile, nb, img, y0, x0);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          if (kb >= conv_kb && p.res_conv) {
            // second contraction (downsample shortcut): aux activations through tmR, aux weights through tmI
            const int j = kb - conv_kb;
            const int rx = x0 * p.r_scale, ry = y0 * p.r_scale, bn = nb * BLOCK_N;
            mbar_expect_tx(fb, (SPLIT ? 2 : 1) * (p.a_bytes + Cfg::B_BYTES));
            tma_load_4d(sa, &tmR, fb, j * 64, rx, ry, img);
            tma_load_4d(sa + Cfg::OFF_B_HI, &tmI, fb, j * 64, bn, 0, 0);
            if (SPLIT) {
              tma_load_4d(sa + Cfg::OFF_A_LO, &tmR, fb, j * 64, rx, ry, img + p.r_lo_img);
              tma_load_4d(sa + Cfg::OFF_B_LO, &tmI, fb, j * 64, bn, 0, 1);
            }
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (kb >= conv_kb) {
            const int j = kb - conv_kb;
            const int rc = nb * BLOCK_N + j * 64;
            mbar_expect_tx(fb, (SPLIT ? 2 : 1) * p.a_bytes + Cfg::B_BYTES);
            tma_load_4d(sa, &tmR, fb, rc, x0, y0, img);
            tma_load_4d(sa + Cfg::OFF_B_HI, &tmI, fb, 0, j * BLOCK_N, 0, 0);
            if (SPLIT) tma_load_4d(sa + Cfg::OFF_A_LO, &tmR, fb, rc, x0, y0, img + p.r_lo_img);
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          const int tap = kb / k_chunks;
          const int c0 = (kb - tap * k_chunks) * IG_BLOCK_K;
          mbar_expect_tx(fb, (SPLIT ? 2 : 1) * (p.a_bytes + Cfg::B_BYTES));
          const int ax = x0 * p.a_scale + p.tap_dx[tap], ay = y0 * p.a_scale + p.tap_dy[tap], ai = img + p.tap_img[tap];
          const int bk = tap * p.Cin + c0, bn = nb * BLOCK_N;
          tma_load_4d(sa, &tmA, fb, c0, ax, ay, ai);
          tma_load_4d(sa + Cfg::OFF_B_HI, &tmB, fb, bk, bn, 0, 0);
          if (SPLIT) {
            tma_load_4d(sa + Cfg::OFF_A_LO, &tmA, fb, c0, ax, ay, ai + p.a_lo_img);
            tma_load_4d(sa + Cfg::OFF_B_LO, &tmB, fb, bk, bn, 0, 1);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      const uint32_t idesc = (1u << 4) | (PL16_MMA_FMT << 7) | (PL16_MMA_FMT << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                             ((uint32_t)(IG_BLOCK_M >> 4) << 24);
      // XSEP: the same instruction shape with N = 2 * BLOCK_N over [B_hi | B_lo]
      const uint32_t idesc_cat = (1u << 4) | (PL16_MMA_FMT << 7) | (PL16_MMA_FMT << 10) |
                                 ((uint32_t)((2 * BLOCK_N) >> 3) << 17) | ((uint32_t)(IG_BLOCK_M >> 4) << 24);
      const uint32_t idesc64 = (1u << 4) | (PL16_MMA_FMT << 7) | (PL16_MMA_FMT << 10) | ((uint32_t)(64 >> 3) << 17) |
                               ((uint32_t)(IG_BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int kc = CHUNKED ? p.kc : num_kb;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        uint32_t d = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int kin = kb % kc;  // position inside the accumulation chunk
          if (kin == 0) {
            mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
            tcgen05_fence_after();
            d = tmem_base + acc * Cfg::ACC_STRIDE;
          }
          mbar_wait(full_bar + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(sa);
          const uint64_t b_hi = umma_desc_sw128(sa + Cfg::OFF_B_HI);
          const uint64_t a_lo = umma_desc_sw128(sa + Cfg::OFF_A_LO);
          const uint64_t b_lo = umma_desc_sw128(sa + Cfg::OFF_B_LO);
#pragma unroll
          for (int k = 0; k < IG_BLOCK_K / IG_UMMA_K; ++k) {
            const uint64_t ko = (uint64_t)((k * IG_UMMA_K * 2) >> 4);  // advance start address inside the swizzle row
            if (kb >= conv_kb && !p.res_conv) {
              // residual k-block: main += R_hi * I, cross += R_lo * I (the lo plane carries LO_SCALE like the cross terms)
              if (XSEP) {
                // block j of the identity routes 64 residual channels to accumulator columns [64 j, 64 j + 64): only its
                // rows [64 j, 64 j + 64) are non-zero, so an N = 64 instruction on those rows and columns adds the same
                // values as the N = BLOCK_N one (the other columns received + 0) at ~2/3 of its tensor-pipe time
                if (p.res_narrow) {
                  const int j = kb - conv_kb;
                  const uint64_t bj = b_hi + (uint64_t)((j * 64 * 128) >> 4) + ko;
                  tcgen05_mma_bf16(d + BLOCK_N + j * 64, a_lo + ko, bj, idesc64, 1);
                  tcgen05_mma_bf16(d + j * 64, a_hi + ko, bj, idesc64, 1);
                } else {
                  tcgen05_mma_bf16(d + BLOCK_N, a_lo + ko, b_hi + ko, idesc, 1);
                  tcgen05_mma_bf16(d, a_hi + ko, b_hi + ko, idesc, 1);
                }
              } else {
                if (SPLIT) tcgen05_mma_bf16(d, a_lo + ko, b_hi + ko, idesc, (kin | k) != 0);
                tcgen05_mma_bf16(d, a_hi + ko, b_hi + ko, idesc, SPLIT ? 1u : (uint32_t)((kin | k) != 0));
              }
            } else if (XSEP) {
              // columns [0, BLOCK_N) += A_hi*B_hi, columns [BLOCK_N, 2*BLOCK_N) += A_hi*B_lo + A_lo*B_hi
              tcgen05_mma_bf16(d, a_hi + ko, b_hi + ko, idesc_cat, (kin | k) != 0);
              tcgen05_mma_bf16(d + BLOCK_N, a_lo + ko, b_hi + ko, idesc, 1);
            } else if (SPLIT) {
              // small cross terms first, dominant term last
              tcgen05_mma_bf16(d, a_lo + ko, b_hi + ko, idesc, (kin | k) != 0);
              tcgen05_mma_bf16(d, a_hi + ko, b_lo + ko, idesc, 1);
              tcgen05_mma_bf16(d, a_hi + ko, b_hi + ko, idesc, 1);
            } else {
              tcgen05_mma_bf16(d, a_hi + ko, b_hi + ko, idesc, (kin | k) != 0);
            }
          }
          tcgen05_commit(empty_bar + 8 * stage);  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          if (kin == kc - 1 || kb == num_kb - 1) {
            tcgen05_commit(tfull_bar + 8 * acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const bool leader = (threadIdx.x == 64);
    const uint32_t stage_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES;  // 1024-aligned
    bool store_pending = false;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int num_chunks = CHUNKED ? (num_kb + p.kc - 1) / p.kc : 1;
    constexpr int NACC = CHUNKED ? (BLOCK_N <= 128 ? BLOCK_N : 32) : 32;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int nb, img, y0, x0;
      tile_coords(p, tile, nb, img, y0, x0);
      const int y = y0 + row / p.tw, x = x0 + row % p.tw;
      const bool valid = (row < p.th * p.tw) && (y < p.H) && (x < p.W);
      long long orow = 0, rrow = 0;
      if (valid) {
        orow = out_row_offset(p, img, y, x);
        if (p.res_mode != RES_NONE) rrow = res_row_offset(p, img, y, x);
      }
      // residual of the first 64-channel group is requested before the accumulator is ready
      uint4 rh[8], rl[8];
      const bool res_on = (p.res_mode != RES_NONE) && valid;
      if (p.tma_store && res_on) load_res64(p, rrow, nb * BLOCK_N, rh, rl);

      float accv[NACC];
      if (CHUNKED) {
        for (int ch = 0; ch < num_chunks; ++ch) {
          mbar_wait(tfull_bar + 8 * acc, acc_phase);
          tcgen05_fence_after();
          const uint32_t t0 = tmem_base + acc * Cfg::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
#pragma unroll
          for (int cc = 0; cc < NACC; cc += 32) {
            uint32_t r[32];
            tmem_ld32(t0 + cc, r);
            if (ch == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) accv[cc + i] = __uint_as_float(r[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) accv[cc + i] = accv[cc + i] + __uint_as_float(r[i]);
            }
            if (XSEP) {  // the chunk's cross-term columns
              tmem_ld32(t0 + BLOCK_N + cc, r);
#pragma unroll
              for (int i = 0; i < 32; ++i) accv[cc + i] = fmaf(__uint_as_float(r[i]), CALD_LO_INV, accv[cc + i]);
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        mbar_wait(tfull_bar + 8 * acc, acc_phase);
        tcgen05_fence_after();
      }
      const uint32_t t0 = tmem_base + acc * Cfg::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);

#pragma unroll
      for (int g = 0; g < BLOCK_N / 64; ++g) {
        const int c0 = nb * BLOCK_N + g * 64;
        float v[64];
        if (CHUNKED) {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = accv[(g * 64 + i) % NACC];
        } else {
          if (XSEP) {
            tmem_ld32_sum2(t0 + g * 64, t0 + BLOCK_N + g * 64, v);
            tmem_ld32_sum2(t0 + g * 64 + 32, t0 + BLOCK_N + g * 64 + 32, v + 32);
          } else {
            uint32_t r[32];
            tmem_ld32(t0 + g * 64, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            tmem_ld32(t0 + g * 64 + 32, r);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[32 + i] = __uint_as_float(r[i]);
          }
          if (g == BLOCK_N / 64 - 1) {  // accumulator fully drained: hand the TMEM stage back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
          }
        }
        if (!p.tma_store) {
          // direct path: fp32 head outputs, Cout not a multiple of 64
          if (valid) {
#pragma unroll
            for (int j = 0; j < 64; j += 8) {
              if (c0 + j < p.Cout) epilogue_store8(p, orow, rrow, c0 + j, &v[j]);
            }
          }
          continue;
        }
        if (c0 >= p.Cout) continue;  // uniform across the CTA
        // ---- bias + residual + ReLU in registers
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + c0 + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (res_on) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t* h = reinterpret_cast<const uint32_t*>(&rh[q]);
            const uint32_t* l = reinterpret_cast<const uint32_t*>(&rl[q]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float a, b;
              if (SPLIT) {
                join_pack2(h[i], l[i], a, b);
              } else {
                unpack2(h[i], a, b);
              }
              v[q * 8 + 2 * i] += a;
              v[q * 8 + 2 * i + 1] += b;
            }
          }
          // request the next group's residual now; it lands while this group is being stored
          if (g + 1 < BLOCK_N / 64 && c0 + 64 < p.Cout) load_res64(p, rrow, c0 + 64, rh, rl);
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        // ---- wait until the previous TMA store has finished reading the staging buffer
        if (store_pending) {
          if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        // ---- split to pl16 planes and write this row's 8 x 16-byte chunks at their 128B-swizzled positions
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_pack2(v[q * 8 + 2 * i], v[q * 8 + 2 * i + 1], ph[i], pl[i]);
          const uint32_t off = (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_base + off), "r"(ph[0]), "r"(ph[1]),
                       "r"(ph[2]), "r"(ph[3])
                       : "memory");
          if (SPLIT)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_base + Cfg::A_BYTES + off), "r"(pl[0]),
                         "r"(pl[1]), "r"(pl[2]), "r"(pl[3])
                         : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {
          tma_store_4d(&tmC, stage_base, c0, x0, y0, img);
          if (SPLIT) tma_store_4d(&tmC, stage_base + Cfg::A_BYTES, c0, x0, y0, img + p.c_lo_img);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        store_pending = true;
      }
      if (!CHUNKED) {
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    if (leader && store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// --------------------------------------------------------------------------------
// SIMT checker: the same contraction with plain fp32 FMAs on the same split inputs.
// Used by the stage-wise parity tests to validate the tensor-core kernel and as the
// engine's CALD_CONV=simt debugging path.  One thread = one pixel x 8 output channels.
// --------------------------------------------------------------------------------
struct SimtOperands {
  const pl16* a_hi; const pl16* a_lo;   // [img'][H_in][W_in][Cin]
  const pl16* b_hi; const pl16* b_lo;   // [Cout_pad][taps*Cin]
  int H_in, W_in, n_img_in;             // input plane geometry (a_lo == null -> pl16 mode)
  int pix_stride;                       // elements between consecutive pixels (Cin, or 16 for the stem window)
  int w_limit;                          // number of valid window start columns (W_in, or W_in - 3 for the stem)
};

__global__ void conv_simt_kernel(ConvParams p, SimtOperands o) {
  const int cgroups = (p.Cout + 7) / 8;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)p.n_img * p.H * p.W * cgroups;
  if (gid >= total) return;
  int cg = (int)(gid % cgroups);
  long long pix = gid / cgroups;
  int x = (int)(pix % p.W);
  int y = (int)((pix / p.W) % p.H);
  int n = (int)(pix / ((long long)p.W * p.H));
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int ktot = p.taps * p.Cin;
  for (int tap = 0; tap < p.taps; ++tap) {
    int iy = y * p.a_scale + p.tap_dy[tap], ix = x * p.a_scale + p.tap_dx[tap], ii = n + p.tap_img[tap];
    if (iy < 0 || iy >= o.H_in || ix < 0 || ix >= o.w_limit) continue;
    const long long abase = (((long long)ii * o.H_in + iy) * o.W_in + ix) * o.pix_stride;
    for (int c = 0; c < p.Cin; ++c) {
      float a = pl16_to_float(o.a_hi[abase + c]);
      if (o.a_lo) a = join_pl(o.a_hi[abase + c], o.a_lo[abase + c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long bi = (long long)(cg * 8 + j) * ktot + tap * p.Cin + c;
        float b = pl16_to_float(o.b_hi[bi]);
        if (o.b_lo) b = join_pl(o.b_hi[bi], o.b_lo[bi]);
        acc[j] = fmaf(a, b, acc[j]);
      }
    }
  }
  long long orow = out_row_offset(p, n, y, x);
  long long rrow = p.res_mode != RES_NONE ? res_row_offset(p, n, y, x) : 0;
  epilogue_store8(p, orow, rrow, cg * 8, acc);
}

}  // namespace cald
