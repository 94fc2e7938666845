// Host side of the implicit-GEMM engine: tensor-map construction, tile choice, launch.
#pragma once
#include <atomic>
#include <mutex>
#include <cstring>
#include <vector>
#include <cstdlib>
#include <map>
#include <string>
#include "igemm.cuh"
#include "igemm2.cuh"
#include "igemm_t.cuh"

// Relative loss of the fp32 TMEM accumulator per tcgen05.mma accumulate.  The tensor core adds into the accumulator
// with truncation (round toward zero): each of the n MMA k-steps of a contraction shrinks the running sum by c on
// average, so a product that enters at step j is shrunk (n - j) times -- measured on B200 as a signed bias of
// -(1.6e-8 * n) of the result for every K from 64 to 12544 (tools/conv_accuracy.py, profiles/r02_conv_accuracy.txt),
// i.e. c = 3.2e-8.  Coherent over ~50 layers this was the engine's largest deviation from an fp32 CPU run.  It is
// removed where it costs nothing: the WEIGHTS of k-step j are uploaded multiplied by 1 + c * (n - j) (RzPlan).
#ifndef CALD_RZ_C_DEFAULT
#define CALD_RZ_C_DEFAULT 3.2e-8
#endif

namespace cald {

// ---- split-pl16 NHWC activation: planes [hi | lo], each [n][h][w][c]
struct Act {
  pl16* hi = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  bool split = true;
  size_t plane_elems() const { return (size_t)n * h * w * c; }
  pl16* lo() const { return split ? hi + plane_elems() : nullptr; }
  size_t bytes() const { return plane_elems() * (split ? 2 : 1) * sizeof(pl16); }
};

// ---- conv / linear weights on device: [2][cout_pad][taps*cin] pl16 (BN already folded) + fp32 bias
struct ConvW {
  pl16* w = nullptr;
  float* bias = nullptr;
  int cout = 0, cout_pad = 0, cin = 0, taps = 1;
  // [2][128][taps*cin]: the stacked operands of the transposed-role kernel (igemm_t.cuh), built on first use for the
  // 64-output-channel layers that take it
  mutable pl16* wt = nullptr;
  size_t plane_elems() const { return (size_t)cout_pad * taps * cin; }
};

// Where a weight tensor's k-steps sit in the accumulation of the launch that uses it (see CALD_RZ_C_DEFAULT):
// steps_before = MMA k-steps (16 K-elements each) accumulated before this tensor's first one, steps_total = all real
// k-steps of the launch (a following identity-routed residual block counts as one), chunk_steps = accumulator restart
// period of a chunked launch (0 = one accumulation).
struct RzPlan {
  int steps_before = 0;
  int steps_total = 0;     // 0 = derive from the tensor itself: taps * cin / 16
  int chunk_steps = -1;    // -1 = derive from the engine's chunking rule
  double c = -1.0;         // < 0 = CALD_RZ_C (environment) or the measured default
  // multiplier of the weights that enter at k-step j of the tensor
  double factor(int j_local, int own_steps) const {
    const double cc = c >= 0 ? c : rz_c();
    const int total = steps_total > 0 ? steps_total : own_steps;
    const int j = steps_before + j_local;
    if (chunk_steps > 0) {
      const int jj = j % chunk_steps, chunk_len = std::min(chunk_steps, total - (j - jj));
      return 1.0 + cc * (double)(chunk_len - jj);
    }
    return 1.0 + cc * (double)(total - j);
  }
  static double rz_c() {   // read at every weight upload: experiments can change it between engines of one process
    const char* e = getenv("CALD_RZ_C");
    return e && *e ? atof(e) : CALD_RZ_C_DEFAULT;
  }
};

struct ConvOpts {
  bool relu = false;
  int stride = 1;               // 3x3: 1 or 2 (2: the A tensor map walks the full-resolution input with element stride 2)
  int res_mode = RES_NONE;
  const Act* res = nullptr;
  float* out_f32 = nullptr;     // fp32 NHWC output instead of / in addition to split pl16
  bool no_bf16_out = false;
  // stem mode: `in` is the padded space-to-depth image [n][Ho+3][Wo+3][16]; the 7x7/s2 conv is a 4x4/s1 conv whose
  // k-block `dy` is the 64-element window (4 pixels x 16 ch) starting at pixel (oy + dy, ox): an overlapping-stride
  // tensor map (pixel stride 32 B, box width 128 B) lets TMA gather it without an im2col buffer.
  bool stem_window = false;
  // 1x1 conv with stride 2 (ResNet downsample shortcut, tv:models/resnet.py): `in` is the full-resolution tensor and
  // the TMA tensor map itself skips every other pixel (element strides 2, 2), so no subsampled copy is materialised
  bool in_stride2 = false;
  // second 1x1 contraction accumulated into the same output tile: out = act(conv(in, w) + conv1x1(aux_in, aux_w) + bias),
  // aux_in read with element stride aux_stride (the projection shortcut of tv:models/resnet.py Bottleneck.downsample).
  // bias_sum = device vector of the two folded biases added together.
  const Act* aux_in = nullptr;
  const ConvW* aux_w = nullptr;
  int aux_stride = 1;
  const float* bias_sum = nullptr;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CALD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    if (!p || qr != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled unavailable");
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 4-D pl16 tensor map, innermost box = 64 elements (128 B) with 128B swizzle.
inline CUtensorMap make_tmap(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t b1,
                             uint32_t b2, uint32_t estride = 1) {
  CUtensorMap tm;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {d0 * 2, d0 * d1 * 2, d0 * d1 * d2 * 2};
  // with an element stride s the box is given in un-strided elements: s * n of them deliver n
  cuuint32_t box[4] = {64, b1 * estride, b2 * estride, 1};
  cuuint32_t es[4] = {1, estride, estride, 1};
  CUresult r = get_encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu,%llu box=%u,%u", (int)r,
             (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3, b1, b2);
    throw std::runtime_error(b);
  }
  return tm;
}

// Same, with explicit byte strides for dims 1..3 (used by the stem's overlapping sliding-window view).
inline CUtensorMap make_tmap_strided(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                                     uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b1, uint32_t b2) {
  CUtensorMap tm;
  cuuint64_t dims[4] = {d0, d1, d2, d3};
  cuuint64_t strides[3] = {s1, s2, s3};
  cuuint32_t box[4] = {64, b1, b2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = get_encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[256];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled(strided) failed (%d) dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu",
             (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3,
             (unsigned long long)s1, (unsigned long long)s2, (unsigned long long)s3);
    throw std::runtime_error(b);
  }
  return tm;
}

// Spatial tile (th x tw <= 128 output pixels).  TMA boxes need not be powers of two, so pick the shape that wastes
// the fewest of the 128 MMA rows (e.g. 3 x 42 = 126 rows on an 84-wide map instead of 4 x 32 with 14 % padding).
inline void choose_tile(int H, int W, int& th, int& tw) {
  double best = -1;
  th = 8; tw = 16;
  for (int w = 1; w <= 128 && w <= W; ++w) {
    int h = 128 / w;
    if (h > H) h = H;
    if (h < 1) continue;
    long long tiles = (long long)((H + h - 1) / h) * ((W + w - 1) / w);
    double eff = (double)H * W / ((double)tiles * 128.0);
    // tie-break towards squarer tiles (better halo reuse in L2 for 3x3 taps)
    double score = eff - 1e-6 * std::abs(h - w);
    if (score > best) { best = score; th = h; tw = w; }
  }
}

enum ConvImpl { CONV_TC = 0, CONV_SIMT = 1 };

// launches of the CTA-pair kernel in this process (lets the parity tests assert which kernel they exercised)
inline std::atomic<long long>& pair_launch_counter() {
  static std::atomic<long long> n{0};  // independent engine handles may launch from different host threads
  return n;
}

inline std::atomic<long long>& tform_launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}

struct ConvEngine {
  int num_sms = 148;
  ConvImpl impl = CONV_TC;
  bool split = true;       // false: single-pass pl16 (hi plane only)
  int force_block_n = 0;   // 0 = auto
  bool use_tma_store = true;  // false: always use the direct (per-thread) store epilogue
  bool use_res_mma = true;    // false: add residuals in the epilogue registers
  bool force_pow2_tiles = false;  // true: restrict spatial tiles to power-of-two shapes
  // CTA-pair kernel (igemm2.cuh) for the split-mode BLOCK_N = 128 launches it covers; CALD_CTA2=0/1 overrides
  bool use_cta2 = env_flag("CALD_CTA2", true);
  int cta2_min_kb = env_int("CALD_CTA2_MIN_KB", 16);
  int cta2_min_kb64 = env_int("CALD_CTA2_MIN_KB64", 9);
  // transposed-role kernel (igemm_t.cuh) for the spatial 64-output-channel layers (stem, layer1 3x3); CALD_TFORM=0
  // falls back to the BLOCK_N = 64 instantiations above
  bool use_tform = env_flag("CALD_TFORM", true);
  // identity-routed residual k-blocks as N = 64 instructions on the 64 accumulator columns they feed (CALD_RES_NARROW=0:
  // N = BLOCK_N over the whole identity block; same sums, the other columns receive + 0)
  bool res_narrow = env_flag("CALD_RES_NARROW", true);
  static double env_double(const char* name, double dflt) {
    const char* v = getenv(name);
    return v && *v ? atof(v) : dflt;
  }
  // chunking rule of run(), for callers that need to know it before the launch (weight upload, RzPlan)
  static bool chunked_for(int num_kb, int block_n_max128 = 1) {
    const int kc_ = env_int("CALD_KC", 8), above = env_int("CALD_CHUNK_ABOVE_KB", 40);
    return kc_ > 0 && num_kb > kc_ && num_kb > above && block_n_max128;
  }
  static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
  }
  static bool env_flag(const char* name, bool dflt) {
    const char* v = getenv(name);
    return v && *v ? (*v != '0') : dflt;
  }
  pl16* ident[3] = {nullptr, nullptr, nullptr};  // identity B tiles for BLOCK_N = 64 / 128 / 256
  // I[j][n][k] = (n == j*64 + k): block j routes residual channels [j*64, j*64+64) to accumulator columns
  const pl16* identity(int BN) {
    int slot = BN == 64 ? 0 : (BN == 128 ? 1 : 2);
    if (!ident[slot]) {
      std::vector<pl16> h((size_t)(BN / 64) * BN * 64);
      for (int j = 0; j < BN / 64; ++j)
        for (int n = 0; n < BN; ++n)
          for (int k = 0; k < 64; ++k) h[((size_t)j * BN + n) * 64 + k] = float_to_pl16(n == j * 64 + k ? 1.f : 0.f);
      CALD_CUDA_CHECK(cudaMalloc((void**)&ident[slot], h.size() * sizeof(pl16)));
      CALD_CUDA_CHECK(cudaMemcpy(ident[slot], h.data(), h.size() * sizeof(pl16), cudaMemcpyHostToDevice));
    }
    return ident[slot];
  }
  int kc = env_int("CALD_KC", 8);  // split mode: k-blocks (of 64) per accumulation chunk; 0 = never chunk
  // chunk only contractions longer than this many k-blocks (K > 2560): up to there the cross-term-separated
  // accumulator (igemm.cuh XSEP) with the truncation pre-compensated in the weights (RzPlan) is as accurate and faster
  int chunk_above_kb = env_int("CALD_CHUNK_ABOVE_KB", 40);
  long long launches = 0;  // kernels launched (bench's gpu_launches)
  double flops = 0;        // algorithmic 2*MAC of the launches
  // optional per-launch timing with CUDA events on the launching stream (bench.py roofline)
  bool profiling = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  double prof_flops = 0;
  long long prof_launches = 0;
  // per-launch description (aligned with the event pairs) and the per-layer aggregate of the last drain
  struct LayerRec { char sig[96]; double flops, bytes; };
  struct LayerAgg { long long count = 0; double ms = 0, flops = 0, bytes = 0; };
  std::vector<LayerRec> recs;
  std::map<std::string, LayerAgg> layer_agg;
  cudaEvent_t next_event() {
    if (ev_used == ev.size()) {
      cudaEvent_t e;
      CALD_CUDA_CHECK(cudaEventCreate(&e));
      ev.push_back(e);
    }
    return ev[ev_used++];
  }
  // sum of kernel durations (ms) since the last call; the stream must be idle
  double drain_profile(double* fl, long long* n) {
    double ms = 0;
    layer_agg.clear();
    for (size_t i = 0; i + 1 < ev_used; i += 2) {
      float t = 0;
      CALD_CUDA_CHECK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
      ms += t;
      if (i / 2 < recs.size()) {
        const LayerRec& r = recs[i / 2];
        LayerAgg& a = layer_agg[r.sig];
        a.count++; a.ms += t; a.flops += r.flops; a.bytes += r.bytes;
      }
    }
    if (fl) *fl = prof_flops;
    if (n) *n = prof_launches;
    ev_used = 0; prof_flops = 0; prof_launches = 0;
    recs.clear();
    return ms;
  }

  template <int BN, bool SP, bool CH>
  void launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                 const CUtensorMap& ti, const ConvParams& p, cudaStream_t st) {
    using Cfg = IgemmCfg<BN, SP>;
    static std::once_flag attr_once;  // independent engine handles may launch from different host threads
    std::call_once(attr_once, [] {
      CALD_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc_kernel<BN, SP, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
    });
    int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    if (profiling) CALD_CUDA_CHECK(cudaEventRecord(next_event(), st));
    igemm_tc_kernel<BN, SP, CH><<<grid, IG_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, tc, tr, ti, p);
    if (profiling) CALD_CUDA_CHECK(cudaEventRecord(next_event(), st));
    CALD_CUDA_CHECK(cudaGetLastError());
  }

  template <int BN, bool CH>
  void launch_tc2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tbh, const CUtensorMap& tc,
                  const ConvParams& p, cudaStream_t st) {
    using Cfg = Igemm2Cfg<BN>;
    static std::once_flag attr_once;
    std::call_once(attr_once, [] {
      CALD_CUDA_CHECK(cudaFuncSetAttribute(igemm_tc2_kernel<BN, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
    });
    const int m_tiles = p.tiles_x * p.tiles_y * p.n_img;
    const int n_pairs = p.n_blocks * ((m_tiles + 1) / 2);
    const int clusters = n_pairs < num_sms / 2 ? n_pairs : num_sms / 2;
    if (profiling) CALD_CUDA_CHECK(cudaEventRecord(next_event(), st));
    igemm_tc2_kernel<BN, CH><<<2 * clusters, IG_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, tbh, tc, p);
    pair_launch_counter()++;
    if (profiling) CALD_CUDA_CHECK(cudaEventRecord(next_event(), st));
    CALD_CUDA_CHECK(cudaGetLastError());
  }

  void launch_t(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const ConvParams& p,
                cudaStream_t st) {
    static std::once_flag attr_once;
    std::call_once(attr_once, [] {
      CALD_CUDA_CHECK(cudaFuncSetAttribute(igemm_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IGT_SMEM_BYTES));
    });
    int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    if (profiling) CALD_CUDA_CHECK(cudaEventRecord(next_event(), st));
    igemm_t_kernel<<<grid, IG_THREADS, IGT_SMEM_BYTES, st>>>(ta, tw, tc, p);
    tform_launch_counter()++;
    if (profiling) CALD_CUDA_CHECK(cudaEventRecord(next_event(), st));
    CALD_CUDA_CHECK(cudaGetLastError());
  }

  // out must be pre-shaped by the caller (n, h, w, c = cout_pad or larger ldc).
  void run(const Act& in, const ConvW& w, Act& out, const ConvOpts& o, cudaStream_t st) {
    if (o.stem_window) {
      if (in.c != 16 || w.cin != 64 || w.taps != 4) throw std::runtime_error("conv: bad stem operands");
    } else if (in.c != w.cin || (w.cin % 64) != 0) {
      throw std::runtime_error("conv: Cin mismatch or not a multiple of 64");
    }
    if (in.split != split) throw std::runtime_error("conv: activation precision mode mismatch");
    ConvParams p;
    memset(&p, 0, sizeof(p));
    const bool dual = o.aux_in != nullptr;
    if (dual) {
      if (!o.aux_w || !o.bias_sum || o.aux_w->taps != 1 || (o.aux_w->cin % 64) != 0 || o.aux_in->c != o.aux_w->cin ||
          o.aux_w->cout_pad != w.cout_pad || o.res_mode != RES_NONE || o.aux_in->split != split || o.aux_in->n != in.n ||
          (o.aux_in->h + o.aux_stride - 1) / o.aux_stride != out.h || (o.aux_in->w + o.aux_stride - 1) / o.aux_stride != out.w ||
          impl != CONV_TC)
        throw std::runtime_error("conv: bad operands for the fused shortcut contraction");
    }
    const bool spatial = (w.taps == 9) || o.res_mode == RES_NEAREST || o.stem_window || o.in_stride2 ||
                         (dual && o.aux_stride != 1);
    p.a_scale = o.in_stride2 ? 2 : 1;
    p.Cin = w.cin;
    p.taps = w.taps;
    p.Cout = w.cout_pad;
    int n_img_in;
    if (spatial) {
      p.n_img = out.n;
      p.H = out.h;
      p.W = out.w;
      n_img_in = in.n;
      if (o.stem_window) {
        for (int t = 0; t < 4; ++t) { p.tap_dy[t] = t; p.tap_dx[t] = 0; p.tap_img[t] = 0; }
      } else if (w.taps == 1) {
        if (o.in_stride2 && (out.h != (in.h + 1) / 2 || out.w != (in.w + 1) / 2))
          throw std::runtime_error("conv: stride-2 1x1 output shape mismatch");
        p.tap_dy[0] = p.tap_dx[0] = p.tap_img[0] = 0;
      } else if (o.stride == 2) {
        // stride-2 3x3 straight from the plain NHWC tensor: the A tensor map walks it with element stride 2, tap (r, s)
        // starts at input pixel (2*y0 + r - 1, 2*x0 + s - 1); out-of-range pixels are TMA zero fill == padding 1
        p.a_scale = 2;
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s) {
            p.tap_dy[r * 3 + s] = r - 1;
            p.tap_dx[r * 3 + s] = s - 1;
            p.tap_img[r * 3 + s] = 0;
          }
      } else {
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s) {
            p.tap_dy[r * 3 + s] = r - 1;
            p.tap_dx[r * 3 + s] = s - 1;
            p.tap_img[r * 3 + s] = 0;
          }
      }
      choose_tile(p.H, p.W, p.th, p.tw);
      if (force_pow2_tiles) {
        static const int cand[8][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
        long long bestA = -1;
        for (auto& c : cand) {
          long long area = (long long)((p.H + c[0] - 1) / c[0]) * c[0] * ((p.W + c[1] - 1) / c[1]) * c[1];
          if (bestA < 0 || area < bestA) { bestA = area; p.th = c[0]; p.tw = c[1]; }
        }
      }
      p.tiles_x = (p.W + p.tw - 1) / p.tw;
      p.tiles_y = (p.H + p.th - 1) / p.th;
      p.a_lo_img = in.n;
    } else {
      // linear: all pixels of all images are rows of one [M][K] matrix
      p.n_img = 1;
      p.H = 1;
      long long M = (long long)in.n * in.h * in.w;
      p.W = (int)M;
      p.th = 1; p.tw = 128;
      p.tiles_x = (int)((M + 127) / 128);
      p.tiles_y = 1;
      p.a_lo_img = 1;
      n_img_in = 1;
    }
    int BN = force_block_n ? force_block_n : (w.cout_pad <= 64 ? 64 : 128);
    if (!force_block_n && !split && w.cout_pad % 256 == 0) BN = 256;
    // spatial convs with 64 output channels (stem, layer1 3x3): channels on M (hi and lo weight planes stacked), a
    // 16 x 16 pixel patch on N -- two N = 256 instructions per k-step instead of N = 128 + N = 64 (igemm_t.cuh).
    // Measured inside the cfg-2 step (profiles/r02_tform.md): layer1 3x3 708 -> 573 us, stem 1469 -> 1315 us per
    // 32-view launch.  CALD_TFORM_STEM=0 keeps the stem on the pixel-major kernel (A/B).
    static const bool tform_stem = env_flag("CALD_TFORM_STEM", true);
    const bool tform = use_tform && impl == CONV_TC && split && spatial && !force_block_n && w.cout_pad == 64 &&
                       out.c == 64 && (w.taps == 9 || (o.stem_window && tform_stem)) && o.stride == 1 && !o.in_stride2 && !dual &&
                       o.res_mode == RES_NONE && !o.no_bf16_out && o.out_f32 == nullptr && use_tma_store &&
                       !(kc > 0 && w.taps * (w.cin / 64) > kc && w.taps * (w.cin / 64) > chunk_above_kb);
    if (tform) {
      p.th = IGT_TH; p.tw = IGT_TW;
      p.tiles_x = (p.W + p.tw - 1) / p.tw;
      p.tiles_y = (p.H + p.th - 1) / p.th;
    }
    p.a_bytes = p.th * p.tw * 128;
    p.n_blocks = (w.cout_pad + BN - 1) / BN;
    p.num_tiles = p.n_blocks * p.tiles_x * p.tiles_y * p.n_img;
    p.bias = dual ? o.bias_sum : w.bias;
    p.r_scale = 1;
    p.relu = o.relu ? 1 : 0;
    p.res_mode = o.res_mode;
    if (o.res_mode != RES_NONE) {
      p.res_hi = o.res->hi;
      p.res_lo = o.res->lo();
      p.res_H = o.res->h;
      p.res_W = o.res->w;
      p.res_ld = o.res->c;
    }
    p.out_hi = o.no_bf16_out ? nullptr : out.hi;
    p.out_lo = o.no_bf16_out ? nullptr : out.lo();
    p.out_f32 = o.out_f32;
    p.ldc = out.c;
    launches++;
    // algorithmic FLOPs of the reference op (the stem's 4x4x16 window holds the 7x7x3 = 147 real taps)
    const double fl = 2.0 * (double)p.n_img * p.H * p.W * (double)w.cout *
                      ((o.stem_window ? 147.0 : (double)w.taps * w.cin) + (dual ? (double)o.aux_w->cin : 0.0));
    flops += fl;
    if (profiling && impl == CONV_TC) { prof_flops += fl; prof_launches++; }

    if (impl == CONV_SIMT) {
      SimtOperands so;
      so.a_hi = in.hi; so.a_lo = in.lo();
      so.b_hi = w.w; so.b_lo = split ? w.w + w.plane_elems() : nullptr;
      if (spatial) { so.H_in = in.h; so.W_in = in.w; } else { so.H_in = 1; so.W_in = p.W; }
      so.n_img_in = n_img_in;
      so.pix_stride = o.stem_window ? 16 : w.cin;
      so.w_limit = o.stem_window ? in.w - 3 : so.W_in;
      long long total = (long long)p.n_img * p.H * p.W * ((p.Cout + 7) / 8);
      int blocks = (int)((total + 127) / 128);
      conv_simt_kernel<<<blocks, 128, 0, st>>>(p, so);
      CALD_CUDA_CHECK(cudaGetLastError());
      return;
    }
    CUtensorMap ta, tb;
    if (o.stem_window)
      ta = make_tmap_strided(in.hi, 64, (uint64_t)in.w - 3, in.h, (uint64_t)in.n * (split ? 2 : 1), 32,
                             (uint64_t)in.w * 32, (uint64_t)in.h * in.w * 32, p.tw, p.th);
    else if (spatial)
      ta = make_tmap(in.hi, in.c, in.w, in.h, (uint64_t)in.n * (split ? 2 : 1), p.tw, p.th, p.a_scale);
    else
      ta = make_tmap(in.hi, in.c, (uint64_t)p.W, 1, split ? 2 : 1, 128, 1);
    tb = make_tmap(w.w, (uint64_t)w.taps * w.cin, w.cout_pad, 1, split ? 2 : 1, BN, 1);
    // epilogue output path: TMA store for plain NHWC pl16 outputs whose channel count is a multiple of 64
    CUtensorMap tc = tb;
    p.tma_store = (!o.no_bf16_out && o.out_f32 == nullptr && (w.cout_pad % 64) == 0 &&
                   out.c == w.cout_pad && use_tma_store) ? 1 : 0;
    if (p.tma_store) {
      if (spatial) {
        // the transposed-role kernel stores slabs of IGT_SLAB_ROWS patch rows
        tc = make_tmap(out.hi, out.c, out.w, out.h, (uint64_t)out.n * (split ? 2 : 1), p.tw, tform ? IGT_SLAB_ROWS : p.th);
        p.c_lo_img = out.n;
      } else {
        tc = make_tmap(out.hi, out.c, (uint64_t)p.W, 1, split ? 2 : 1, 128, 1);
        p.c_lo_img = 1;
      }
    }
    // residual on the tensor core (see igemm.cuh): same-shape shortcut, channel count a multiple of BLOCK_N
    CUtensorMap tr = tb, ti = tb;
    p.res_kb = 0;
    // (Measured, round 2: adding the shortcut of the 256 -> 1024 / 512 -> 2048 launches in the epilogue registers
    // instead -- 2 of their 6 / 10 k-blocks are identity MMAs -- was 1.3 % SLOWER end to end; the MMA form stays.)
    if (use_res_mma && o.res_mode == RES_SAME && p.tma_store && (w.cout_pad % BN) == 0 && o.res->c == w.cout_pad &&
        o.res->split == split) {
      if (spatial) {
        tr = make_tmap(o.res->hi, o.res->c, o.res->w, o.res->h, (uint64_t)o.res->n * (split ? 2 : 1), p.tw, p.th);
        p.r_lo_img = o.res->n;
      } else {
        tr = make_tmap(o.res->hi, o.res->c, (uint64_t)p.W, 1, split ? 2 : 1, 128, 1);
        p.r_lo_img = 1;
      }
      ti = make_tmap(identity(BN), 64, (uint64_t)(BN / 64) * BN, 1, 1, BN, 1);
      p.res_kb = BN / 64;
      p.res_narrow = res_narrow ? 1 : 0;
      p.res_mode = RES_NONE;  // the epilogue no longer sees a residual
    }
    if (dual) {
      const Act& ax = *o.aux_in;
      if (spatial) {
        tr = make_tmap(ax.hi, ax.c, ax.w, ax.h, (uint64_t)ax.n * (split ? 2 : 1), p.tw, p.th, o.aux_stride);
        p.r_lo_img = ax.n;
      } else {
        tr = make_tmap(ax.hi, ax.c, (uint64_t)p.W, 1, split ? 2 : 1, 128, 1);
        p.r_lo_img = 1;
      }
      ti = make_tmap(o.aux_w->w, (uint64_t)o.aux_w->cin, o.aux_w->cout_pad, 1, split ? 2 : 1, BN, 1);
      p.res_kb = o.aux_w->cin / 64;
      p.res_conv = 1;
      p.r_scale = o.aux_stride;
    }
    const int num_kb = w.taps * (w.cin / 64) + p.res_kb;
    const bool chunked = split && kc > 0 && num_kb > kc && num_kb > chunk_above_kb && BN <= 128;
    p.kc = chunked ? kc : num_kb;
    if (split && BN > 128)
      throw std::runtime_error("conv: the scaled-lo half format needs the cross-term accumulator (BLOCK_N <= 128)");
    // the pair kernel pays a cross-CTA handshake per tile: at BLOCK_N = 128 it wins from 16 k-blocks per tile up
    // (measured, +8..22 % on the 3x3 and K >= 1024 layers) and loses below 8 (the short-K layers are epilogue / HBM
    // bound).  BLOCK_N = 64: the layer1 3x3 convs gain 5 %; the stem (4 k-blocks) loses 12 % and stays on one CTA.
    bool pair = !tform && use_cta2 && split && p.res_kb == 0 && !o.stem_window &&
                ((BN == 128 && num_kb >= cta2_min_kb) ||
                 (BN == 64 && w.taps == 9 && num_kb >= cta2_min_kb64 && !chunked));
    if (profiling) {
      // algorithmic HBM bytes: every operand element once at its stored width (activations 4 B split / 2 B pl16)
      const double eb = split ? 4.0 : 2.0;
      const double pix = (double)p.n_img * p.H * p.W;
      double by = pix * (o.stem_window ? 16.0 : (double)w.cin) * eb * (o.stride == 2 ? 1.0 : 1.0) +
                  (double)w.cout_pad * w.taps * w.cin * eb;
      if (o.out_f32) by += pix * out.c * 4.0;
      if (!o.no_bf16_out) by += pix * out.c * eb;
      if (o.res_mode != RES_NONE) by += (o.res_mode == RES_NEAREST ? 0.25 : 1.0) * pix * w.cout_pad * eb;
      if (dual) by += pix * o.aux_w->cin * eb + (double)w.cout_pad * o.aux_w->cin * eb;
      LayerRec r;
      snprintf(r.sig, sizeof(r.sig), "%dx%dx%d k%d%s cin%d cout%d BN%d%s%s%s%s%s", p.n_img, p.H, p.W,
               o.stem_window ? 7 : (w.taps == 9 ? 3 : 1), o.stride == 2 ? "s2" : "", o.stem_window ? 3 : w.cin, w.cout, BN,
               chunked ? " chunk" : "", tform ? " tform" : (pair ? " pair" : ""),
               dual ? (o.aux_stride == 2 ? " +ds2" : " +ds") : (p.res_kb ? " resmma" : (o.res_mode != RES_NONE ? " res" : "")),
               p.tma_store ? " tma" : " direct", o.relu ? " relu" : "");
      r.flops = fl; r.bytes = by;
      recs.push_back(r);
    }
    if (tform) {
      const int K = w.taps * w.cin;
      if (!w.wt) {
        CALD_CUDA_CHECK(cudaMalloc((void**)&w.wt, (size_t)2 * 128 * K * sizeof(pl16)));
        igemm_t_stack_weights_kernel<<<64, 256, 0, st>>>(w.w, w.wt, K);
        CALD_CUDA_CHECK(cudaGetLastError());
      }
      const CUtensorMap tws = make_tmap(w.wt, (uint64_t)K, 128, 1, 2, 128, 1);
      launch_t(ta, tws, tc, p, st);
    } else if (pair) {
      const CUtensorMap tbh = make_tmap(w.w, (uint64_t)w.taps * w.cin, w.cout_pad, 1, 2, BN / 2, 1);
      if (BN == 64) launch_tc2<64, false>(ta, tb, tbh, tc, p, st);
      else if (chunked) launch_tc2<128, true>(ta, tb, tbh, tc, p, st);
      else launch_tc2<128, false>(ta, tb, tbh, tc, p, st);
    } else if (split) {
      if (BN == 64) { if (chunked) launch_tc<64, true, true>(ta, tb, tc, tr, ti, p, st); else launch_tc<64, true, false>(ta, tb, tc, tr, ti, p, st); }
      else if (BN == 128) { if (chunked) launch_tc<128, true, true>(ta, tb, tc, tr, ti, p, st); else launch_tc<128, true, false>(ta, tb, tc, tr, ti, p, st); }
      else launch_tc<256, true, false>(ta, tb, tc, tr, ti, p, st);
    } else {
      if (BN == 64) launch_tc<64, false, false>(ta, tb, tc, tr, ti, p, st);
      else if (BN == 128) launch_tc<128, false, false>(ta, tb, tc, tr, ti, p, st);
      else launch_tc<256, false, false>(ta, tb, tc, tr, ti, p, st);
    }
  }
};

}  // namespace cald
