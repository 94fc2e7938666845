// Stage-level C ABI (see include/cald_b200_ops.h).  Host buffers in, host buffers out.
#include "layers.cuh"
#include "../../include/cald_b200_ops.h"

using namespace cald;

static thread_local std::string g_ops_err;
extern "C" const char* cald_ops_last_error(void) { return g_ops_err.c_str(); }

extern "C" long long cald_ops_pair_launches(void) { return pair_launch_counter().load(); }
extern "C" long long cald_ops_tform_launches(void) { return tform_launch_counter().load(); }

#define OPS_TRY try {
#define OPS_CATCH                                   \
  }                                                 \
  catch (const std::exception& e) {                 \
    g_ops_err = e.what();                           \
    cudaGetLastError();                             \
    return -1;                                      \
  }                                                 \
  return 0;

namespace cald {
// Upload torch-layout weights [cout][cin][k][k] as [2][cout_pad][(r,s,cin)] split pl16 (+ fp32 bias).
// plan: position of this tensor's k-steps in its launch (RzPlan, conv_host.cuh): the truncating tensor-core accumulate
// is pre-compensated here, in double, before the value is split into its two half planes.
ConvW upload_conv_weight(const float* w, const float* bias, int cout, int cin, int k, bool split,
                         const float* scale /*per-cout or null*/, RzPlan plan) {
  ConvW cw;
  cw.cout = cout;
  cw.cout_pad = (cout + 7) / 8 * 8;
  cw.cin = cin;
  cw.taps = k * k;
  size_t pe = cw.plane_elems();
  std::vector<pl16> h(pe * (split ? 2 : 1));
  for (auto& v : h) v = float_to_pl16(0.f);
  const int own_steps = cw.taps * cin / 16;
  if (plan.chunk_steps < 0) {
    const int num_kb = (plan.steps_total > 0 ? plan.steps_total : own_steps) / 4;
    plan.chunk_steps = (split && ConvEngine::chunked_for(num_kb)) ? ConvEngine::env_int("CALD_KC", 8) * 4 : 0;
  }
  std::vector<float> fac(std::max(1, own_steps), 1.0f);
  if (split)
    for (int j = 0; j < own_steps; ++j) fac[j] = (float)plan.factor(j, own_steps);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int r = 0; r < k; ++r)
        for (int s = 0; s < k; ++s) {
          float v = w[(((size_t)o * cin + c) * k + r) * k + s];
          if (scale) v *= scale[o];
          size_t idx = (size_t)o * cw.taps * cin + (size_t)(r * k + s) * cin + c;
          if (split && own_steps > 0) v = (float)((double)v * (double)fac[((size_t)(r * k + s) * cin + c) / 16]);
          pl16 hi, lo;
          split_pl(v, hi, lo);
          h[idx] = hi;
          if (split) h[pe + idx] = lo;
        }
  CALD_CUDA_CHECK(cudaMalloc((void**)&cw.w, h.size() * sizeof(pl16)));
  CALD_CUDA_CHECK(cudaMemcpy(cw.w, h.data(), h.size() * sizeof(pl16), cudaMemcpyHostToDevice));
  std::vector<float> b(cw.cout_pad, 0.f);
  if (bias) for (int o = 0; o < cout; ++o) b[o] = bias[o];
  CALD_CUDA_CHECK(cudaMalloc((void**)&cw.bias, b.size() * sizeof(float)));
  CALD_CUDA_CHECK(cudaMemcpy(cw.bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  return cw;
}
void free_conv_weight(ConvW& w) {
  if (w.w) cudaFree(w.w);
  if (w.bias) cudaFree(w.bias);
  if (w.wt) cudaFree(w.wt);
  w.w = nullptr; w.bias = nullptr; w.wt = nullptr;
}
}  // namespace cald

extern "C" int cald_op_conv2d(const float* x, int n, int h, int w, int cin, const float* weight, const float* bias,
                              int cout, int k, int stride, int relu, const float* res, int res_mode, int res_h,
                              int res_w, int prec, int impl, int block_n, int kc, float* out) {
  OPS_TRY
  const bool split = (prec == 0);
  cudaStream_t st = 0;
  Arena ar;
  size_t in_e = (size_t)n * h * w * cin;
  int ho = (stride == 2) ? (h + 1) / 2 : h, wo = (stride == 2) ? (w + 1) / 2 : w;
  size_t out_e = (size_t)n * ho * wo * ((cout + 7) / 8 * 8);
  ar.init((in_e + out_e) * 16 + ((size_t)64 << 20) + (size_t)n * res_h * res_w * cout * 16);
  ConvEngine eng;
  int dev;
  CALD_CUDA_CHECK(cudaGetDevice(&dev));
  CALD_CUDA_CHECK(cudaDeviceGetAttribute(&eng.num_sms, cudaDevAttrMultiProcessorCount, dev));
  eng.impl = impl ? CONV_SIMT : CONV_TC;
  eng.split = split;
  eng.force_block_n = block_n;
  if (kc >= 0) eng.kc = kc;
  RzPlan plan;
  if (impl) plan.c = 0.0;   // the SIMT checker accumulates with fp32 FMAs: nothing to compensate
  {
    const int num_kb = k * k * (cin / 64);
    const bool ch = split && eng.kc > 0 && num_kb > eng.kc && num_kb > eng.chunk_above_kb;
    plan.chunk_steps = ch ? eng.kc * 4 : 0;
  }
  ConvW cw = upload_conv_weight(weight, bias, cout, cin, k, split, nullptr, plan);
  float* dx = (float*)ar.alloc(in_e * 4);
  CALD_CUDA_CHECK(cudaMemcpy(dx, x, in_e * 4, cudaMemcpyHostToDevice));
  Act a = alloc_act(ar, n, h, w, cin, split);
  f32_to_split(dx, a, st);
  Act rs;
  if (res_mode) {
    size_t re = (size_t)n * res_h * res_w * cw.cout_pad;
    if (cw.cout_pad != cout) throw std::runtime_error("test op: residual needs cout % 8 == 0");
    float* dr = (float*)ar.alloc(re * 4);
    CALD_CUDA_CHECK(cudaMemcpy(dr, res, re * 4, cudaMemcpyHostToDevice));
    rs = alloc_act(ar, n, res_h, res_w, cout, split);
    f32_to_split(dr, rs, st);
  }
  Act in = a;
  ConvOpts o;
  // both stride-2 forms read the full-resolution tensor through an element-strided A tensor map
  if (stride == 2 && k == 1) o.in_stride2 = true;
  o.relu = relu != 0;
  o.stride = (k == 3) ? stride : 1;
  o.res_mode = res_mode;
  o.res = res_mode ? &rs : nullptr;
  Act y = alloc_act(ar, n, ho, wo, cw.cout_pad, split);
  eng.run(in, cw, y, o, st);
  if (getenv("CALD_OP_TIMING")) {
    // experiment hook (tools/conv_micro.py): re-run the launch a few times bracketed by CUDA events
    eng.profiling = true;
    for (int it = 0; it < 6; ++it) eng.run(in, cw, y, o, st);
    CALD_CUDA_CHECK(cudaStreamSynchronize(st));
    double fl = 0;
    long long nl = 0;
    double ms = eng.drain_profile(&fl, &nl);
    fprintf(stderr, "[conv-timing] %.4f ms/launch  %.1f TFLOP/s algorithmic\n", ms / nl, fl / ms / 1e9);
    eng.profiling = false;
  }
  std::vector<float> hy(y.plane_elems());
  float* dy = (float*)ar.alloc(hy.size() * 4);
  split_to_f32(y, dy, st);
  CALD_CUDA_CHECK(cudaMemcpyAsync(hy.data(), dy, hy.size() * 4, cudaMemcpyDeviceToHost, st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i)
    for (int yy = 0; yy < ho; ++yy)
      for (int xx = 0; xx < wo; ++xx)
        for (int c = 0; c < cout; ++c) {
          const size_t src = (((size_t)i * ho + yy) * wo + xx) * cw.cout_pad + c;
          out[(((size_t)i * ho + yy) * wo + xx) * cout + c] = hy[src];
        }
  free_conv_weight(cw);
  ar.destroy();
  OPS_CATCH
}

extern "C" int cald_op_conv2d_dual(const float* x, int n, int h, int w, int cin, const float* weight, const float* bias,
                                   int cout, int k, const float* x2, int h2, int w2, int cin2, const float* weight2,
                                   const float* bias2, int stride2, int relu, float* out) {
  OPS_TRY
  cudaStream_t st = 0;
  Arena ar;
  const size_t in_e = (size_t)n * h * w * cin, in2_e = (size_t)n * h2 * w2 * cin2;
  const int cout_pad = (cout + 7) / 8 * 8;
  const size_t out_e = (size_t)n * h * w * cout_pad;
  ar.init((in_e + in2_e + out_e) * 16 + ((size_t)64 << 20));
  ConvEngine eng;
  int dev;
  CALD_CUDA_CHECK(cudaGetDevice(&dev));
  CALD_CUDA_CHECK(cudaDeviceGetAttribute(&eng.num_sms, cudaDevAttrMultiProcessorCount, dev));
  RzPlan p1, p2;   // one accumulation: the main contraction's k-steps first, then the second one's
  p1.steps_total = p2.steps_total = k * k * cin / 16 + cin2 / 16;
  p2.steps_before = k * k * cin / 16;
  ConvW cw = upload_conv_weight(weight, bias, cout, cin, k, true, nullptr, p1);
  ConvW cw2 = upload_conv_weight(weight2, bias2, cout, cin2, 1, true, nullptr, p2);
  std::vector<float> bs(cout_pad, 0.f);
  for (int o = 0; o < cout; ++o) bs[o] = (bias ? bias[o] : 0.f) + (bias2 ? bias2[o] : 0.f);
  float* dbs = (float*)ar.alloc(bs.size() * 4);
  CALD_CUDA_CHECK(cudaMemcpy(dbs, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice));
  float* dx = (float*)ar.alloc(in_e * 4);
  CALD_CUDA_CHECK(cudaMemcpy(dx, x, in_e * 4, cudaMemcpyHostToDevice));
  Act a = alloc_act(ar, n, h, w, cin, true);
  f32_to_split(dx, a, st);
  float* dx2 = (float*)ar.alloc(in2_e * 4);
  CALD_CUDA_CHECK(cudaMemcpy(dx2, x2, in2_e * 4, cudaMemcpyHostToDevice));
  Act a2 = alloc_act(ar, n, h2, w2, cin2, true);
  f32_to_split(dx2, a2, st);
  ConvOpts o;
  o.relu = relu != 0;
  o.aux_in = &a2;
  o.aux_w = &cw2;
  o.aux_stride = stride2;
  o.bias_sum = dbs;
  Act y = alloc_act(ar, n, h, w, cw.cout_pad, true);
  eng.run(a, cw, y, o, st);
  std::vector<float> hy(y.plane_elems());
  float* dy = (float*)ar.alloc(hy.size() * 4);
  split_to_f32(y, dy, st);
  CALD_CUDA_CHECK(cudaMemcpyAsync(hy.data(), dy, hy.size() * 4, cudaMemcpyDeviceToHost, st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(st));
  for (size_t pix = 0; pix < (size_t)n * h * w; ++pix)
    for (int c = 0; c < cout; ++c) out[pix * cout + c] = hy[pix * cw.cout_pad + c];
  free_conv_weight(cw);
  free_conv_weight(cw2);
  ar.destroy();
  OPS_CATCH
}


// ---------------------------------------------------------------- measurement probe (tools/overlap_probe.py)
// Can a small kernel on a second stream run WHILE the persistent conv kernels occupy every SM (one CTA per SM with
// ~226 KB of shared memory)?  The pool-ingest decode relies on it.  A 32-block, one-warp kernel that spins for spin_ms
// is timed alone and launched in the middle of a train of conv launches; the conv train is timed with and without it.
__global__ void spin_kernel(long long ns, int smem_probe) {
  extern __shared__ unsigned char probe_smem[];
  if (smem_probe && threadIdx.x == 0) probe_smem[0] = 1;
  if (threadIdx.x != 0) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < ns);
}

extern "C" int cald_op_overlap_probe(int n, int h, int w, int cin, int cout, int k, int iters, int spin_ms,
                                     int spin_smem_bytes, double* out /*[4]*/) {
  OPS_TRY
  cudaStream_t st, st2;
  CALD_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CALD_CUDA_CHECK(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
  Arena ar;
  size_t in_e = (size_t)n * h * w * cin, out_e = (size_t)n * h * w * ((cout + 7) / 8 * 8);
  ar.init((in_e + out_e) * 8 + ((size_t)64 << 20));
  ConvEngine eng;
  int dev;
  CALD_CUDA_CHECK(cudaGetDevice(&dev));
  CALD_CUDA_CHECK(cudaDeviceGetAttribute(&eng.num_sms, cudaDevAttrMultiProcessorCount, dev));
  std::vector<float> wh((size_t)cout * cin * k * k, 0.01f);
  ConvW cw = upload_conv_weight(wh.data(), nullptr, cout, cin, k, true, nullptr, RzPlan());
  Act a = alloc_act(ar, n, h, w, cin, true), y = alloc_act(ar, n, h, w, cw.cout_pad, true);
  CALD_CUDA_CHECK(cudaMemsetAsync(a.hi, 0, a.bytes(), st));
  ConvOpts o;
  o.relu = true;
  cudaEvent_t e0, e1, s0, s1;
  for (cudaEvent_t* ev : {&e0, &e1, &s0, &s1}) CALD_CUDA_CHECK(cudaEventCreate(ev));
  auto train = [&](bool with_spin, double& conv_ms, double& spin_ms_out) {
    CALD_CUDA_CHECK(cudaDeviceSynchronize());
    CALD_CUDA_CHECK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) {
      eng.run(a, cw, y, o, st);
      if (with_spin && i == 1) {
        CALD_CUDA_CHECK(cudaEventRecord(s0, st2));
        spin_kernel<<<32, 32, spin_smem_bytes, st2>>>((long long)spin_ms * 1000000ll, spin_smem_bytes > 0);
        CALD_CUDA_CHECK(cudaEventRecord(s1, st2));
      }
    }
    CALD_CUDA_CHECK(cudaEventRecord(e1, st));
    CALD_CUDA_CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CALD_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    conv_ms = ms;
    if (with_spin) { CALD_CUDA_CHECK(cudaEventElapsedTime(&ms, s0, s1)); spin_ms_out = ms; }
  };
  double d;
  train(false, out[0], d);          // warm-up
  train(false, out[0], d);
  CALD_CUDA_CHECK(cudaEventRecord(s0, st2));
  spin_kernel<<<32, 32, spin_smem_bytes, st2>>>((long long)spin_ms * 1000000ll, spin_smem_bytes > 0);
  CALD_CUDA_CHECK(cudaEventRecord(s1, st2));
  CALD_CUDA_CHECK(cudaDeviceSynchronize());
  float ms = 0;
  CALD_CUDA_CHECK(cudaEventElapsedTime(&ms, s0, s1));
  out[2] = ms;
  train(true, out[1], out[3]);
  free_conv_weight(cw);
  ar.destroy();
  cudaStreamDestroy(st); cudaStreamDestroy(st2);
  OPS_CATCH
}
