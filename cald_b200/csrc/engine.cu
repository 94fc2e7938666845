// Host side of the scoring engine: weight folding, the Faster R-CNN forward pass as a stream of
// hand-written kernels, the CALD scoring loop, and the public C ABI (include/cald_b200.h).
#include <map>
#include <cmath>
#include <algorithm>
#include <memory>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <thread>
#include "layers.cuh"
#include "det.cuh"
#include "pre.cuh"
#include "cons.cuh"
#include "ret.cuh"
#include "jpeg.cuh"
#include "select.cuh"
#include "../../include/cald_b200.h"

using namespace cald;

namespace cald {
ConvW upload_conv_weight(const float* w, const float* bias, int cout, int cin, int k, bool split, const float* scale,
                         RzPlan plan);
void free_conv_weight(ConvW& w);
}

namespace {

struct HostTensor {
  std::vector<float> v;
  std::vector<int64_t> shape;
};

struct Block {
  ConvW c1, c2, c3, ds;
  bool has_ds = false;
  int stride = 1;
  float* bias_c3ds = nullptr;  // device [cout_pad]: folded bn3 bias + folded downsample-bn bias (fused shortcut launch)
};

// fixed-capacity detections of a set of views (device)
struct ViewSet {
  int V = 0;
  DetOut det{};
  float* scores = nullptr;    // [V][cap][C]
  float* prob_max = nullptr;  // [V][cap] (per proposal)
};

static std::string g_create_err;

}  // namespace

struct cald_engine {
  cald_config cfg;
  std::string err;
  cudaStream_t st = nullptr;
  ConvEngine conv;
  Arena arena;
  bool split = true;
  int C = 0;          // classes incl. background
  int cap = 1000;     // proposals per view
  int det_cap = 100;
  long long launches = 0;

  std::map<std::string, HostTensor> staged;
  bool weights_ready = false;
  ConvW stem;
  std::vector<std::vector<Block>> layers;
  ConvW fpn_inner[4], fpn_layer[4], rpn_conv, rpn_out, fc6, fc7, pred;
  int head_ld = 0;
  // RetinaNet (retinanet_cal.py:584-625): FPN on C3..C5 (+P6, P7), two 4-conv towers, 3x3 output convs
  bool retina = false;
  ConvW ret_p6, ret_p7, ret_cls_tower[4], ret_reg_tower[4], ret_cls_out, ret_reg_out;
  int ret_per_class = 300;   // detections_per_img is applied per class (retinanet_cal.py:463)
  int* d_overflow = nullptr; // raised by the RetinaNet post-processing when a fixed capacity is exceeded

  // device constants
  int* d_lut = nullptr;  // [(det_cap+1)][50]
  std::map<std::string, std::vector<float>> dbg;
  std::map<long long, std::pair<int*, int*>> pil_cache;  // (in,out,filter) -> device bounds, kk
  std::map<long long, int> pil_ksize;
  std::vector<float> last_per_view;
  std::vector<int> last_ref_counts;  // detections of every image's reference view in the last scoring call
  int last_A = 0;
  // debug = 1: detections of every view of the last scoring call, ragged, (image, [reference, aug 0, aug 1, ...]) order
  struct DbgViews {
    std::vector<int> counts;
    std::vector<float> boxes, scores, prob_max;
    std::vector<int> labels;
    void clear() { counts.clear(); boxes.clear(); scores.clear(); prob_max.clear(); labels.clear(); }
  } dbg_views;
  cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // CALD_TRACE=1: host timestamp + CUDA event at named points of a scoring call, printed at the end of the call
  struct TracePt { const char* name; double host_us; cudaEvent_t ev; };
  std::vector<TracePt> trace_pts;
  std::vector<cudaEvent_t> trace_pool;
  std::chrono::steady_clock::time_point trace_t0;
  bool tracing = getenv("CALD_TRACE") != nullptr;
  void trace(const char* name) {
    if (!tracing) return;
    if (trace_pts.empty()) trace_t0 = std::chrono::steady_clock::now();
    cudaEvent_t ev;
    if (trace_pts.size() < trace_pool.size()) ev = trace_pool[trace_pts.size()];
    else { cudaEventCreate(&ev); trace_pool.push_back(ev); }
    cudaEventRecord(ev, st);
    trace_pts.push_back({name, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - trace_t0).count(), ev});
  }
  void trace_dump() {
    if (!tracing || trace_pts.empty()) return;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[trace]");
    for (size_t i = 0; i < trace_pts.size(); ++i) {
      float ms = 0;
      if (i) cudaEventElapsedTime(&ms, trace_pts[0].ev, trace_pts[i].ev);
      fprintf(stderr, " %s host=%.1fms gpu=%.1fms |", trace_pts[i].name, trace_pts[i].host_us / 1000.0, ms);
    }
    fprintf(stderr, "\n");
    trace_pts.clear();
  }
  // Image upload pipeline: chunk i+1's u8 images travel host -> device on `copy_st` while chunk i computes on `st`.
  // Pageable caller buffers are first gathered into one of two page-locked staging buffers (one CPU memcpy, then a
  // single link-speed DMA instead of n driver-staged pageable copies).
  cudaStream_t copy_st = nullptr;
  uint8_t* pinned[2] = {nullptr, nullptr};
  size_t pinned_cap[2] = {0, 0};
  cudaEvent_t pinned_free[2] = {nullptr, nullptr};  // the DMA out of staging buffer k has completed
  cudaEvent_t upload_done[2] = {nullptr, nullptr};  // device slab k holds its chunk
  int views_per_pass = 8;      // resolved cfg.max_views_per_pass (0 = sized from the arena, cald_create)

  // device copies of the detector's weights (re-uploaded by every cald_load_weights: the AL cycle retrains the model)
  void free_weights() {
    auto fw = [](ConvW& w) { free_conv_weight(w); };
    fw(stem); fw(rpn_conv); fw(rpn_out); fw(fc6); fw(fc7); fw(pred);
    fw(ret_p6); fw(ret_p7); fw(ret_cls_out); fw(ret_reg_out);
    for (int i = 0; i < 4; ++i) { fw(ret_cls_tower[i]); fw(ret_reg_tower[i]); }
    for (int i = 0; i < 4; ++i) { fw(fpn_inner[i]); fw(fpn_layer[i]); }
    for (auto& l : layers) for (auto& b : l) {
      fw(b.c1); fw(b.c2); fw(b.c3);
      if (b.has_ds) fw(b.ds);
      if (b.bias_c3ds) cudaFree(b.bias_c3ds);
    }
    layers.clear();
    weights_ready = false;
  }
  ~cald_engine() {
    if (d_lut) cudaFree(d_lut);
    for (int k = 0; k < 2; ++k) {
      if (pinned[k]) cudaFreeHost(pinned[k]);
      if (pinned_free[k]) cudaEventDestroy(pinned_free[k]);
      if (upload_done[k]) cudaEventDestroy(upload_done[k]);
    }
    if (copy_st) cudaStreamDestroy(copy_st);
    for (auto& kv : pil_cache) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
    free_weights();
    if (d_overflow) cudaFree(d_overflow);
    arena.destroy();
    if (st) cudaStreamDestroy(st);
  }
};

namespace {

#define KLAUNCH(e) ((e)->launches++)

const HostTensor& need(cald_engine* e, const std::string& name) {
  auto it = e->staged.find(name);
  if (it == e->staged.end()) throw std::runtime_error("missing weight tensor: " + name);
  return it->second;
}

std::string canonical_key(const std::string& n) {
  // torchvision 0.8.2 spellings (README.md:10-11 pin) -> current
  if (n == "rpn.head.conv.weight") return "rpn.head.conv.0.0.weight";
  if (n == "rpn.head.conv.bias") return "rpn.head.conv.0.0.bias";
  for (const char* grp : {"backbone.fpn.inner_blocks.", "backbone.fpn.layer_blocks."}) {
    std::string g(grp);
    if (n.compare(0, g.size(), g) == 0) {
      std::string rest = n.substr(g.size());  // "i.weight" or "i.0.weight"
      size_t dot = rest.find('.');
      if (dot != std::string::npos && (rest.substr(dot + 1) == "weight" || rest.substr(dot + 1) == "bias"))
        return g + rest.substr(0, dot) + ".0." + rest.substr(dot + 1);
    }
  }
  return n;
}

// truncation pre-compensation only applies to the tensor-core kernels (the SIMT checker accumulates with fp32 FMAs)
RzPlan base_plan(cald_engine* e) {
  RzPlan p;
  if (e->cfg.conv_impl == CALD_CONV_SIMT) p.c = 0.0;
  return p;
}

// conv + FrozenBatchNorm2d folded: scale = w_bn * rsqrt(var + eps), bias = b_bn - mean * scale (tv:ops/misc.py:54-63)
ConvW fold_conv_bn(cald_engine* e, const std::string& conv, const std::string& bn, RzPlan plan) {
  const HostTensor& w = need(e, conv + ".weight");
  const HostTensor& g = need(e, bn + ".weight");
  const HostTensor& b = need(e, bn + ".bias");
  const HostTensor& m = need(e, bn + ".running_mean");
  const HostTensor& v = need(e, bn + ".running_var");
  int cout = (int)w.shape[0], cin = (int)w.shape[1], k = (int)w.shape[2];
  std::vector<float> scale(cout), bias(cout);
  for (int o = 0; o < cout; ++o) {
    scale[o] = g.v[o] * (1.0f / sqrtf(v.v[o] + 1e-5f));
    bias[o] = b.v[o] - m.v[o] * scale[o];
  }
  return upload_conv_weight(w.v.data(), bias.data(), cout, cin, k, e->split, scale.data(), plan);
}

ConvW plain_conv(cald_engine* e, const std::string& name) {
  const HostTensor& w = need(e, name + ".weight");
  const HostTensor& b = need(e, name + ".bias");
  int k = w.shape.size() == 4 ? (int)w.shape[2] : 1;
  return upload_conv_weight(w.v.data(), b.v.data(), (int)w.shape[0], (int)w.shape[1], k, e->split, nullptr, base_plan(e));
}

void finalize_weights(cald_engine* e) {
  const int depth = e->cfg.depth;
  static const int blocks50[4] = {3, 4, 6, 3}, blocks101[4] = {3, 4, 23, 3};
  const int* nb = depth == 101 ? blocks101 : blocks50;
  // ---- stem: [64][3][7][7] with BN folded -> 4x4 conv over the space-to-depth image:
  //      k = dy*64 + dx*16 + (py*2+px)*3 + c  with  r = 2*dy + py - 1,  s = 2*dx + px - 1  (zero where r, s fall outside 0..6)
  {
    const HostTensor& w = need(e, "backbone.body.conv1.weight");
    const HostTensor& g = need(e, "backbone.body.bn1.weight");
    const HostTensor& b = need(e, "backbone.body.bn1.bias");
    const HostTensor& m = need(e, "backbone.body.bn1.running_mean");
    const HostTensor& v = need(e, "backbone.body.bn1.running_var");
    std::vector<float> w2((size_t)64 * 256, 0.f), bias(64);
    for (int o = 0; o < 64; ++o) {
      float scale = g.v[o] * (1.0f / sqrtf(v.v[o] + 1e-5f));
      bias[o] = b.v[o] - m.v[o] * scale;
      for (int dy = 0; dy < 4; ++dy)
        for (int dx = 0; dx < 4; ++dx)
          for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
              int r = 2 * dy + py - 1, s2 = 2 * dx + px - 1;
              if (r < 0 || r > 6 || s2 < 0 || s2 > 6) continue;
              for (int c = 0; c < 3; ++c)
                w2[(size_t)o * 256 + dy * 64 + dx * 16 + (py * 2 + px) * 3 + c] =
                    w.v[(((size_t)o * 3 + c) * 7 + r) * 7 + s2] * scale;
            }
    }
    // upload as a "4-tap" conv with Cin = 64: layout [o][tap][64] == [o][256]
    ConvW cw = upload_conv_weight(w2.data(), bias.data(), 64, 256, 1, e->split, nullptr, base_plan(e));
    cw.cin = 64;
    cw.taps = 4;
    e->stem = cw;
  }
  e->layers.assign(4, {});
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < nb[li]; ++bi) {
      Block blk;
      std::string pre = "backbone.body.layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      blk.c1 = fold_conv_bn(e, pre + ".conv1", pre + ".bn1", base_plan(e));
      blk.c2 = fold_conv_bn(e, pre + ".conv2", pre + ".bn2", base_plan(e));
      blk.stride = (bi == 0 && li > 0) ? 2 : 1;
      // conv3's accumulation continues after its own k-steps: with the fused projection shortcut (first block of
      // layer1-3, run_body) by the shortcut's k-steps, otherwise by the identity-routed residual (one more accumulate)
      const HostTensor& w3 = need(e, pre + ".conv3.weight");
      const int c3_steps = (int)w3.shape[1] / 16;
      const bool fused_ds = bi == 0 && li < 3 && ConvEngine::env_flag("CALD_FUSE_DS", true) &&
                            e->cfg.conv_impl != CALD_CONV_SIMT;
      RzPlan p3 = base_plan(e), pd = base_plan(e);
      if (bi == 0) {
        const HostTensor& wd = need(e, pre + ".downsample.0.weight");
        const int ds_steps = (int)wd.shape[1] / 16;
        if (fused_ds) { p3.steps_total = pd.steps_total = c3_steps + ds_steps; pd.steps_before = c3_steps; }
      } else {
        p3.steps_total = c3_steps + 1;
      }
      blk.c3 = fold_conv_bn(e, pre + ".conv3", pre + ".bn3", p3);
      if (bi == 0) {
        blk.has_ds = true;
        blk.ds = fold_conv_bn(e, pre + ".downsample.0", pre + ".downsample.1", pd);
        // out = relu(bn3(conv3(y)) + bn_d(conv_d(x))): with both BNs folded the two biases simply add
        std::vector<float> b3(blk.c3.cout_pad), bd(blk.ds.cout_pad);
        if (b3.size() != bd.size()) throw std::runtime_error("downsample / conv3 channel mismatch");
        CALD_CUDA_CHECK(cudaMemcpy(b3.data(), blk.c3.bias, b3.size() * 4, cudaMemcpyDeviceToHost));
        CALD_CUDA_CHECK(cudaMemcpy(bd.data(), blk.ds.bias, bd.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < b3.size(); ++i) b3[i] += bd[i];
        CALD_CUDA_CHECK(cudaMalloc((void**)&blk.bias_c3ds, b3.size() * 4));
        CALD_CUDA_CHECK(cudaMemcpy(blk.bias_c3ds, b3.data(), b3.size() * 4, cudaMemcpyHostToDevice));
      }
      e->layers[li].push_back(blk);
    }
  }
  if (e->retina) {
    for (int i = 0; i < 3; ++i) {
      e->fpn_inner[i] = plain_conv(e, "backbone.fpn.inner_blocks." + std::to_string(i) + ".0");
      e->fpn_layer[i] = plain_conv(e, "backbone.fpn.layer_blocks." + std::to_string(i) + ".0");
    }
    e->ret_p6 = plain_conv(e, "backbone.fpn.extra_blocks.p6");
    e->ret_p7 = plain_conv(e, "backbone.fpn.extra_blocks.p7");
    for (int i = 0; i < 4; ++i) {
      e->ret_cls_tower[i] = plain_conv(e, "head.classification_head.conv." + std::to_string(2 * i));
      e->ret_reg_tower[i] = plain_conv(e, "head.regression_head.conv." + std::to_string(2 * i));
    }
    e->ret_cls_out = plain_conv(e, "head.classification_head.cls_logits");
    e->ret_reg_out = plain_conv(e, "head.regression_head.bbox_reg");
    if (e->ret_cls_out.cout != RET_A * e->C) throw std::runtime_error("cls_logits: num_classes mismatch");
    if (e->ret_reg_out.cout != RET_A * 4) throw std::runtime_error("bbox_reg: expected 36 output channels");
    e->staged.clear();
    e->weights_ready = true;
    return;
  }
  for (int i = 0; i < 4; ++i) {
    e->fpn_inner[i] = plain_conv(e, "backbone.fpn.inner_blocks." + std::to_string(i) + ".0");
    e->fpn_layer[i] = plain_conv(e, "backbone.fpn.layer_blocks." + std::to_string(i) + ".0");
  }
  e->rpn_conv = plain_conv(e, "rpn.head.conv.0.0");
  {
    // 1x1 -> 3 objectness + 1x1 -> 12 deltas fused into one [16][256] matrix (rows 0..2 logits, 3..14 deltas)
    const HostTensor& wc = need(e, "rpn.head.cls_logits.weight");
    const HostTensor& bc = need(e, "rpn.head.cls_logits.bias");
    const HostTensor& wb = need(e, "rpn.head.bbox_pred.weight");
    const HostTensor& bb = need(e, "rpn.head.bbox_pred.bias");
    std::vector<float> w(15 * 256), b(15);
    memcpy(w.data(), wc.v.data(), 3 * 256 * 4);
    memcpy(w.data() + 3 * 256, wb.v.data(), 12 * 256 * 4);
    for (int i = 0; i < 3; ++i) b[i] = bc.v[i];
    for (int i = 0; i < 12; ++i) b[3 + i] = bb.v[i];
    e->rpn_out = upload_conv_weight(w.data(), b.data(), 15, 256, 1, e->split, nullptr, base_plan(e));
  }
  {
    // fc6: torch flattens [256][7][7] as c*49 + ph*7 + pw; RoIAlign writes (ph*7+pw)*256 + c
    const HostTensor& w = need(e, "roi_heads.box_head.fc6.weight");
    const HostTensor& b = need(e, "roi_heads.box_head.fc6.bias");
    int out = (int)w.shape[0];
    std::vector<float> w2(w.v.size());
    for (int o = 0; o < out; ++o)
      for (int c = 0; c < 256; ++c)
        for (int s = 0; s < 49; ++s) w2[(size_t)o * 12544 + s * 256 + c] = w.v[(size_t)o * 12544 + c * 49 + s];
    e->fc6 = upload_conv_weight(w2.data(), b.v.data(), out, 12544, 1, e->split, nullptr, base_plan(e));
  }
  e->fc7 = plain_conv(e, "roi_heads.box_head.fc7");
  {
    const HostTensor& wc = need(e, "roi_heads.box_predictor.cls_score.weight");
    const HostTensor& bc = need(e, "roi_heads.box_predictor.cls_score.bias");
    const HostTensor& wb = need(e, "roi_heads.box_predictor.bbox_pred.weight");
    const HostTensor& bb = need(e, "roi_heads.box_predictor.bbox_pred.bias");
    const int C = e->C;
    if ((int)wc.shape[0] != C || (int)wb.shape[0] != 4 * C) throw std::runtime_error("box predictor: num_classes mismatch");
    std::vector<float> w((size_t)5 * C * 1024), b(5 * C);
    memcpy(w.data(), wc.v.data(), (size_t)C * 1024 * 4);
    memcpy(w.data() + (size_t)C * 1024, wb.v.data(), (size_t)4 * C * 1024 * 4);
    for (int i = 0; i < C; ++i) b[i] = bc.v[i];
    for (int i = 0; i < 4 * C; ++i) b[C + i] = bb.v[i];
    e->pred = upload_conv_weight(w.data(), b.data(), 5 * C, 1024, 1, e->split, nullptr, base_plan(e));
    e->head_ld = e->pred.cout_pad;
  }
  e->staged.clear();
  e->weights_ready = true;
}

// torchvision cell anchors (tv:anchor_utils.py:58-75): fp32 sqrt, round-half-even
void cell_anchors(float size, float out[3][4]) {
  const float ratios[3] = {0.5f, 1.0f, 2.0f};
  for (int i = 0; i < 3; ++i) {
    float hr = sqrtf(ratios[i]);
    float wr = 1.0f / hr;
    float ws = wr * size, hs = hr * size;
    out[i][0] = nearbyintf(-ws / 2.f);
    out[i][1] = nearbyintf(-hs / 2.f);
    out[i][2] = nearbyintf(ws / 2.f);
    out[i][3] = nearbyintf(hs / 2.f);
  }
}

void dbg_store_act(cald_engine* e, const char* name, const Act& a) {
  if (!e->cfg.debug) return;
  std::vector<float>& v = e->dbg[name];
  v.resize(a.plane_elems());
  float* tmp = (float*)e->arena.alloc(v.size() * 4);
  split_to_f32(a, tmp, e->st);
  CALD_CUDA_CHECK(cudaMemcpyAsync(v.data(), tmp, v.size() * 4, cudaMemcpyDeviceToHost, e->st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
  e->arena.free(tmp);
}
void dbg_store_f32(cald_engine* e, const char* name, const float* d, size_t n) {
  if (!e->cfg.debug) return;
  std::vector<float>& v = e->dbg[name];
  v.resize(n);
  CALD_CUDA_CHECK(cudaMemcpyAsync(v.data(), d, n * 4, cudaMemcpyDeviceToHost, e->st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
}

Act conv(cald_engine* e, const Act& in, const ConvW& w, int n, int h, int wd, const ConvOpts& o) {
  Act out = alloc_act(e->arena, n, h, wd, w.cout_pad, e->split);
  e->conv.run(in, w, out, o, e->st);
  KLAUNCH(e);
  return out;
}

// ------------------------------------------------------------------------------------------------------------
// One forward pass over V views that share the padded input size (Hp, Wp).
// d_views / d_cuts / d_image_hw / d_ratio live on the device.  Results go to `vs` (local view order).
// ------------------------------------------------------------------------------------------------------------
// transform + ResNet body: fills cfeat[0..3] = C2..C5 (tv:models/resnet.py, FrozenBN folded)
void run_body(cald_engine* e, int V, int Hp, int Wp, const ViewDesc* d_views, const CutRects* d_cuts, Act* cfeat) {
  cudaStream_t st = e->st;
  Arena& ar = e->arena;
  const bool split = e->split;

  // ---- transform (normalise + resize + pad) fused with the stem's space-to-depth layout
  const int Ho = Hp / 2, Wo = Wp / 2;
  Act sin = alloc_act(ar, V, Ho + 3, Wo + 3, 16, split);
  {
    dim3 grid((Wo + 3 + 127) / 128, Ho + 3, V);
    view_stem_input_kernel<<<grid, 128, 0, st>>>(d_views, d_cuts, Ho + 3, Wo + 3, sin.hi, sin.lo());
    CALD_CUDA_CHECK(cudaGetLastError());
    KLAUNCH(e);
  }
  if (e->cfg.debug) {
    float* img = (float*)ar.alloc((size_t)V * Hp * Wp * 3 * 4);
    dim3 grid((Wp * 3 + 191) / 192, Hp, V);
    view_preprocess_kernel<<<grid, 192, 0, st>>>(d_views, d_cuts, Hp, Wp, img);
    dbg_store_f32(e, "input", img, (size_t)V * Hp * Wp * 3);
    ar.free(img);
  }
  // ---- stem: 7x7/2 conv (+BN+ReLU) as a windowed 4x4 implicit GEMM, then maxpool
  ConvOpts relu_o;
  relu_o.relu = true;
  Act x;
  {
    ConvOpts so;
    so.relu = true;
    so.stem_window = true;
    x = conv(e, sin, e->stem, V, Ho, Wo, so);
  }
  free_act(ar, sin);
  Act xp = maxpool3x3s2(ar, x, st);
  KLAUNCH(e);
  free_act(ar, x);
  x = xp;
  // ---- residual stages
  for (int li = 0; li < 4; ++li) {
    for (size_t bi = 0; bi < e->layers[li].size(); ++bi) {
      const Block& b = e->layers[li][bi];
      const int ho = b.stride == 2 ? (x.h + 1) / 2 : x.h, wo = b.stride == 2 ? (x.w + 1) / 2 : x.w;
      // The projection shortcut of a stage's first block is accumulated INSIDE the conv3 launch (second contraction
      // over x, conv_host.cuh ConvOpts::aux_*): its 4x-wide output is never written or re-read.  layer4's two
      // contractions (K = 512 + 1024) are tensor-bound and stay separate launches on the CTA-pair kernel.
      static const bool fuse_env = ConvEngine::env_flag("CALD_FUSE_DS", true);
      const bool fuse_ds = b.has_ds && fuse_env && e->conv.impl == CONV_TC && li < 3;
      Act idt;
      if (b.has_ds && !fuse_ds) {
        ConvOpts o;
        if (b.stride == 2) {
          o.in_stride2 = true;  // the A tensor map skips every other pixel: no subsampled copy
          idt = conv(e, x, b.ds, V, ho, wo, o);
        } else {
          idt = conv(e, x, b.ds, V, ho, wo, o);
        }
      }
      // conv1 at full resolution; a stride-2 conv2 reads it through an element-strided TMA map (no phase-split copy)
      Act t1 = conv(e, x, b.c1, V, x.h, x.w, relu_o);
      ConvOpts o2;
      o2.relu = true;
      o2.stride = b.stride;
      Act t2 = conv(e, t1, b.c2, V, ho, wo, o2);
      free_act(ar, t1);
      ConvOpts o3;
      o3.relu = true;
      if (fuse_ds) {
        o3.aux_in = &x;
        o3.aux_w = &b.ds;
        o3.aux_stride = b.stride;
        o3.bias_sum = b.bias_c3ds;
      } else {
        o3.res_mode = RES_SAME;
        o3.res = b.has_ds ? &idt : &x;
      }
      Act y = conv(e, t2, b.c3, V, ho, wo, o3);
      free_act(ar, t2);
      if (b.has_ds && !fuse_ds) free_act(ar, idt);
      const bool keep_x = (bi == 0 && li > 0);  // x is the previous stage's output (C2..C4), needed by the FPN
      if (!keep_x) free_act(ar, x);
      x = y;
    }
    cfeat[li] = x;
  }
  if (e->cfg.debug) {
    dbg_store_act(e, "c2", cfeat[0]); dbg_store_act(e, "c3", cfeat[1]);
    dbg_store_act(e, "c4", cfeat[2]); dbg_store_act(e, "c5", cfeat[3]);
  }
}

void forward_pass_retina(cald_engine* e, int V, int Hp, int Wp, const ViewDesc* d_views, const CutRects* d_cuts,
                         const int* d_image_hw, const float* d_ratio, ViewSet& vs);

void forward_pass(cald_engine* e, int V, int Hp, int Wp, const ViewDesc* d_views, const CutRects* d_cuts,
                  const int* d_image_hw, const float* d_ratio, ViewSet& vs) {
  if (e->retina) {
    forward_pass_retina(e, V, Hp, Wp, d_views, d_cuts, d_image_hw, d_ratio, vs);
    return;
  }
  cudaStream_t st = e->st;
  Arena& ar = e->arena;
  const bool split = e->split;
  const int C = e->C, cap = e->cap;
  ConvOpts relu_o;
  relu_o.relu = true;
  Act cfeat[4];
  run_body(e, V, Hp, Wp, d_views, d_cuts, cfeat);
  // ---- FPN (tv:ops/feature_pyramid_network.py:172-221)
  Act pf[5];
  {
    ConvOpts o;
    Act last = conv(e, cfeat[3], e->fpn_inner[3], V, cfeat[3].h, cfeat[3].w, o);
    pf[3] = conv(e, last, e->fpn_layer[3], V, last.h, last.w, o);
    for (int i = 2; i >= 0; --i) {
      ConvOpts oi;
      oi.res_mode = RES_NEAREST;
      oi.res = &last;
      Act inner = conv(e, cfeat[i], e->fpn_inner[i], V, cfeat[i].h, cfeat[i].w, oi);
      free_act(ar, last);
      last = inner;
      pf[i] = conv(e, last, e->fpn_layer[i], V, last.h, last.w, o);
    }
    free_act(ar, last);
    for (int i = 0; i < 4; ++i) free_act(ar, cfeat[i]);
    pf[4] = subsample2(ar, pf[3], st);  // LastLevelMaxPool: kernel 1, stride 2
    e->launches += split ? 2 : 1;
  }
  if (e->cfg.debug) {
    const char* nm[5] = {"p2", "p3", "p4", "p5", "p6"};
    for (int i = 0; i < 5; ++i) dbg_store_act(e, nm[i], pf[i]);
  }
  // ---- RPN head (tv:rpn.py:71-78): 3x3+ReLU then fused 1x1 -> 15 (+1 pad) fp32
  RpnLevels L;
  memset(&L, 0, sizeof(L));
  float* rpn_raw[5];
  static const float sizes[5] = {32.f, 64.f, 128.f, 256.f, 512.f};
  int off = 0;
  for (int l = 0; l < 5; ++l) {
    rpn_raw[l] = (float*)ar.alloc((size_t)V * pf[l].h * pf[l].w * 16 * 4);
    {
      Act t = conv(e, pf[l], e->rpn_conv, V, pf[l].h, pf[l].w, relu_o);
      Act dummy;
      dummy.n = V; dummy.h = pf[l].h; dummy.w = pf[l].w; dummy.c = 16; dummy.split = split; dummy.hi = nullptr;
      ConvOpts o;
      o.out_f32 = rpn_raw[l];
      o.no_bf16_out = true;
      e->conv.run(t, e->rpn_out, dummy, o, st);
      KLAUNCH(e);
      free_act(ar, t);
    }
    RpnLevel& lv = L.lv[l];
    lv.out = rpn_raw[l];
    lv.h = pf[l].h; lv.w = pf[l].w;
    lv.stride_h = Hp / pf[l].h; lv.stride_w = Wp / pf[l].w;
    cell_anchors(sizes[l], lv.base);
    lv.n = pf[l].h * pf[l].w * 3;
    lv.off = off;
    off += lv.n;
    if (e->cfg.debug) {
      std::string nm = "rpn" + std::to_string(l);
      dbg_store_f32(e, nm.c_str(), rpn_raw[l], (size_t)V * pf[l].h * pf[l].w * 16);
    }
  }
  L.total = off;
  free_act(ar, pf[4]);
  // ---- proposals (tv:rpn.py:242-297)
  const int G = V * RPN_LEVELS;
  unsigned long long* keys = (unsigned long long*)ar.alloc((size_t)V * L.total * 8);
  unsigned long long* sel = (unsigned long long*)ar.alloc((size_t)G * TOPK_MAX * 8);
  int* sel_count = (int*)ar.alloc((size_t)G * 4);
  float4* lv_boxes = (float4*)ar.alloc((size_t)G * TOPK_MAX * 16);
  float* lv_scores = (float*)ar.alloc((size_t)G * TOPK_MAX * 4);
  int* lv_count = (int*)ar.alloc((size_t)G * 4);
  int* keep_idx = (int*)ar.alloc((size_t)G * TOPK_MAX * 4);
  int* keep_count = (int*)ar.alloc((size_t)G * 4);
  float4* props = (float4*)ar.alloc((size_t)V * cap * 16);
  float* prop_scores = (float*)ar.alloc((size_t)V * cap * 4);
  int* prop_count = (int*)ar.alloc((size_t)V * 4);
  {
    long long tot = (long long)V * L.total;
    rpn_keys_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(L, V, keys);
    TopkGroups tg;
    memset(&tg, 0, sizeof(tg));
    tg.keys = keys;
    tg.stride_outer = L.total;
    tg.inner = RPN_LEVELS;
    for (int l = 0; l < 5; ++l) { tg.inner_off[l] = L.lv[l].off; tg.inner_n[l] = L.lv[l].n; }
    tg.k = std::min(e->cfg.rpn_pre_nms_top_n, TOPK_MAX);
    topk_select_kernel<<<G, 1024, 0, st>>>(tg, sel, sel_count);
    rpn_decode_kernel<<<G, 1024, 0, st>>>(L, sel, sel_count, d_image_hw, 1e-3f, lv_boxes, lv_scores, lv_count);
    nms_groups_kernel<<<G, 1024, NMS_SMEM, st>>>(lv_boxes, lv_count, (double)e->cfg.rpn_nms_thresh, keep_idx,
                                                 keep_count);
    rpn_merge_kernel<<<V, 1024, MERGE_CAP * 8, st>>>(lv_boxes, lv_scores, keep_idx, keep_count,
                                                     std::min(e->cfg.rpn_post_nms_top_n, cap), props, prop_scores,
                                                     prop_count, cap);
    CALD_CUDA_CHECK(cudaGetLastError());
    e->launches += 5;
  }
  if (e->cfg.debug) {
    dbg_store_f32(e, "proposals", (const float*)props, (size_t)V * cap * 4);
    std::vector<int> pc(V);
    CALD_CUDA_CHECK(cudaMemcpyAsync(pc.data(), prop_count, V * 4, cudaMemcpyDeviceToHost, st));
    CALD_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<float>& v = e->dbg["proposal_count"];
    v.assign(pc.begin(), pc.end());
  }
  for (int l = 0; l < 5; ++l) ar.free(rpn_raw[l]);
  ar.free(keys); ar.free(sel); ar.free(sel_count); ar.free(lv_boxes); ar.free(lv_scores); ar.free(lv_count);
  ar.free(keep_idx); ar.free(keep_count);
  // ---- RoIAlign on P2..P5 -> [V*cap][49*256]
  Act roi = alloc_act(ar, 1, 1, V * cap, 49 * 256, split);
  {
    RoiFeats F;
    for (int l = 0; l < 4; ++l) {
      F.hi[l] = pf[l].hi; F.lo[l] = pf[l].lo();
      F.h[l] = pf[l].h; F.w[l] = pf[l].w;
      // tv:ops/poolers.py _infer_scale: 2 ** round(log2(feat / padded_input))
      F.scale[l] = (float)std::pow(2.0, std::round(std::log2((double)pf[l].h / (double)Hp)));
    }
    F.C = 256;
    static const bool roi_fma = ConvEngine::env_flag("CALD_ROI_FMA", true);
    if (roi_fma) roialign_kernel<true><<<dim3(cap, V), ROI_THREADS, 0, st>>>(F, props, prop_count, cap, roi.hi, roi.lo());
    else roialign_kernel<false><<<dim3(cap, V), ROI_THREADS, 0, st>>>(F, props, prop_count, cap, roi.hi, roi.lo());
    CALD_CUDA_CHECK(cudaGetLastError());
    KLAUNCH(e);
  }
  dbg_store_act(e, "pooled", roi);
  for (int l = 0; l < 4; ++l) free_act(ar, pf[l]);
  // ---- box head (tv:faster_rcnn.py:286-307, 347-372)
  Act f6 = conv(e, roi, e->fc6, 1, 1, V * cap, relu_o);
  free_act(ar, roi);
  Act f7 = conv(e, f6, e->fc7, 1, 1, V * cap, relu_o);
  free_act(ar, f6);
  float* head = (float*)ar.alloc((size_t)V * cap * e->head_ld * 4);
  {
    Act dummy;
    dummy.n = 1; dummy.h = 1; dummy.w = V * cap; dummy.c = e->head_ld; dummy.split = split; dummy.hi = nullptr;
    ConvOpts o;
    o.out_f32 = head;
    o.no_bf16_out = true;
    e->conv.run(f7, e->pred, dummy, o, st);
    KLAUNCH(e);
  }
  free_act(ar, f7);
  dbg_store_f32(e, "head", head, (size_t)V * cap * e->head_ld);
  // ---- postprocess_detections (frcnn_la.py:32-87) + transform.postprocess (292-315)
  {
    long long rows = (long long)V * cap;
    softmax_rows_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(head, e->head_ld, C, rows, vs.scores,
                                                                             vs.prob_max);
    const int kept_cap = (C - 1) * cap;
    unsigned long long* kept = (unsigned long long*)ar.alloc((size_t)V * kept_cap * 8);
    int* kept_count = (int*)ar.alloc((size_t)V * 4);
    unsigned long long* top = (unsigned long long*)ar.alloc((size_t)V * TOPK_MAX * 8);
    int* top_count = (int*)ar.alloc((size_t)V * 4);
    CALD_CUDA_CHECK(cudaMemsetAsync(kept_count, 0, V * 4, st));
    det_class_nms_kernel<<<dim3(C - 1, V), 1024, NMS_SMEM + TOPK_MAX * 8 + TOPK_MAX * 4, st>>>(
        head, e->head_ld, C, vs.scores, props, prop_count, cap, d_image_hw, e->cfg.box_score_thresh,
        (double)e->cfg.box_nms_thresh, kept, kept_count, kept_cap);
    TopkGroups tg;
    memset(&tg, 0, sizeof(tg));
    tg.keys = kept;
    tg.stride_outer = kept_cap;
    tg.inner = 1;
    tg.dyn_n = kept_count;
    tg.k = e->det_cap;
    topk_select_kernel<<<V, 1024, 0, st>>>(tg, top, top_count);
    det_gather_kernel<<<V, 128, 0, st>>>(top, top_count, head, e->head_ld, C, vs.scores, vs.prob_max, props, cap,
                                         d_image_hw, d_ratio, e->det_cap, vs.det);
    CALD_CUDA_CHECK(cudaGetLastError());
    e->launches += 4;
    ar.free(kept); ar.free(kept_count); ar.free(top); ar.free(top_count);
  }
  ar.free(head); ar.free(props); ar.free(prop_scores); ar.free(prop_count);
}

// ------------------------------------------------------------------------------------------------------------
// RetinaNet forward (detection/retinanet_cal.py:492-575): body -> FPN P3..P5 + P6/P7 -> two towers + output convs on
// every level -> per-class post-processing.  vs.scores receives one sigmoid row per DETECTION ([V][det_cap][K]).
// ------------------------------------------------------------------------------------------------------------
void ret_cell_anchors(int level, float out[RET_A][4]) {
  // retinanet_cal.py:347-348 sizes; tv:anchor_utils.py:58-75 (ratio-major, scale-minor; fp32; round-half-even)
  static const int bases[5] = {32, 64, 128, 256, 512};
  const int x = bases[level];
  const float scales[3] = {(float)x, (float)(int)(x * std::pow(2.0, 1.0 / 3)), (float)(int)(x * std::pow(2.0, 2.0 / 3))};
  const float ratios[3] = {0.5f, 1.0f, 2.0f};
  for (int r = 0; r < 3; ++r) {
    const float hr = sqrtf(ratios[r]);
    const float wr = 1.0f / hr;
    for (int sidx = 0; sidx < 3; ++sidx) {
      const float ws = wr * scales[sidx], hs = hr * scales[sidx];
      float* o = out[r * 3 + sidx];
      o[0] = nearbyintf(-ws / 2.f); o[1] = nearbyintf(-hs / 2.f);
      o[2] = nearbyintf(ws / 2.f);  o[3] = nearbyintf(hs / 2.f);
    }
  }
}

Act relu_act(cald_engine* e, const Act& in) {
  Act o = alloc_act(e->arena, in.n, in.h, in.w, in.c, in.split);
  long long n = (long long)in.plane_elems();
  relu_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->st>>>(in.hi, in.lo(), o.hi, o.lo(), n);
  CALD_CUDA_CHECK(cudaGetLastError());
  KLAUNCH(e);
  return o;
}

void forward_pass_retina(cald_engine* e, int V, int Hp, int Wp, const ViewDesc* d_views, const CutRects* d_cuts,
                         const int* d_image_hw, const float* d_ratio, ViewSet& vs) {
  cudaStream_t st = e->st;
  Arena& ar = e->arena;
  const bool split = e->split;
  const int K = e->C;
  ConvOpts relu_o;
  relu_o.relu = true;
  Act cfeat[4];
  run_body(e, V, Hp, Wp, d_views, d_cuts, cfeat);
  free_act(ar, cfeat[0]);  // C2 is not a returned layer (returned_layers=[2,3,4], retinanet_cal.py:618)
  // ---- FPN on C3..C5, then P6 = conv3x3/2(P5), P7 = conv3x3/2(relu(P6))
  Act pf[5];
  {
    ConvOpts o;
    Act last = conv(e, cfeat[3], e->fpn_inner[2], V, cfeat[3].h, cfeat[3].w, o);
    pf[2] = conv(e, last, e->fpn_layer[2], V, last.h, last.w, o);
    for (int i = 1; i >= 0; --i) {
      ConvOpts oi;
      oi.res_mode = RES_NEAREST;
      oi.res = &last;
      Act inner = conv(e, cfeat[i + 1], e->fpn_inner[i], V, cfeat[i + 1].h, cfeat[i + 1].w, oi);
      free_act(ar, last);
      last = inner;
      pf[i] = conv(e, last, e->fpn_layer[i], V, last.h, last.w, o);
    }
    free_act(ar, last);
    for (int i = 1; i < 4; ++i) free_act(ar, cfeat[i]);
    ConvOpts s2;
    s2.stride = 2;
    pf[3] = conv(e, pf[2], e->ret_p6, V, (pf[2].h + 1) / 2, (pf[2].w + 1) / 2, s2);
    Act r6 = relu_act(e, pf[3]);
    pf[4] = conv(e, r6, e->ret_p7, V, (pf[3].h + 1) / 2, (pf[3].w + 1) / 2, s2);
    free_act(ar, r6);
  }
  if (e->cfg.debug) {
    const char* nm[5] = {"p3", "p4", "p5", "p6", "p7"};
    for (int i = 0; i < 5; ++i) dbg_store_act(e, nm[i], pf[i]);
  }
  // ---- heads (retinanet_cal.py:135-151, 225-241): fp32 NHWC outputs, channel a*K + k / a*4 + j
  RetLevels L;
  memset(&L, 0, sizeof(L));
  L.K = K;
  float* cls_raw[5];
  float* reg_raw[5];
  int off = 0;
  for (int l = 0; l < 5; ++l) {
    const int h = pf[l].h, w = pf[l].w;
    for (int tower = 0; tower < 2; ++tower) {
      const ConvW* tw = tower == 0 ? e->ret_cls_tower : e->ret_reg_tower;
      const ConvW& ow = tower == 0 ? e->ret_cls_out : e->ret_reg_out;
      Act t = conv(e, pf[l], tw[0], V, h, w, relu_o);
      for (int i = 1; i < 4; ++i) {
        Act t2 = conv(e, t, tw[i], V, h, w, relu_o);
        free_act(ar, t);
        t = t2;
      }
      float* raw = (float*)ar.alloc((size_t)V * h * w * ow.cout_pad * 4);
      Act dummy;
      dummy.n = V; dummy.h = h; dummy.w = w; dummy.c = ow.cout_pad; dummy.split = split; dummy.hi = nullptr;
      ConvOpts o;
      o.out_f32 = raw;
      o.no_bf16_out = true;
      e->conv.run(t, ow, dummy, o, st);
      KLAUNCH(e);
      free_act(ar, t);
      (tower == 0 ? cls_raw : reg_raw)[l] = raw;
    }
    RetLevel& lv = L.lv[l];
    lv.cls = cls_raw[l]; lv.reg = reg_raw[l];
    lv.h = h; lv.w = w;
    lv.ld_cls = e->ret_cls_out.cout_pad; lv.ld_reg = e->ret_reg_out.cout_pad;
    lv.stride_h = Hp / h; lv.stride_w = Wp / w;
    ret_cell_anchors(l, lv.base);
    lv.n = h * w * RET_A;
    lv.off = off;
    off += lv.n;
    if (e->cfg.debug) {
      std::string nm = "cls" + std::to_string(l);
      dbg_store_f32(e, nm.c_str(), cls_raw[l], (size_t)V * h * w * lv.ld_cls);
      nm = "reg" + std::to_string(l);
      dbg_store_f32(e, nm.c_str(), reg_raw[l], (size_t)V * h * w * lv.ld_reg);
    }
  }
  L.total = off;
  for (int l = 0; l < 5; ++l) free_act(ar, pf[l]);
  // ---- postprocess_detections (retinanet_cal.py:402-490) + transform.postprocess
  {
    const int pc = e->ret_per_class;
    unsigned long long* keys = (unsigned long long*)ar.alloc((size_t)V * K * RET_CAND * 8);
    int* counts = (int*)ar.alloc((size_t)V * K * 4);
    RetKept kept;
    kept.boxes = (float4*)ar.alloc((size_t)V * K * pc * 16);
    kept.scores = (float*)ar.alloc((size_t)V * K * pc * 4);
    kept.anchor = (int*)ar.alloc((size_t)V * K * pc * 4);
    kept.count = (int*)ar.alloc((size_t)V * K * 4);
    CALD_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)V * K * 4, st));
    ret_candidates_kernel<<<dim3(e->conv.num_sms * 8, V), 256, 0, st>>>(L, e->cfg.box_score_thresh, keys, counts);
    ret_class_nms_kernel<<<dim3(K, V), 1024, RET_NMS_SMEM, st>>>(L, keys, counts, d_image_hw, e->cfg.box_score_thresh,
                                                                 1e-2f, (double)e->cfg.box_nms_thresh, pc, kept);
    ret_gather_kernel<<<V, 1024, (K + 1) * 4, st>>>(L, kept, pc, d_ratio, e->det_cap, vs.det, vs.scores, e->d_overflow);
    CALD_CUDA_CHECK(cudaGetLastError());
    e->launches += 3;
    ar.free(keys); ar.free(counts); ar.free(kept.boxes); ar.free(kept.scores); ar.free(kept.anchor); ar.free(kept.count);
  }
  for (int l = 0; l < 5; ++l) { ar.free(cls_raw[l]); ar.free(reg_raw[l]); }
}

// ------------------------------------------------------------------------------------------------------------
ViewSet alloc_viewset(cald_engine* e, int V) {
  Arena& ar = e->arena;
  ViewSet vs;
  vs.V = V;
  const int dc = e->det_cap;
  vs.det.count = (int*)ar.alloc((size_t)V * 4);
  vs.det.boxes = (float4*)ar.alloc((size_t)V * dc * 16);
  vs.det.props = (float4*)ar.alloc((size_t)V * dc * 16);
  vs.det.scores = (float*)ar.alloc((size_t)V * dc * 4);
  vs.det.prob_max = (float*)ar.alloc((size_t)V * dc * 4);
  vs.det.labels = (int*)ar.alloc((size_t)V * dc * 4);
  vs.det.prop_idx = (int*)ar.alloc((size_t)V * dc * 4);
  vs.scores = (float*)ar.alloc((size_t)V * e->cap * e->C * 4);
  vs.prob_max = (float*)ar.alloc((size_t)V * e->cap * 4);
  return vs;
}
void free_viewset(cald_engine* e, ViewSet& vs) {
  Arena& ar = e->arena;
  ar.free(vs.det.count); ar.free(vs.det.boxes); ar.free(vs.det.props); ar.free(vs.det.scores);
  ar.free(vs.det.prob_max); ar.free(vs.det.labels); ar.free(vs.det.prop_idx); ar.free(vs.scores);
  ar.free(vs.prob_max);
}

// host description of one view before upload
struct HostView {
  const uint8_t* src;
  int sh, sw;
  int flip;
  int cut_slot;
  const float* noise = nullptr;  // device planes [3][sh][sw]
  int noise_mode = 0;
  float n0 = 0, n1 = 0;
  const int* mm = nullptr;       // device [2]: min / max of the source image (salt / pepper values)
  int perm = PERM_IDENTITY;
};

void resized_hw(const cald_config& c, int h, int w, int& rh, int& rw) {
  // tv:transform.py:57-62 + F.interpolate(recompute_scale_factor=True): python doubles
  double s = std::min((double)c.min_size / (double)std::min(h, w), (double)c.max_size / (double)std::max(h, w));
  rh = (int)std::floor((double)h * s);
  rw = (int)std::floor((double)w * s);
}
inline int pad32(int v) { return (int)(std::ceil((double)v / 32.0) * 32.0); }
constexpr double ARENA_BYTES_PER_PIXEL = 420.0;

// Run the detector over `views` (any mix of sizes): group by padded shape, forward each group, results in view order.
void detect_views(cald_engine* e, const std::vector<HostView>& views, const CutRects* d_cuts, ViewSet& out) {
  const int V = (int)views.size();
  std::vector<ViewDesc> vd(V);
  std::vector<int> hw(V * 2), php(V), pwp(V);
  std::vector<float> ratio(V * 2);
  for (int i = 0; i < V; ++i) {
    int rh, rw;
    resized_hw(e->cfg, views[i].sh, views[i].sw, rh, rw);
    vd[i] = ViewDesc{views[i].src, views[i].sh, views[i].sw, rh, rw, views[i].flip, views[i].cut_slot,
                     views[i].noise, views[i].noise_mode, views[i].n0, views[i].n1, views[i].mm, views[i].perm};
    hw[i * 2] = rh; hw[i * 2 + 1] = rw;
    if (e->retina) {  // tv:transform.py:306-311: fp32 tensors divided in fp32
      ratio[i * 2] = (float)views[i].sh / (float)rh;
      ratio[i * 2 + 1] = (float)views[i].sw / (float)rw;
    } else {          // frcnn_la.py:307-315: python doubles, then multiplied into the fp32 tensor
      ratio[i * 2] = (float)((double)views[i].sh / (double)rh);
      ratio[i * 2 + 1] = (float)((double)views[i].sw / (double)rw);
    }
    php[i] = pad32(rh); pwp[i] = pad32(rw);
  }
  // order views by (Hp, Wp) groups
  std::vector<int> order(V);
  for (int i = 0; i < V; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return php[a] != php[b] ? php[a] < php[b] : pwp[a] < pwp[b];
  });
  Arena& ar = e->arena;
  cudaStream_t st = e->st;
  const int maxv = e->views_per_pass;
  int pos = 0;
  while (pos < V) {
    int end = pos + 1;
    while (end < V && end - pos < maxv && php[order[end]] == php[order[pos]] && pwp[order[end]] == pwp[order[pos]]) ++end;
    const int n = end - pos;
    std::vector<ViewDesc> lvd(n);
    std::vector<int> lhw(n * 2);
    std::vector<float> lr(n * 2);
    for (int j = 0; j < n; ++j) {
      int g = order[pos + j];
      lvd[j] = vd[g];
      lhw[j * 2] = hw[g * 2]; lhw[j * 2 + 1] = hw[g * 2 + 1];
      lr[j * 2] = ratio[g * 2]; lr[j * 2 + 1] = ratio[g * 2 + 1];
    }
    ViewDesc* d_vd = (ViewDesc*)ar.alloc(n * sizeof(ViewDesc));
    int* d_hw = (int*)ar.alloc(n * 8);
    float* d_r = (float*)ar.alloc(n * 8);
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_vd, lvd.data(), n * sizeof(ViewDesc), cudaMemcpyHostToDevice, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_hw, lhw.data(), n * 8, cudaMemcpyHostToDevice, st));
    // (pageable-host cudaMemcpyAsync stages the source before returning, so the vectors may go out of scope)
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_r, lr.data(), n * 8, cudaMemcpyHostToDevice, st));
    const int dc = e->det_cap;
    const size_t srow = (size_t)e->cap * e->C * 4;
    bool contiguous = true;
    for (int j = 0; j < n; ++j) contiguous &= (order[pos + j] == order[pos] + j);
    if (contiguous) {
      // the group occupies consecutive global slots: let the pass write its results in place
      const int g0 = order[pos];
      ViewSet sub;
      sub.V = n;
      sub.det.count = out.det.count + g0;
      sub.det.boxes = out.det.boxes + (size_t)g0 * dc;
      sub.det.props = out.det.props + (size_t)g0 * dc;
      sub.det.scores = out.det.scores + (size_t)g0 * dc;
      sub.det.prob_max = out.det.prob_max + (size_t)g0 * dc;
      sub.det.labels = out.det.labels + (size_t)g0 * dc;
      sub.det.prop_idx = out.det.prop_idx + (size_t)g0 * dc;
      sub.scores = out.scores + (size_t)g0 * e->cap * e->C;
      sub.prob_max = out.prob_max + (size_t)g0 * e->cap;
      forward_pass(e, n, php[order[pos]], pwp[order[pos]], d_vd, d_cuts, d_hw, d_r, sub);
      ar.free(d_vd); ar.free(d_hw); ar.free(d_r);
      pos = end;
      continue;
    }
    ViewSet local = alloc_viewset(e, n);
    forward_pass(e, n, php[order[pos]], pwp[order[pos]], d_vd, d_cuts, d_hw, d_r, local);
    // scatter to global view slots
    for (int j = 0; j < n; ++j) {
      int g = order[pos + j];
      auto cp = [&](void* dst, const void* src, size_t bytes) {
        CALD_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
      };
      cp(out.det.count + g, local.det.count + j, 4);
      cp(out.det.boxes + (size_t)g * dc, local.det.boxes + (size_t)j * dc, dc * 16);
      cp(out.det.props + (size_t)g * dc, local.det.props + (size_t)j * dc, dc * 16);
      cp(out.det.scores + (size_t)g * dc, local.det.scores + (size_t)j * dc, dc * 4);
      cp(out.det.prob_max + (size_t)g * dc, local.det.prob_max + (size_t)j * dc, dc * 4);
      cp(out.det.labels + (size_t)g * dc, local.det.labels + (size_t)j * dc, dc * 4);
      cp(out.det.prop_idx + (size_t)g * dc, local.det.prop_idx + (size_t)j * dc, dc * 4);
      cp((char*)out.scores + (size_t)g * srow, (char*)local.scores + (size_t)j * srow, srow);
      cp(out.prob_max + (size_t)g * e->cap, local.prob_max + (size_t)j * e->cap, (size_t)e->cap * 4);
    }
    free_viewset(e, local);
    ar.free(d_vd); ar.free(d_hw); ar.free(d_r);
    pos = end;
  }
}

uint8_t* pil_resize_device(cald_engine* e, const uint8_t* src, int h, int w, int oh, int ow, int filter);
void check_overflow(cald_engine* e);
}  // namespace

// stage-level entry point (include/cald_b200_ops.h); needs no weights
extern "C" int cald_op_aug_image(int kind, const uint8_t* img, int h, int w, uint8_t* out, int* out_h, int* out_w) {
  try {
    std::unique_ptr<cald_engine> e(new cald_engine());
    memset(&e->cfg, 0, sizeof(e->cfg));
    CALD_CUDA_CHECK(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    e->arena.init((size_t)64 << 20);
    uint8_t* d = (uint8_t*)e->arena.alloc((size_t)h * w * 3);
    CALD_CUDA_CHECK(cudaMemcpyAsync(d, img, (size_t)h * w * 3, cudaMemcpyHostToDevice, e->st));
    uint8_t* r = nullptr;
    int oh = h, ow = w;
    if (kind == CALD_AUG_SMALLER_RESIZE) {
      ow = (int)(w * 0.8); oh = (int)(h * 0.8);
      r = pil_resize_device(e.get(), d, h, w, oh, ow, 0);
    } else if (kind == CALD_AUG_ROTATION) {
      RotateGeom rg = pil_rotate_geom(w, h, 5.0);
      uint8_t* t = (uint8_t*)e->arena.alloc((size_t)rg.nh * rg.nw * 3);
      pil_rotate_nearest_kernel<<<dim3((rg.nw + 127) / 128, rg.nh), 128, 0, e->st>>>(d, t, w, h, rg);
      r = pil_resize_device(e.get(), t, rg.nh, rg.nw, h, w, 1);
    } else {
      throw std::runtime_error("cald_op_aug_image: kind must be 2 (smaller_resize) or 3 (rotation)");
    }
    CALD_CUDA_CHECK(cudaMemcpyAsync(out, r, (size_t)oh * ow * 3, cudaMemcpyDeviceToHost, e->st));
    CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
    *out_h = oh; *out_w = ow;
    return 0;
  } catch (const std::exception& ex) {
    g_create_err = ex.what();
    cudaGetLastError();
    return -1;
  }
}

// stage-level entry point: ColorAdjust(image, factor) on the device (cald_helper.py:65-69)
extern "C" int cald_op_color_adjust(const uint8_t* img, int h, int w, double factor, uint8_t* out) {
  try {
    const long long npix = (long long)h * w;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    unsigned long long* d_sum = nullptr;
    CALD_CUDA_CHECK(cudaMalloc((void**)&d_in, npix * 3));
    CALD_CUDA_CHECK(cudaMalloc((void**)&d_out, npix * 3));
    CALD_CUDA_CHECK(cudaMalloc((void**)&d_sum, 8));
    CALD_CUDA_CHECK(cudaMemcpy(d_in, img, npix * 3, cudaMemcpyHostToDevice));
    CALD_CUDA_CHECK(cudaMemset(d_sum, 0, 8));
    const int blocks = (int)std::min<long long>((npix + 255) / 256, 148 * 8);
    pil_brightness_kernel<<<blocks, 256>>>(d_in, d_out, npix, (float)factor, d_sum);
    pil_contrast_saturation_kernel<<<blocks, 256>>>(d_out, npix, (float)factor, d_sum);
    CALD_CUDA_CHECK(cudaGetLastError());
    CALD_CUDA_CHECK(cudaMemcpy(out, d_out, npix * 3, cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_sum);
    return 0;
  } catch (const std::exception& ex) {
    g_create_err = ex.what();
    cudaGetLastError();
    return -1;
  }
}

namespace {
// Pillow-exact resize of a device u8 image (horizontal pass then vertical pass).
// A real pool has thousands of distinct image sizes: the coefficient cache is emptied once it holds PIL_CACHE_MAX
// tables.  Called at the start of a chunk only (tables handed out during a chunk stay valid until its kernels are
// enqueued) and after draining the stream (earlier kernels may still read them).
void pil_cache_trim(cald_engine* e) {
  constexpr size_t PIL_CACHE_MAX = 256;
  if (e->pil_cache.size() < PIL_CACHE_MAX) return;
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
  for (auto& kv : e->pil_cache) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
  e->pil_cache.clear();
  e->pil_ksize.clear();
}

// device coefficient tables of one Pillow resampling pass (cached per (in, out, filter))
void pil_coeffs(cald_engine* e, int in, int out, int filter, int*& d_bounds, int*& d_kk, int& ksize) {
  {
    long long key = ((long long)in << 34) | ((long long)out << 4) | filter;
    auto it = e->pil_cache.find(key);
    if (it == e->pil_cache.end()) {
      PilCoeffs pc = pil_precompute(in, out, filter);
      int *b, *k;
      CALD_CUDA_CHECK(cudaMalloc((void**)&b, pc.bounds.size() * 4));
      CALD_CUDA_CHECK(cudaMalloc((void**)&k, pc.kk.size() * 4));
      CALD_CUDA_CHECK(cudaMemcpy(b, pc.bounds.data(), pc.bounds.size() * 4, cudaMemcpyHostToDevice));
      CALD_CUDA_CHECK(cudaMemcpy(k, pc.kk.data(), pc.kk.size() * 4, cudaMemcpyHostToDevice));
      e->pil_cache[key] = {b, k};
      e->pil_ksize[key] = pc.ksize;
      it = e->pil_cache.find(key);
    }
    d_bounds = it->second.first;
    d_kk = it->second.second;
    ksize = e->pil_ksize[key];
  }
}

uint8_t* pil_resize_device(cald_engine* e, const uint8_t* src, int h, int w, int oh, int ow, int filter) {
  Arena& ar = e->arena;
  cudaStream_t st = e->st;
  auto coeffs = [&](int in, int out, int*& d_bounds, int*& d_kk, int& ksize) {
    pil_coeffs(e, in, out, filter, d_bounds, d_kk, ksize);
  };
  const uint8_t* cur = src;
  uint8_t* tmp = nullptr;
  if (ow != w) {
    int *b, *k, ks;
    coeffs(w, ow, b, k, ks);
    tmp = (uint8_t*)ar.alloc((size_t)h * ow * 3);
    pil_resample_h_kernel<<<dim3((ow * 3 + 255) / 256, h), 256, 0, st>>>(cur, tmp, h, w, ow, b, k, ks);
    KLAUNCH(e);
    cur = tmp;
  }
  uint8_t* out = (uint8_t*)ar.alloc((size_t)oh * ow * 3);
  if (oh != h) {
    int *b, *k, ks;
    coeffs(h, oh, b, k, ks);
    pil_resample_v_kernel<<<dim3((ow * 3 + 255) / 256, oh), 256, 0, st>>>(cur, out, h, oh, ow, b, k, ks);
    KLAUNCH(e);
  } else {
    CALD_CUDA_CHECK(cudaMemcpyAsync(out, cur, (size_t)oh * ow * 3, cudaMemcpyDeviceToDevice, st));
  }
  CALD_CUDA_CHECK(cudaGetLastError());
  if (tmp) ar.free(tmp);
  return out;
}

// The affine of cald_helper.rotate's box transform (cald_helper.py:146-195): float64 math, then .float()
void rotate_box_geom(int w, int h, double angle_deg, int rot_w, int rot_h, AugGeom& g) {
  const double PI = 3.14159265358979323846;
  double ang = angle_deg * PI / 180.0;  // np.radians
  double alpha = cos(ang), beta = sin(ang);
  double cx = w / 2.0, cy = h / 2.0;
  double m[6] = {alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy};
  double c = fabs(m[0]), s = fabs(m[1]);
  int nW = (int)((h * s) + (w * c));
  int nH = (int)((h * c) + (w * s));
  m[2] += (nW / 2.0) - cx;
  m[5] += (nH / 2.0) - cy;
  for (int i = 0; i < 6; ++i) g.m[i] = (float)m[i];
  g.sx = (float)((double)rot_w / w);
  g.sy = (float)((double)rot_h / h);
}

// min / max of a device u8 image as to_tensor values (salt / pepper of cald_helper.py:81-82)
// grid = (blocks, B): one launch for the whole chunk; out[b] = {min, max}, pre-set to {255, 0}
__global__ void u8_minmax_kernel(const ViewDesc* __restrict__ imgs, int* __restrict__ out_all) {
  const ViewDesc d = imgs[blockIdx.y];
  const uint8_t* img = d.src;
  const long long n = (long long)d.sh * d.sw * 3;
  int* out = out_all + blockIdx.y * 2;
  int lo = 255, hi = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int v = img[i];
    lo = min(lo, v); hi = max(hi, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&out[0], lo); atomicMax(&out[1], hi); }
}

void score_chunk(cald_engine* e, int B, const uint8_t* const* d_images, const int* hs, const int* ws,
                 const std::vector<cald_aug>& augs, double bp, const double* d_u, int n_u, int* d_cursor,
                 const float* const* d_noise /* [B][n_noise] device planes */,
                 const int* swap_perms /* [B][n_swap] host, or null */, double* out_cons, double* out_cls,
                 int scorer = 0 /* 0: CALD consistency, 1: LS+C stability (ls_c_train.py:108-155) */,
                 const std::function<void()>& before_wait = nullptr /* runs once everything is enqueued */) {
  Arena& ar = e->arena;
  cudaStream_t st = e->st;
  const int A = (int)augs.size();
  const int C = e->C, ncls1 = C - 1, dc = e->det_cap;
  // ---------------- reference views
  std::vector<HostView> rv(B);
  for (int b = 0; b < B; ++b) rv[b] = HostView{d_images[b], hs[b], ws[b], 0, -1};
  ViewSet ref = alloc_viewset(e, B);
  pil_cache_trim(e);
  e->trace("chunk_start");
  detect_views(e, rv, nullptr, ref);
  e->trace("ref_pass_enqueued");
  // ---------------- reference sub-sample, class vectors, cutout rects, boxes in aug coordinates
  RefSet rs;
  rs.n = (int*)ar.alloc(B * 4);
  rs.n_det = (int*)ar.alloc(B * 4);
  rs.boxes = (float4*)ar.alloc((size_t)B * REF_CAP * 16);
  rs.prob_max = (float*)ar.alloc((size_t)B * REF_CAP * 4);
  rs.prop_idx = (int*)ar.alloc((size_t)B * REF_CAP * 4);
  float* d_cls = (float*)ar.alloc((size_t)B * (1 + A) * ncls1 * 4);  // [B] ref rows, then [B*A] aug rows
  ref_prepare_kernel<<<B, 64, 0, st>>>(ref.det, dc, e->d_lut, rs);
  class_max_kernel<<<B, 128, ncls1 * 4, st>>>(ref.det, dc, ncls1, e->d_lut, 1, d_cls);
  e->launches += 2;
  // ---------------- cutout rectangles: every cutout view consumes the caller's RNG stream in view order
  std::vector<int> cut_nums;
  int n_noise = 0;
  for (const cald_aug& a : augs) {
    if (a.kind == CALD_AUG_CUTOUT) cut_nums.push_back((int)a.param);
    if (a.kind == CALD_AUG_GAUSS || a.kind == CALD_AUG_SALTPEPPER) n_noise++;
  }
  const int n_cut = (int)cut_nums.size();
  int n_swap = 0;
  for (const cald_aug& a : augs) n_swap += (a.kind == CALD_AUG_COLOR_SWAP);
  for (int c : cut_nums) if (c < 1 || c > MAX_CUT) throw std::runtime_error("cutout: cut_num must be 1..4");
  CutRects* d_cuts = (CutRects*)ar.alloc((size_t)B * std::max(1, n_cut) * sizeof(CutRects));
  CALD_CUDA_CHECK(cudaMemsetAsync(d_cuts, 0, (size_t)B * std::max(1, n_cut) * sizeof(CutRects), st));
  std::vector<int> img_hw(B * 2);
  for (int b = 0; b < B; ++b) { img_hw[b * 2] = hs[b]; img_hw[b * 2 + 1] = ws[b]; }
  int* d_img_hw = (int*)ar.alloc(B * 8);
  CALD_CUDA_CHECK(cudaMemcpyAsync(d_img_hw, img_hw.data(), B * 8, cudaMemcpyHostToDevice, st));
  if (n_cut) {
    int* d_cn = (int*)ar.alloc(n_cut * 4);
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_cn, cut_nums.data(), n_cut * 4, cudaMemcpyHostToDevice, st));
    cutout_kernel<<<1, 64, 0, st>>>(rs, d_img_hw, B, n_cut, d_cn, d_u, n_u, d_cuts, d_cursor);
    KLAUNCH(e);
    ar.free(d_cn);
  }
  // salt / pepper values need each image's min / max: one launch for the chunk, the values stay on the device
  // (the stem-input kernel reads them through ViewDesc::mm), so there is no host round trip
  bool has_sp = false;
  for (const cald_aug& a : augs) has_sp |= (a.kind == CALD_AUG_SALTPEPPER);
  int* d_mm = nullptr;
  if (has_sp) {
    d_mm = (int*)ar.alloc(B * 8);
    std::vector<int> init(B * 2);
    std::vector<ViewDesc> srcs(B);
    for (int b = 0; b < B; ++b) {
      init[b * 2] = 255; init[b * 2 + 1] = 0;
      memset(&srcs[b], 0, sizeof(ViewDesc));
      srcs[b].src = d_images[b]; srcs[b].sh = hs[b]; srcs[b].sw = ws[b];
    }
    ViewDesc* d_srcs = (ViewDesc*)ar.alloc(B * sizeof(ViewDesc));
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_mm, init.data(), B * 8, cudaMemcpyHostToDevice, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_srcs, srcs.data(), B * sizeof(ViewDesc), cudaMemcpyHostToDevice, st));
    u8_minmax_kernel<<<dim3(64, B), 256, 0, st>>>(d_srcs, d_mm);
    CALD_CUDA_CHECK(cudaGetLastError());
    KLAUNCH(e);
    ar.free(d_srcs);  // stream-ordered reuse
  }
  // ---------------- augmented views
  std::vector<HostView> av((size_t)B * A);
  std::vector<AugGeom> geom((size_t)B * A);
  std::vector<uint8_t*> temps;
  // Pillow-exact resize / rotate of the whole chunk in three launches (nearest rotate, horizontal pass, vertical pass)
  std::vector<RotJob> rot_jobs;
  std::vector<PilJob> h_jobs, v_jobs;
  std::vector<uint8_t*> pil_scratch;
  // two-pass resample src (h x w) -> (oh x ow) queued as one horizontal and one vertical job; returns the output
  auto queue_resize = [&](const uint8_t* src, int h, int w, int oh, int ow, int filter) -> uint8_t* {
    if (ow == w || oh == h) return nullptr;  // a pass Pillow would skip: the per-image path handles it
    uint8_t* tmp = (uint8_t*)ar.alloc((size_t)h * ow * 3);
    uint8_t* out = (uint8_t*)ar.alloc((size_t)oh * ow * 3);
    pil_scratch.push_back(tmp);
    PilJob hj{src, tmp, h, w, h, ow, nullptr, nullptr, 0}, vj{tmp, out, h, ow, oh, ow, nullptr, nullptr, 0};
    int *b, *k, ks;
    pil_coeffs(e, w, ow, filter, b, k, ks);
    hj.bounds = b; hj.kk = k; hj.ksize = ks;
    pil_coeffs(e, h, oh, filter, b, k, ks);
    vj.bounds = b; vj.kk = k; vj.ksize = ks;
    h_jobs.push_back(hj);
    v_jobs.push_back(vj);
    return out;
  };
  for (int b = 0; b < B; ++b) {
    int cut_i = 0, noise_i = 0, swap_i = 0;
    for (int a = 0; a < A; ++a) {
      AugGeom& g = geom[(size_t)b * A + a];
      memset(&g, 0, sizeof(g));
      g.w = (float)ws[b]; g.h = (float)hs[b];
      HostView hv{d_images[b], hs[b], ws[b], 0, -1};
      const double prm = augs[a].param;
      switch (augs[a].kind) {
        case CALD_AUG_FLIP: g.kind = AUG_FLIP; hv.flip = 1; break;
        case CALD_AUG_CUTOUT: g.kind = AUG_IDENT; hv.cut_slot = b * n_cut + cut_i++; break;
        case CALD_AUG_RESIZE: {
          g.kind = AUG_RESIZE;
          g.ratio = (float)prm;
          int ow = (int)(ws[b] * prm), oh = (int)(hs[b] * prm);
          uint8_t* t = queue_resize(d_images[b], hs[b], ws[b], oh, ow, 0);
          if (!t) t = pil_resize_device(e, d_images[b], hs[b], ws[b], oh, ow, 0);
          temps.push_back(t);
          hv.src = t; hv.sh = oh; hv.sw = ow;
          break;
        }
        case CALD_AUG_ROTATION: {
          g.kind = AUG_ROTATE;
          RotateGeom rg = pil_rotate_geom(ws[b], hs[b], prm);
          uint8_t* r = (uint8_t*)ar.alloc((size_t)rg.nh * rg.nw * 3);
          uint8_t* t = queue_resize(r, rg.nh, rg.nw, hs[b], ws[b], 1);
          if (t) {
            rot_jobs.push_back(RotJob{d_images[b], r, ws[b], hs[b], rg});
            pil_scratch.push_back(r);
          } else {
            pil_rotate_nearest_kernel<<<dim3((rg.nw + 127) / 128, rg.nh), 128, 0, st>>>(d_images[b], r, ws[b], hs[b], rg);
            KLAUNCH(e);
            t = pil_resize_device(e, r, rg.nh, rg.nw, hs[b], ws[b], 1);
            ar.free(r);
          }
          temps.push_back(t);
          hv.src = t;
          rotate_box_geom(ws[b], hs[b], prm, rg.nw, rg.nh, g);
          break;
        }
        case CALD_AUG_GAUSS:
        case CALD_AUG_SALTPEPPER: {
          g.kind = AUG_IDENT;
          if (!d_noise) throw std::runtime_error("noise augmentation requested but no noise planes were passed");
          hv.noise = d_noise[(size_t)b * n_noise + noise_i++];
          if (augs[a].kind == CALD_AUG_GAUSS) {
            hv.noise_mode = 1;
            hv.n0 = (float)prm;  // torch: randn * std (python scalar -> fp32) / 255.0
          } else {
            hv.noise_mode = 2;
            hv.n0 = (float)(prm / 2.0);
            hv.n1 = (float)(1.0 - prm / 2.0);
            hv.mm = d_mm + b * 2;
          }
          break;
        }
        case CALD_AUG_COLOR_ADJUST: {
          g.kind = AUG_IDENT;
          const long long npix = (long long)hs[b] * ws[b];
          uint8_t* t = (uint8_t*)ar.alloc((size_t)npix * 3);
          unsigned long long* d_sum = (unsigned long long*)ar.alloc(8);
          CALD_CUDA_CHECK(cudaMemsetAsync(d_sum, 0, 8, st));
          const int blocks = (int)std::min<long long>((npix + 255) / 256, (long long)e->conv.num_sms * 8);
          pil_brightness_kernel<<<blocks, 256, 0, st>>>(d_images[b], t, npix, (float)prm, d_sum);
          pil_contrast_saturation_kernel<<<blocks, 256, 0, st>>>(t, npix, (float)prm, d_sum);
          CALD_CUDA_CHECK(cudaGetLastError());
          e->launches += 2;
          ar.free(d_sum);  // stream-ordered reuse: the kernels above are already enqueued
          temps.push_back(t);
          hv.src = t;
          break;
        }
        case CALD_AUG_COLOR_SWAP: {
          // image[swap, :, :] with swap = perms[random.randint(0, 5)] drawn by the caller (cald_helper.py:56-62)
          static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
          g.kind = AUG_IDENT;
          const int pi = swap_perms ? swap_perms[(size_t)b * n_swap + swap_i++] : (int)prm;
          if (pi < 0 || pi > 5) throw std::runtime_error("color_swap: permutation index must be 0..5");
          hv.perm = perms[pi][0] | (perms[pi][1] << 2) | (perms[pi][2] << 4);
          break;
        }
        default: throw std::runtime_error("unsupported augmentation kind");
      }
      av[(size_t)b * A + a] = hv;
    }
  }
  if (!h_jobs.empty()) {
    const size_t nr = rot_jobs.size(), nj = h_jobs.size();
    RotJob* d_rot = nr ? (RotJob*)ar.alloc(nr * sizeof(RotJob)) : nullptr;
    PilJob* d_hv = (PilJob*)ar.alloc(2 * nj * sizeof(PilJob));
    if (nr) CALD_CUDA_CHECK(cudaMemcpyAsync(d_rot, rot_jobs.data(), nr * sizeof(RotJob), cudaMemcpyHostToDevice, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_hv, h_jobs.data(), nj * sizeof(PilJob), cudaMemcpyHostToDevice, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_hv + nj, v_jobs.data(), nj * sizeof(PilJob), cudaMemcpyHostToDevice, st));
    int rw = 0, rh = 0, hw = 0, hh = 0, vw = 0, vh = 0;
    for (const RotJob& j : rot_jobs) { rw = std::max(rw, j.g.nw); rh = std::max(rh, j.g.nh); }
    for (const PilJob& j : h_jobs) { hw = std::max(hw, j.out_w); hh = std::max(hh, j.in_h); }
    for (const PilJob& j : v_jobs) { vw = std::max(vw, j.in_w); vh = std::max(vh, j.out_h); }
    if (nr) {
      pil_rotate_nearest_batched_kernel<<<dim3((rw + 127) / 128, rh, (unsigned)nr), 128, 0, st>>>(d_rot);
      KLAUNCH(e);
    }
    pil_resample_h_batched_kernel<<<dim3((hw + 127) / 128, hh, (unsigned)nj), 128, 0, st>>>(d_hv);
    pil_resample_v_batched_kernel<<<dim3(((vw * 3 + 3) / 4 + 127) / 128, vh, (unsigned)nj), 128, 0, st>>>(d_hv + nj);
    CALD_CUDA_CHECK(cudaGetLastError());
    e->launches += 2;
    if (d_rot) ar.free(d_rot);
    ar.free(d_hv);
    for (uint8_t* t : pil_scratch) ar.free(t);  // stream-ordered reuse: the passes above are enqueued
  }
  ViewSet aug;
  float* d_cons = nullptr;
  if (A > 0) {
    AugGeom* d_geom = (AugGeom*)ar.alloc(geom.size() * sizeof(AugGeom));
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_geom, geom.data(), geom.size() * sizeof(AugGeom), cudaMemcpyHostToDevice, st));
    float4* d_augb = (float4*)ar.alloc((size_t)B * A * REF_CAP * 16);
    aug_boxes_kernel<<<B * A, 64, 0, st>>>(rs, d_geom, A, d_augb);
    KLAUNCH(e);
    aug = alloc_viewset(e, B * A);
    e->trace("aug_prep_enqueued");
    detect_views(e, av, d_cuts, aug);
    e->trace("aug_pass_enqueued");
    if (scorer == 1) {
      // LS+C: top-30 reference boxes by prob_max, mean over the A noise views of the best IoU, prob_max-weighted mean
      if (A > 8) throw std::runtime_error("LS+C: at most 8 augmented views");
      unsigned long long* keys = (unsigned long long*)ar.alloc((size_t)B * dc * 8);
      unsigned long long* top = (unsigned long long*)ar.alloc((size_t)B * TOPK_MAX * 8);
      int* top_count = (int*)ar.alloc((size_t)B * 4);
      double* d_out = (double*)ar.alloc((size_t)B * 8);
      lsc_keys_kernel<<<B, 128, 0, st>>>(ref.det, dc, keys);
      TopkGroups tg;
      memset(&tg, 0, sizeof(tg));
      tg.keys = keys; tg.stride_outer = dc; tg.inner = 1; tg.dyn_n = ref.det.count; tg.k = LSC_REF;
      topk_select_kernel<<<B, 1024, 0, st>>>(tg, top, top_count);
      lsc_kernel<<<B, 32 * CONS_WARPS, 0, st>>>(ref.det, aug.det, dc, A, top, d_out);
      CALD_CUDA_CHECK(cudaGetLastError());
      e->launches += 3;
      // (a device -> PAGEABLE-host cudaMemcpyAsync returns only when the copy is done, i.e. after everything queued
      // before it: the next chunk's upload has to be started BEFORE the result copies, not after them)
      if (before_wait) before_wait();
      CALD_CUDA_CHECK(cudaMemcpyAsync(out_cons, d_out, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
      std::vector<int> h_rc(B);
      CALD_CUDA_CHECK(cudaMemcpyAsync(h_rc.data(), ref.det.count, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
      CALD_CUDA_CHECK(cudaStreamSynchronize(st));
      e->last_ref_counts.insert(e->last_ref_counts.end(), h_rc.begin(), h_rc.end());
      check_overflow(e);
      ar.free(keys); ar.free(top); ar.free(top_count); ar.free(d_out);
    } else {
      class_max_kernel<<<B * A, 128, ncls1 * 4, st>>>(aug.det, dc, ncls1, e->d_lut, 0, d_cls + (size_t)B * ncls1);
      d_cons = (float*)ar.alloc((size_t)B * A * 4);
      ConsArgs ca;
      ca.ref = rs; ca.aug_boxes = d_augb; ca.det = aug.det; ca.det_cap = dc;
      ca.ref_scores = ref.scores; ca.aug_scores = aug.scores; ca.cap = e->cap; ca.C = C; ca.A = A;
      ca.bp = (float)bp; ca.out = d_cons;
      consistency_kernel<<<dim3(A, B), 32 * CONS_WARPS, 0, st>>>(ca);
      CALD_CUDA_CHECK(cudaGetLastError());
      e->launches += 2;
    }
    ar.free(d_geom);
    ar.free(d_augb);
  }
  // ---------------- results to host; final means in double as numpy does (cald_train.py:225-228)
  if (scorer == 0) {
    std::vector<float> h_cons((size_t)B * std::max(A, 1)), h_cls((size_t)B * (1 + A) * ncls1);
    std::vector<int> h_ndet(B);
    // Everything of this chunk is enqueued: start the next chunk's upload / decode NOW.  The result copies below go to
    // pageable host memory, and a device -> pageable cudaMemcpyAsync only returns once the copy has completed -- i.e.
    // after the whole chunk has run.  (Round 2 first had the prefetch behind these copies: the timeline showed chunk
    // c+1's decode starting exactly when chunk c's compute ended, nothing overlapped.)
    if (before_wait) before_wait();
    if (A > 0) CALD_CUDA_CHECK(cudaMemcpyAsync(h_cons.data(), d_cons, (size_t)B * A * 4, cudaMemcpyDeviceToHost, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(h_cls.data(), d_cls, h_cls.size() * 4, cudaMemcpyDeviceToHost, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(h_ndet.data(), rs.n_det, B * 4, cudaMemcpyDeviceToHost, st));
    CALD_CUDA_CHECK(cudaStreamSynchronize(st));
    check_overflow(e);
    e->trace("results_on_host");
    e->last_ref_counts.insert(e->last_ref_counts.end(), h_ndet.begin(), h_ndet.end());
    for (int b = 0; b < B; ++b) {
      double* cls = out_cls + (size_t)b * ncls1;
      if (h_ndet[b] == 0 || A == 0) {
        // empty reference prediction: consistency 0.0, class vector = the (all-zero) reference row (cald_train.py:118-121);
        // no augmentation at all: np.mean([]) = nan (cald_train.py:225) and the class vector is the reference row
        out_cons[b] = h_ndet[b] == 0 ? 0.0 : std::nan("");
        for (int c = 0; c < ncls1; ++c) cls[c] = (double)h_cls[(size_t)b * ncls1 + c];
        for (int a = 0; a < A; ++a) e->last_per_view.push_back(0.f);
        continue;
      }
      double s = 0.0;
      for (int a = 0; a < A; ++a) { s += (double)h_cons[(size_t)b * A + a]; e->last_per_view.push_back(h_cons[(size_t)b * A + a]); }
      out_cons[b] = s / A;
      for (int c = 0; c < ncls1; ++c) {
        double t = (double)h_cls[(size_t)b * ncls1 + c];
        for (int a = 0; a < A; ++a) t += (double)h_cls[((size_t)B + (size_t)b * A + a) * ncls1 + c];
        cls[c] = t / (1 + A);
      }
    }
  }
  if (e->cfg.debug && scorer == 0) {
    auto fetch = [&](const ViewSet& vs, int V, std::vector<int>& cnt, std::vector<float>& bx, std::vector<float>& sc,
                     std::vector<float>& pm, std::vector<int>& lb) {
      cnt.resize(V); bx.resize((size_t)V * dc * 4); sc.resize((size_t)V * dc); pm.resize((size_t)V * dc); lb.resize((size_t)V * dc);
      CALD_CUDA_CHECK(cudaMemcpyAsync(cnt.data(), vs.det.count, (size_t)V * 4, cudaMemcpyDeviceToHost, st));
      CALD_CUDA_CHECK(cudaMemcpyAsync(bx.data(), vs.det.boxes, bx.size() * 4, cudaMemcpyDeviceToHost, st));
      CALD_CUDA_CHECK(cudaMemcpyAsync(sc.data(), vs.det.scores, sc.size() * 4, cudaMemcpyDeviceToHost, st));
      CALD_CUDA_CHECK(cudaMemcpyAsync(pm.data(), vs.det.prob_max, pm.size() * 4, cudaMemcpyDeviceToHost, st));
      CALD_CUDA_CHECK(cudaMemcpyAsync(lb.data(), vs.det.labels, lb.size() * 4, cudaMemcpyDeviceToHost, st));
      CALD_CUDA_CHECK(cudaStreamSynchronize(st));
    };
    std::vector<int> rc, ac, rl, al;
    std::vector<float> rb, rsc, rpm, ab, asc, apm;
    fetch(ref, B, rc, rb, rsc, rpm, rl);
    if (A > 0) fetch(aug, B * A, ac, ab, asc, apm, al);
    auto push = [&](int v, const std::vector<int>& cnt, const std::vector<float>& bx, const std::vector<float>& sc,
                    const std::vector<float>& pm, const std::vector<int>& lb) {
      const int n = cnt[v];
      e->dbg_views.counts.push_back(n);
      e->dbg_views.boxes.insert(e->dbg_views.boxes.end(), bx.begin() + (size_t)v * dc * 4, bx.begin() + ((size_t)v * dc + n) * 4);
      e->dbg_views.scores.insert(e->dbg_views.scores.end(), sc.begin() + (size_t)v * dc, sc.begin() + (size_t)v * dc + n);
      e->dbg_views.prob_max.insert(e->dbg_views.prob_max.end(), pm.begin() + (size_t)v * dc, pm.begin() + (size_t)v * dc + n);
      e->dbg_views.labels.insert(e->dbg_views.labels.end(), lb.begin() + (size_t)v * dc, lb.begin() + (size_t)v * dc + n);
    };
    for (int b = 0; b < B; ++b) {
      push(b, rc, rb, rsc, rpm, rl);
      for (int a = 0; a < A; ++a) push(b * A + a, ac, ab, asc, apm, al);
    }
  }
  if (A > 0) { free_viewset(e, aug); if (d_cons) ar.free(d_cons); }
  for (uint8_t* t : temps) ar.free(t);
  if (d_mm) ar.free(d_mm);
  ar.free(d_img_hw); ar.free(d_cuts); ar.free(d_cls);
  ar.free(rs.n); ar.free(rs.n_det); ar.free(rs.boxes); ar.free(rs.prob_max); ar.free(rs.prop_idx);
  free_viewset(e, ref);
}

// ------------------------------------------------------------------------------------------------------------
// Image upload pipeline.  A call's images are scored in chunks; chunk i+1 is copied host -> device on the engine's copy
// stream into the other of two fixed device slabs while chunk i computes, so the H2D time disappears behind the
// forward passes.  Caller buffers that are page-locked are DMA'd in place; pageable ones are gathered into a
// page-locked staging buffer first (one CPU memcpy, then a single link-speed DMA).
// ------------------------------------------------------------------------------------------------------------
struct DeviceImages {
  std::vector<const uint8_t*> ptr;
};
struct UploadPipe {
  cald_engine* e;
  int n_images, per_chunk;
  const uint8_t* const* imgs;   // raw mode: u8 HWC images; JPEG mode: the files
  const int* hs;
  const int* ws;
  uint8_t* slab[2] = {nullptr, nullptr};
  size_t slab_bytes = 0;
  bool all_pinned = true;
  int started = 0;  // chunks whose copy has been enqueued
  // ---- JPEG mode (pool ingest, jpeg.cuh): the compressed scans travel instead of pixels and the chunk is decoded
  //      into its slab by three kernels on the copy stream, concurrently with the previous chunk's forward passes
  bool jpeg = false;
  // Where the entropy-coded segment is walked.  Host (default): a few host threads walk the chunk's scans into the
  // page-locked staging buffer while the previous chunk computes (the walk is a serial bit parser, ~1 ms per image on a
  // CPU core against ~50 ms on one GPU lane) and the coefficients travel; the device does dequantisation, inverse DCT,
  // upsampling and colour conversion with short-lived blocks.  Device (CALD_JPEG_WALK=device): the compressed scan
  // travels (20x fewer bytes) and one GPU lane per image walks it -- but those blocks live ~100 ms beside the
  // persistent conv CTAs and cost the conv kernels ~15 % while they run (measured, profiles/r02_summary.md).
  bool host_walk = true;
  std::vector<JpegImage> meta;                 // per file, offsets relative to its chunk
  std::vector<std::vector<JpegHuff>> tabs;     // per file
  std::vector<size_t> scan_begin, scan_end;
  std::vector<int> jh, jw;
  size_t cap_bytes = 0, cap_coef = 0, cap_planes = 0, cap_tabs = 0;
  uint8_t* d_bytes[2] = {nullptr, nullptr};
  JpegImage* d_meta[2] = {nullptr, nullptr};
  JpegHuff* d_tabs[2] = {nullptr, nullptr};
  short* d_coef[2] = {nullptr, nullptr};
  uint8_t* d_planes[2] = {nullptr, nullptr};
  int* d_sched[2] = {nullptr, nullptr};        // jpeg_huffman_kernel's image counter + per-SM claim flags
  std::vector<JpegImage> h_meta[2];            // host copies must outlive the async H2D of their slot
  std::vector<JpegHuff> h_tabs[2];
  // CALD_TRACE_JPEG=1: device time of every chunk's decode (copy stream), printed when the pipe is destroyed
  bool trace_jpeg = getenv("CALD_TRACE_JPEG") != nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trace_ev;
  std::vector<cudaEvent_t> trace_chunk;   // compute stream: the point where chunk c may start (after its wait)
  cudaEvent_t trace_base = nullptr;
  ~UploadPipe() {
    if (!trace_jpeg || trace_ev.empty()) return;
    cudaStreamSynchronize(e->copy_st);
    cudaStreamSynchronize(e->st);
    fprintf(stderr, "[jpeg timeline, ms since call start] decode(c) = start..end on the copy stream; compute(c) starts at\n");
    for (size_t c = 0; c < trace_ev.size(); ++c) {
      float a = 0, b = 0, g = 0;
      cudaEventElapsedTime(&a, trace_base, trace_ev[c].first);
      cudaEventElapsedTime(&b, trace_base, trace_ev[c].second);
      if (c < trace_chunk.size()) cudaEventElapsedTime(&g, trace_base, trace_chunk[c]);
      fprintf(stderr, "  chunk %zu: decode %.1f..%.1f  compute starts %.1f\n", c, a, b, g);
      cudaEventDestroy(trace_ev[c].first); cudaEventDestroy(trace_ev[c].second);
    }
    for (cudaEvent_t ev : trace_chunk) cudaEventDestroy(ev);
    cudaEventDestroy(trace_base);
  }

  static size_t padded(int h, int w) { return ((size_t)h * w * 3 + 255) & ~(size_t)255; }
  int n_chunks() const { return (n_images + per_chunk - 1) / per_chunk; }

  // Both slabs are taken from the arena before anything else of the call and live until its end: arena blocks freed
  // and re-used during the call are only ordered on the compute stream, never against the copy stream.
  UploadPipe(cald_engine* eng, int n, const uint8_t* const* images, const int* heights, const int* widths, int chunk,
             const size_t* file_sizes = nullptr)
      : e(eng), n_images(n), per_chunk(std::max(1, chunk)), imgs(images), hs(heights), ws(widths) {
    jpeg = file_sizes != nullptr;
    if (const char* m = getenv("CALD_JPEG_WALK")) host_walk = strcmp(m, "device") != 0;
    if (jpeg && trace_jpeg) { cudaEventCreate(&trace_base); cudaEventRecord(trace_base, e->st); }
    if (jpeg) {
      meta.resize(n); tabs.resize(n); scan_begin.resize(n); scan_end.resize(n); jh.resize(n); jw.resize(n);
      for (int i = 0; i < n; ++i) {
        try {
          jpeg_parse(images[i], file_sizes[i], meta[i], tabs[i], scan_begin[i], scan_end[i]);
        } catch (const std::exception& ex) {
          throw std::runtime_error("file " + std::to_string(i) + ": " + ex.what());
        }
        jh[i] = meta[i].height; jw[i] = meta[i].width;
      }
      hs = jh.data(); ws = jw.data();
    }
    for (int c = 0; c < n_chunks(); ++c) {
      size_t b = 0, cb = 0, cc = 0, cp = 0, ct = 0;
      for (int i = c * per_chunk; i < std::min(n, (c + 1) * per_chunk); ++i) {
        b += padded(hs[i], ws[i]);
        if (jpeg) {
          cb += (scan_end[i] - scan_begin[i] + 15) & ~(size_t)15;
          for (int k = 0; k < meta[i].ncomp; ++k) {
            const size_t blocks = (size_t)meta[i].comp[k].blocks_w * meta[i].comp[k].blocks_h;
            cc += blocks * 64;
            cp += blocks * 64;
          }
          ct += tabs[i].size();
        }
      }
      slab_bytes = std::max(slab_bytes, b);
      cap_bytes = std::max(cap_bytes, cb); cap_coef = std::max(cap_coef, cc);
      cap_planes = std::max(cap_planes, cp); cap_tabs = std::max(cap_tabs, ct);
    }
    if (!jpeg) {
      for (int i = 0; i < n && all_pinned; ++i) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, imgs[i]) != cudaSuccess) { cudaGetLastError(); all_pinned = false; break; }
        all_pinned = (at.type == cudaMemoryTypeHost);
      }
    } else {
      all_pinned = false;
    }
    const int slots = n == 0 ? 0 : (n_chunks() > 1 ? 2 : 1);
    for (int k = 0; k < slots; ++k) {
      slab[k] = (uint8_t*)e->arena.alloc(slab_bytes);
      if (jpeg) {
        d_bytes[k] = (uint8_t*)e->arena.alloc(cap_bytes + 16);
        d_meta[k] = (JpegImage*)e->arena.alloc((size_t)per_chunk * sizeof(JpegImage));
        d_tabs[k] = (JpegHuff*)e->arena.alloc(std::max<size_t>(1, cap_tabs) * sizeof(JpegHuff));
        d_coef[k] = (short*)e->arena.alloc(cap_coef * 2 + 16);
        d_planes[k] = (uint8_t*)e->arena.alloc(cap_planes + 16);
        d_sched[k] = (int*)e->arena.alloc(JPEG_SCHED_INTS * 4);
      }
    }
  }
  // page-locked staging buffer of slot k with at least `bytes`, free for writing
  uint8_t* staging(int k, size_t bytes) {
    if (e->pinned_cap[k] < bytes) {
      CALD_CUDA_CHECK(cudaEventSynchronize(e->pinned_free[k]));
      if (e->pinned[k]) cudaFreeHost(e->pinned[k]);
      e->pinned[k] = nullptr;
      CALD_CUDA_CHECK(cudaMallocHost((void**)&e->pinned[k], bytes));
      e->pinned_cap[k] = bytes;
    }
    CALD_CUDA_CHECK(cudaEventSynchronize(e->pinned_free[k]));  // the previous DMA out of this staging buffer
    return e->pinned[k];
  }
  // enqueue the copy (and, for files, the decode) of chunk c; no-op if already started or out of range
  void start(int c) {
    if (c != started || c >= n_chunks()) return;
    const int k = c & 1, i0 = c * per_chunk, i1 = std::min(n_images, i0 + per_chunk);
    cudaStream_t cs = e->copy_st;
    if (jpeg) {
      uint8_t* stage = staging(k, host_walk ? cap_coef * 2 + 16 : cap_bytes + 16);
      h_meta[k].clear(); h_tabs[k].clear();
      size_t ob = 0, oc = 0, op = 0, oo = 0;
      int max_w = 0, max_h = 0, max_blocks = 0;
      for (int i = i0; i < i1; ++i) {
        JpegImage im = meta[i];
        const size_t len = scan_end[i] - scan_begin[i];
        if (!host_walk) memcpy(stage + ob, imgs[i] + scan_begin[i], len);
        im.scan_off = (long long)ob; im.scan_len = (long long)len;
        ob += (len + 15) & ~(size_t)15;
        for (int q = 0; q < im.ncomp; ++q) {
          const size_t blocks = (size_t)im.comp[q].blocks_w * im.comp[q].blocks_h;
          im.comp[q].coef_off = (long long)oc; im.comp[q].plane_off = (long long)op;
          oc += blocks * 64; op += blocks * 64;
          max_blocks = std::max(max_blocks, (int)blocks);
        }
        const int tbase = (int)h_tabs[k].size();
        for (int q = 0; q < 2; ++q) {
          if (im.huff_dc[q] >= 0) im.huff_dc[q] += tbase;
          if (im.huff_ac[q] >= 0) im.huff_ac[q] += tbase;
        }
        h_tabs[k].insert(h_tabs[k].end(), tabs[i].begin(), tabs[i].end());
        im.out_off = (long long)oo;
        oo += padded(im.height, im.width);
        max_w = std::max(max_w, im.width); max_h = std::max(max_h, im.height);
        h_meta[k].push_back(im);
      }
      const int nb = i1 - i0;
      cudaEvent_t ta = nullptr, tb = nullptr;
      if (trace_jpeg) { cudaEventCreate(&ta); cudaEventCreate(&tb); cudaEventRecord(ta, cs); }
      CALD_CUDA_CHECK(cudaMemcpyAsync(d_meta[k], h_meta[k].data(), (size_t)nb * sizeof(JpegImage), cudaMemcpyHostToDevice, cs));
      if (host_walk) {
        // walk the scans on a few host threads straight into the page-locked staging buffer (zeroed first: only the
        // non-zero coefficients are written), then one DMA of the chunk's coefficients
        short* hc = reinterpret_cast<short*>(stage);
        const int nthr = std::max(1, std::min(4, nb));
        auto work = [&](int t) {
          for (int j = t; j < nb; j += nthr) {
            const JpegImage& im = h_meta[k][j];
            size_t lo = (size_t)im.comp[0].coef_off, hi = lo;
            for (int q = 0; q < im.ncomp; ++q)
              hi = std::max(hi, (size_t)im.comp[q].coef_off + (size_t)im.comp[q].blocks_w * im.comp[q].blocks_h * 64);
            memset(hc + lo, 0, (hi - lo) * 2);
            jpeg_walk(im, h_tabs[k].data(), imgs[i0 + j] + scan_begin[i0 + j], hc);
          }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(work, t);
        work(0);
        for (std::thread& th : pool) th.join();
        CALD_CUDA_CHECK(cudaMemcpyAsync(d_coef[k], stage, oc * 2, cudaMemcpyHostToDevice, cs));
        CALD_CUDA_CHECK(cudaEventRecord(e->pinned_free[k], cs));
      } else {
        CALD_CUDA_CHECK(cudaMemcpyAsync(d_bytes[k], stage, ob, cudaMemcpyHostToDevice, cs));
        CALD_CUDA_CHECK(cudaEventRecord(e->pinned_free[k], cs));
        if (!h_tabs[k].empty())
          CALD_CUDA_CHECK(cudaMemcpyAsync(d_tabs[k], h_tabs[k].data(), h_tabs[k].size() * sizeof(JpegHuff), cudaMemcpyHostToDevice, cs));
        CALD_CUDA_CHECK(cudaMemsetAsync(d_coef[k], 0, oc * 2, cs));
        CALD_CUDA_CHECK(cudaMemsetAsync(d_sched[k], 0, JPEG_SCHED_INTS * 4, cs));
        jpeg_huffman_kernel<<<JPEG_WALK_BLOCKS, 32, 0, cs>>>(d_meta[k], d_tabs[k], d_bytes[k], d_coef[k], nb, d_sched[k]);
        e->launches += 1;
      }
      jpeg_idct_kernel<<<dim3((max_blocks + 127) / 128, nb * 3), 128, 0, cs>>>(d_meta[k], d_coef[k], d_planes[k]);
      jpeg_rgb_kernel<<<dim3((max_w + 127) / 128, max_h, nb), 128, 0, cs>>>(d_meta[k], d_planes[k], slab[k]);
      CALD_CUDA_CHECK(cudaGetLastError());
      if (trace_jpeg) { cudaEventRecord(tb, cs); trace_ev.push_back({ta, tb}); }
      e->launches += 2;
    } else if (all_pinned) {
      size_t off = 0;
      for (int i = i0; i < i1; ++i) {
        CALD_CUDA_CHECK(cudaMemcpyAsync(slab[k] + off, imgs[i], (size_t)hs[i] * ws[i] * 3, cudaMemcpyHostToDevice, cs));
        off += padded(hs[i], ws[i]);
      }
    } else {
      uint8_t* stage = staging(k, slab_bytes);
      // gather the chunk into the page-locked staging buffer with a few host threads: one core copies ~6 GB/s, and
      // the first chunk of a call has nothing to hide behind (32 images of 3.2 MB = 17 ms single-threaded)
      std::vector<size_t> offs(i1 - i0);
      size_t off = 0;
      for (int i = i0; i < i1; ++i) { offs[i - i0] = off; off += padded(hs[i], ws[i]); }
      const int nthr = std::max(1, std::min(4, i1 - i0));
      auto work = [&](int t) {
        for (int i = i0 + t; i < i1; i += nthr) memcpy(stage + offs[i - i0], imgs[i], (size_t)hs[i] * ws[i] * 3);
      };
      std::vector<std::thread> pool;
      for (int t = 1; t < nthr; ++t) pool.emplace_back(work, t);
      work(0);
      for (std::thread& th : pool) th.join();
      CALD_CUDA_CHECK(cudaMemcpyAsync(slab[k], stage, off, cudaMemcpyHostToDevice, cs));
      CALD_CUDA_CHECK(cudaEventRecord(e->pinned_free[k], cs));
    }
    CALD_CUDA_CHECK(cudaEventRecord(e->upload_done[k], cs));
    started = c + 1;
  }
  // device pointers of chunk c; the compute stream waits for its copy
  DeviceImages get(int c) {
    start(c);
    const int k = c & 1, i0 = c * per_chunk, i1 = std::min(n_images, i0 + per_chunk);
    CALD_CUDA_CHECK(cudaStreamWaitEvent(e->st, e->upload_done[k], 0));
    if (jpeg && trace_jpeg) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, e->st); trace_chunk.push_back(ev); }
    DeviceImages d;
    size_t off = 0;
    for (int i = i0; i < i1; ++i) { d.ptr.push_back(slab[k] + off); off += padded(hs[i], ws[i]); }
    return d;
  }
};

// RetinaNet capacities are fixed; a pool image that exceeds them must fail the call, not be scored differently
void check_overflow(cald_engine* e) {
  if (!e->retina) return;
  int flag = 0;
  CALD_CUDA_CHECK(cudaMemcpyAsync(&flag, e->d_overflow, 4, cudaMemcpyDeviceToHost, e->st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
  if (flag) {
    CALD_CUDA_CHECK(cudaMemsetAsync(e->d_overflow, 0, 4, e->st));
    throw std::runtime_error("RetinaNet: more detections in one image than retina_max_detections (the default, "
                             "300 * num_classes, cannot overflow)");
  }
}

void check_ready(cald_engine* e) {
  if (!e->weights_ready) throw std::runtime_error("cald_load_weights has not been called");
  CALD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
}

}  // namespace

#define API_TRY(e) try {
#define API_CATCH(e)                                  \
  }                                                   \
  catch (const std::exception& ex) {                  \
    (e)->err = ex.what();                             \
    cudaGetLastError();                               \
    return -1;                                        \
  }                                                   \
  return 0;

extern "C" {

int cald_config_default(cald_config* cfg, int arch, int depth, int num_classes, int min_size, int max_size) {
  if (!cfg) return -1;
  memset(cfg, 0, sizeof(*cfg));
  cfg->arch = arch; cfg->depth = depth; cfg->num_classes = num_classes;
  cfg->min_size = min_size; cfg->max_size = max_size;
  cfg->rpn_pre_nms_top_n = 1000; cfg->rpn_post_nms_top_n = 1000; cfg->rpn_nms_thresh = 0.7f;
  cfg->box_score_thresh = 0.05f; cfg->box_nms_thresh = 0.5f; cfg->box_detections_per_img = 100;
  if (arch == CALD_ARCH_RETINANET) cfg->box_detections_per_img = 300;  /* per class, retinanet_cal.py:333,463 */
  /* the reference keeps up to 300 detections per class (retinanet_cal.py:333, 463): 300 * K rows can never overflow */
  cfg->retina_max_detections = std::min(32768, 300 * num_classes);
  cfg->device = 0; cfg->precision = CALD_PREC_F16X3; cfg->conv_impl = CALD_CONV_TCGEN05;
  cfg->max_views_per_pass = 0; cfg->workspace_bytes = 0; cfg->debug = 0;
  return 0;
}

const char* cald_last_error(const cald_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int cald_create(const cald_config* cfg, cald_engine** out) {
  if (!cfg || !out) return -1;
  cald_engine* e = nullptr;
  try {
    if (cfg->arch != CALD_ARCH_FRCNN && cfg->arch != CALD_ARCH_RETINANET) throw std::runtime_error("unknown arch");
    if (cfg->depth != 50 && cfg->depth != 101) throw std::runtime_error("depth must be 50 or 101");
    const bool retina = cfg->arch == CALD_ARCH_RETINANET;
    if (!retina && (cfg->box_detections_per_img > 128 || cfg->rpn_post_nms_top_n > 1000 || cfg->rpn_pre_nms_top_n > 1000))
      throw std::runtime_error("capacity limits: detections <= 128, rpn top-n <= 1000");
    if (retina && (cfg->box_detections_per_img < 1 || cfg->box_detections_per_img > RET_CAND ||
                   cfg->retina_max_detections < 1 || cfg->retina_max_detections > 32768))
      throw std::runtime_error("capacity limits: RetinaNet detections per class <= 4096, per image <= 32768");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
      throw std::runtime_error("no CUDA device: the CALD B200 engine has no CPU fallback");
    CALD_CUDA_CHECK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CALD_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) throw std::runtime_error("this library contains sm_100a code only (Blackwell B200 required)");
    e = new cald_engine();
    e->cfg = *cfg;
    e->split = cfg->precision == CALD_PREC_F16X3;
    e->C = cfg->num_classes;
    e->retina = retina;
    if (retina) {
      // one score row per detection: the "proposal" capacity of the generic view buffers is the detection capacity
      e->ret_per_class = cfg->box_detections_per_img;
      e->det_cap = cfg->retina_max_detections;
      e->cap = e->det_cap;
    } else {
      e->cap = 1000;
      e->det_cap = cfg->box_detections_per_img;
    }
    CALD_CUDA_CHECK(cudaMalloc((void**)&e->d_overflow, 4));
    CALD_CUDA_CHECK(cudaMemset(e->d_overflow, 0, 4));
    CALD_CUDA_CHECK(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    e->conv.num_sms = prop.multiProcessorCount;
    e->conv.impl = cfg->conv_impl == CALD_CONV_SIMT ? CONV_SIMT : CONV_TC;
    e->conv.split = e->split;
    size_t ws = cfg->workspace_bytes;
    if (!ws) {
      size_t fr = 0, tot = 0;
      CALD_CUDA_CHECK(cudaMemGetInfo(&fr, &tot));
      ws = std::min<size_t>(fr / 2, (size_t)64 << 30);
    }
    e->arena.init(ws);
    CALD_CUDA_CHECK(cudaStreamCreateWithFlags(&e->copy_st, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CALD_CUDA_CHECK(cudaEventCreateWithFlags(&e->pinned_free[k], cudaEventDisableTiming));
      CALD_CUDA_CHECK(cudaEventCreateWithFlags(&e->upload_done[k], cudaEventDisableTiming));
    }
    if (cfg->max_views_per_pass > 0) {
      e->views_per_pass = cfg->max_views_per_pass;
    } else {
      // auto: as many views as the arena holds at the detector's largest padded input, up to 32.  Measured on one box
      // (profiles/r02_summary.md): 16 / 64 / 128 views per pass give 128.8 / 129.7 / 130.1 img/s -- pass size no
      // longer matters for the kernels -- while a chunk of 32 images lets a 64-image call overlap half of its host ->
      // device traffic with compute.  Peak arena use per view is ~ARENA_BYTES_PER_PIXEL of the padded input (measured
      // with cald_arena_peak: 290 B; activations of the widest point of the pass + RoI features + per-view result
      // buffers); two chunk-sized image slabs come on top.
      const double per_view = (double)pad32(cfg->min_size) * (double)pad32(cfg->max_size) * ARENA_BYTES_PER_PIXEL +
                              (retina ? (double)e->det_cap * (cfg->num_classes + 16) * 4.0 : 8.0e6);
      const double fit = 0.85 * (double)ws / per_view;
      e->views_per_pass = (int)std::max(4.0, std::min(32.0, std::floor(fit)));
    }
    CALD_CUDA_CHECK(cudaFuncSetAttribute(nms_groups_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_SMEM));
    CALD_CUDA_CHECK(cudaFuncSetAttribute(det_class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         NMS_SMEM + TOPK_MAX * 12));
    CALD_CUDA_CHECK(cudaFuncSetAttribute(rpn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_CAP * 8));
    CALD_CUDA_CHECK(cudaFuncSetAttribute(ret_class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RET_NMS_SMEM));
    // sub-sampling LUT: np.round(np.linspace(0, n-1, 50)).astype(int) (cald_train.py:110-111)
    std::vector<int> lut((size_t)(e->det_cap + 1) * 50, 0);
    for (int n = 41; n <= e->det_cap; ++n) {
      double step = (double)(n - 1) / 49.0;
      for (int i = 0; i < 50; ++i) {
        double y = (i == 49) ? (double)(n - 1) : (double)i * step;
        lut[(size_t)n * 50 + i] = (int)nearbyint(y);
      }
    }
    CALD_CUDA_CHECK(cudaMalloc((void**)&e->d_lut, lut.size() * 4));
    CALD_CUDA_CHECK(cudaMemcpy(e->d_lut, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice));
    *out = e;
    return 0;
  } catch (const std::exception& ex) {
    g_create_err = ex.what();
    delete e;
    return -1;
  }
}

void cald_destroy(cald_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  delete e;
}

int cald_load_weights(cald_engine* e, int n, const char* const* names, const float* const* data, const int* ndim,
                      const int64_t* shapes) {
  API_TRY(e)
  CALD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));  // nothing may still be reading the previous weights
  e->free_weights();
  for (int i = 0; i < n; ++i) {
    HostTensor t;
    size_t cnt = 1;
    for (int d = 0; d < ndim[i]; ++d) { t.shape.push_back(shapes[i * 4 + d]); cnt *= (size_t)shapes[i * 4 + d]; }
    t.v.assign(data[i], data[i] + cnt);
    e->staged[canonical_key(names[i])] = std::move(t);
  }
  finalize_weights(e);
  API_CATCH(e)
}

static int score_impl(cald_engine* e, int n_images, const uint8_t* const* imgs, bool on_device, const int* heights,
                      const int* widths, int n_augs, const cald_aug* aug_list, double bp, const double* rng_uniforms,
                      int n_uniforms, int* uniforms_consumed, const float* const* noise, const int* swap_perms,
                      double* out_consistency, double* out_cls, int scorer = 0,
                      const size_t* file_sizes = nullptr /* imgs are JPEG files */, int* out_heights = nullptr,
                      int* out_widths = nullptr) {
  API_TRY(e)
  check_ready(e);
  e->arena.reset();
  e->last_per_view.clear();
  e->last_ref_counts.clear();
  e->dbg_views.clear();
  e->last_A = n_augs;
  std::vector<cald_aug> augs(aug_list, aug_list + n_augs);
  int n_noise = 0, n_swap = 0;
  for (const cald_aug& a : augs) {
    if (a.kind == CALD_AUG_GAUSS || a.kind == CALD_AUG_SALTPEPPER) n_noise++;
    if (a.kind == CALD_AUG_COLOR_SWAP) n_swap++;
  }
  if (n_noise && !noise) throw std::runtime_error("noise augmentation requested but noise == NULL");
  double* d_u = nullptr;
  int* d_cursor = (int*)e->arena.alloc(4);
  CALD_CUDA_CHECK(cudaMemsetAsync(d_cursor, 0, 4, e->st));
  if (n_uniforms > 0 && rng_uniforms) {
    d_u = (double*)e->arena.alloc((size_t)n_uniforms * 8);
    CALD_CUDA_CHECK(cudaMemcpyAsync(d_u, rng_uniforms, (size_t)n_uniforms * 8, cudaMemcpyHostToDevice, e->st));
  } else {
    n_uniforms = 0;
  }
  const int maxv = e->views_per_pass;
  // A chunk is as many IMAGES as one pass holds views: its reference views run as one full pass and its B * A
  // augmented views as A more (detect_views splits them), so every launch of the chunk works on a full-size batch.
  // (Round 1 used maxv / A images per chunk: the 16-view reference pass ran its launches 14 - 19 % less efficiently
  // than the 64-view augmented pass, 3.4 ms of a 127 ms step.)
  const int Bmax = std::max(1, maxv);
  e->trace("call_start");
  std::unique_ptr<UploadPipe> pipe;
  if (!on_device) pipe.reset(new UploadPipe(e, n_images, imgs, heights, widths, Bmax, file_sizes));
  if (file_sizes) {
    heights = pipe->hs; widths = pipe->ws;   // sizes come from the files' frame headers
    for (int i = 0; i < n_images; ++i) {
      if (out_heights) out_heights[i] = heights[i];
      if (out_widths) out_widths[i] = widths[i];
    }
  }
  // no copy may still be reading the caller's buffers when the call returns, on the error path either
  struct CopyFence { cudaStream_t s; ~CopyFence() { cudaStreamSynchronize(s); } } copy_fence{e->copy_st};
  int chunk = 0;
  for (int pos = 0; pos < n_images; pos += Bmax, ++chunk) {
    const int B = std::min(Bmax, n_images - pos);
    DeviceImages di;
    const uint8_t* const* dptr;
    std::vector<const float*> nz;
    std::vector<float*> nz_owned;
    if (on_device) {
      dptr = imgs + pos;
      for (int i = 0; i < B * n_noise; ++i) nz.push_back(noise[(size_t)pos * n_noise + i]);
    } else {
      di = pipe->get(chunk);
      dptr = di.ptr.data();
      for (int b = 0; b < B; ++b)
        for (int j = 0; j < n_noise; ++j) {
          const float* src = noise[(size_t)(pos + b) * n_noise + j];
          if (!src) { nz.push_back(nullptr); continue; }  // image without reference detections: never read
          size_t bytes = (size_t)heights[pos + b] * widths[pos + b] * 3 * 4;
          float* d = (float*)e->arena.alloc(bytes);
          CALD_CUDA_CHECK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, e->st));
          nz.push_back(d);
          nz_owned.push_back(d);
        }
    }
    // the next chunk's images start travelling once this chunk's work is enqueued, before its results are awaited
    UploadPipe* pp = pipe.get();
    const int next = chunk + 1;
    std::function<void()> prefetch = [pp, next]() { if (pp) pp->start(next); };
    score_chunk(e, B, dptr, heights + pos, widths + pos, augs, bp, d_u, n_uniforms, d_cursor,
                n_noise ? nz.data() : nullptr, swap_perms ? swap_perms + (size_t)pos * n_swap : nullptr,
                out_consistency + pos, out_cls ? out_cls + (size_t)pos * (e->C - 1) : nullptr, scorer, prefetch);
    for (float* d : nz_owned) e->arena.free(d);
  }
  int consumed = 0;
  CALD_CUDA_CHECK(cudaMemcpyAsync(&consumed, d_cursor, 4, cudaMemcpyDeviceToHost, e->st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
  if (uniforms_consumed) *uniforms_consumed = consumed;
  e->trace_dump();
  API_CATCH(e)
}

int cald_score(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
               int n_augs, const cald_aug* augs, double bp, const double* rng_uniforms, int n_uniforms,
               int* uniforms_consumed, const float* const* noise, const int* swap_perms, double* out_consistency,
               double* out_cls) {
  return score_impl(e, n_images, images, false, heights, widths, n_augs, augs, bp, rng_uniforms, n_uniforms,
                    uniforms_consumed, noise, swap_perms, out_consistency, out_cls);
}
int cald_score_device(cald_engine* e, int n_images, const uint8_t* const* d_images, const int* heights,
                      const int* widths, int n_augs, const cald_aug* augs, double bp, const double* rng_uniforms,
                      int n_uniforms, int* uniforms_consumed, const float* const* d_noise, const int* swap_perms,
                      double* out_consistency, double* out_cls) {
  return score_impl(e, n_images, d_images, true, heights, widths, n_augs, augs, bp, rng_uniforms, n_uniforms,
                    uniforms_consumed, d_noise, swap_perms, out_consistency, out_cls);
}

int cald_score_jpeg(cald_engine* e, int n_files, const uint8_t* const* files, const size_t* file_sizes, int n_augs,
                    const cald_aug* augs, double bp, const double* rng_uniforms, int n_uniforms, int* uniforms_consumed,
                    const int* swap_perms, double* out_consistency, double* out_cls, int* out_heights, int* out_widths) {
  if (!file_sizes) { if (e) e->err = "cald_score_jpeg: file_sizes == NULL"; return -1; }
  for (int i = 0; i < n_augs; ++i)
    if (augs[i].kind == CALD_AUG_GAUSS || augs[i].kind == CALD_AUG_SALTPEPPER) {
      e->err = "cald_score_jpeg: noise views need the image size before the call; decode with cald_jpeg_decode first";
      return -1;
    }
  return score_impl(e, n_files, files, false, nullptr, nullptr, n_augs, augs, bp, rng_uniforms, n_uniforms,
                    uniforms_consumed, nullptr, swap_perms, out_consistency, out_cls, 0, file_sizes, out_heights,
                    out_widths);
}

int cald_jpeg_info(const uint8_t* file, size_t size, int* height, int* width, int* components) {
  try {
    JpegImage im;
    std::vector<JpegHuff> t;
    size_t a, b;
    jpeg_parse(file, size, im, t, a, b);
    if (height) *height = im.height;
    if (width) *width = im.width;
    if (components) *components = im.ncomp;
    return 0;
  } catch (const std::exception& ex) {
    g_create_err = ex.what();
    return -1;
  }
}

int cald_jpeg_coefficients(const uint8_t* file, size_t size, int16_t* out, size_t capacity, size_t* n_coef, int* blocks_w,
                           int* blocks_h) {
  try {
    JpegImage im;
    std::vector<JpegHuff> t;
    size_t a, b;
    jpeg_parse(file, size, im, t, a, b);
    size_t total = 0;
    for (int q = 0; q < im.ncomp; ++q) {
      im.comp[q].coef_off = (long long)total;
      total += (size_t)im.comp[q].blocks_w * im.comp[q].blocks_h * 64;
      if (blocks_w) blocks_w[q] = im.comp[q].blocks_w;
      if (blocks_h) blocks_h[q] = im.comp[q].blocks_h;
    }
    if (n_coef) *n_coef = total;
    if (!out || capacity < total) throw std::runtime_error("cald_jpeg_coefficients: output buffer too small");
    im.scan_off = 0; im.scan_len = (long long)(b - a);
    memset(out, 0, total * sizeof(int16_t));
    jpeg_walk(im, t.data(), file + a, out);     // the function UploadPipe's host threads (and the device kernel) run
    return 0;
  } catch (const std::exception& ex) {
    g_create_err = ex.what();
    return -1;
  }
}

int cald_jpeg_decode(cald_engine* e, int n_files, const uint8_t* const* files, const size_t* file_sizes,
                     uint8_t* const* out_images) {
  API_TRY(e)
  CALD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
  e->arena.reset();
  const int per = 16;
  UploadPipe pipe(e, n_files, files, nullptr, nullptr, per, file_sizes);
  struct CopyFence { cudaStream_t s; ~CopyFence() { cudaStreamSynchronize(s); } } copy_fence{e->copy_st};
  int chunk = 0;
  for (int pos = 0; pos < n_files; pos += per, ++chunk) {
    const int B = std::min(per, n_files - pos);
    DeviceImages di = pipe.get(chunk);
    pipe.start(chunk + 1);
    for (int b = 0; b < B; ++b)
      CALD_CUDA_CHECK(cudaMemcpyAsync(out_images[pos + b], di.ptr[b], (size_t)pipe.hs[pos + b] * pipe.ws[pos + b] * 3,
                                      cudaMemcpyDeviceToHost, e->st));
    CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
  }
  API_CATCH(e)
}

int cald_select(cald_engine* e, int n, const double* uncertainty, const double* cls, int c1, const double* mean_hist,
                int budget, int n_cand, int uniform, int* out_positions, int out_capacity, int* n_picked) {
  API_TRY(e)
  CALD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
  if (n <= 0 || c1 <= 0 || budget <= 0) throw std::runtime_error("cald_select: empty input");
  const int m = std::min(n, n_cand);
  if (m > SEL_MAX_CAND) throw std::runtime_error("cald_select: more than 8192 candidates (int(mr * budget))");
  if (out_capacity < m) throw std::runtime_error("cald_select: out_positions must hold n_cand entries");
  static std::once_flag once;
  std::call_once(once, [] {
    CALD_CUDA_CHECK(cudaFuncSetAttribute(select_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SEL_MAX_CAND * 12));
  });
  e->arena.reset();
  Arena& ar = e->arena;
  cudaStream_t st = e->st;
  double* d_unc = (double*)ar.alloc((size_t)n * 8);
  double* d_cls = (double*)ar.alloc((size_t)n * c1 * 8);
  double* d_hist = (double*)ar.alloc((size_t)c1 * 8);
  int* d_cand = (int*)ar.alloc((size_t)m * 4);
  double* d_js = (double*)ar.alloc((size_t)m * 8);
  int* d_zero = (int*)ar.alloc((size_t)m * 4);
  int* d_pick = (int*)ar.alloc((size_t)m * 4 + 4);
  CALD_CUDA_CHECK(cudaMemcpyAsync(d_unc, uncertainty, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CALD_CUDA_CHECK(cudaMemcpyAsync(d_cls, cls, (size_t)n * c1 * 8, cudaMemcpyHostToDevice, st));
  CALD_CUDA_CHECK(cudaMemcpyAsync(d_hist, mean_hist, (size_t)c1 * 8, cudaMemcpyHostToDevice, st));
  select_candidates_kernel<<<1, 1024, SEL_MAX_CAND * 12, st>>>(d_unc, n, m, d_cand);
  select_js_kernel<<<(m * 32 + 255) / 256, 256, 0, st>>>(d_cls, d_cand, m, c1, d_hist, uniform, d_js, d_zero);
  select_pick_kernel<<<1, 1024, 0, st>>>(d_js, d_zero, m, budget, uniform, d_pick, d_pick + m);
  CALD_CUDA_CHECK(cudaGetLastError());
  e->launches += 3;
  std::vector<int> h_cand(m), h_pick(m + 1);
  CALD_CUDA_CHECK(cudaMemcpyAsync(h_cand.data(), d_cand, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
  CALD_CUDA_CHECK(cudaMemcpyAsync(h_pick.data(), d_pick, (size_t)(m + 1) * 4, cudaMemcpyDeviceToHost, st));
  CALD_CUDA_CHECK(cudaStreamSynchronize(st));
  const int np = std::min(h_pick[m], m);
  for (int i = 0; i < np; ++i) out_positions[i] = h_cand[h_pick[i]];   // subset[arg][picked]: pool positions
  if (n_picked) *n_picked = np;
  API_CATCH(e)
}

int cald_score_lsc(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
                   const float* const* noise, double* out_stability) {
  // ls_c_train.get_uncertainty: 6 Gaussian-noise views with std 8, 16, ..., 48 (ls_c_train.py:127-129)
  cald_aug views[6];
  for (int i = 0; i < 6; ++i) { views[i].kind = CALD_AUG_GAUSS; views[i].param = 8.0 * (i + 1); }
  return score_impl(e, n_images, images, false, heights, widths, 6, views, 0.0, nullptr, 0, nullptr, noise, nullptr,
                    out_stability, nullptr, 1);
}

int cald_score_ltc(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
                   double* out_uncertainty) {
  API_TRY(e)
  check_ready(e);
  if (e->retina) throw std::runtime_error("LT/C needs the proposals of a Faster R-CNN ('props', frcnn_la.py:131-141)");
  e->arena.reset();
  const int dc = e->det_cap;
  const int maxv = e->views_per_pass;
  UploadPipe pipe(e, n_images, images, heights, widths, maxv);
  struct CopyFence { cudaStream_t s; ~CopyFence() { cudaStreamSynchronize(s); } } copy_fence{e->copy_st};
  int chunk = 0;
  for (int pos = 0; pos < n_images; pos += maxv, ++chunk) {
    const int B = std::min(maxv, n_images - pos);
    DeviceImages di = pipe.get(chunk);
    std::vector<HostView> hv(B);
    for (int b = 0; b < B; ++b) hv[b] = HostView{di.ptr[b], heights[pos + b], widths[pos + b], 0, -1};
    ViewSet vs = alloc_viewset(e, B);
    detect_views(e, hv, nullptr, vs);
    float* d_out = (float*)e->arena.alloc((size_t)B * 4);
    ltc_kernel<<<B, 128, 0, e->st>>>(vs.det, dc, d_out);
    CALD_CUDA_CHECK(cudaGetLastError());
    KLAUNCH(e);
    std::vector<float> h(B);
    pipe.start(chunk + 1);   // before the (host-blocking) copy into pageable memory
    CALD_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_out, (size_t)B * 4, cudaMemcpyDeviceToHost, e->st));
    CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
    for (int b = 0; b < B; ++b) out_uncertainty[pos + b] = (double)h[b];
    e->arena.free(d_out);
    free_viewset(e, vs);
  }
  API_CATCH(e)
}

int cald_detect(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
                int* counts, float* boxes, float* scores, int64_t* labels, float* props, float* prob_max,
                float* scores_cls) {
  API_TRY(e)
  check_ready(e);
  e->arena.reset();
  const int dc = e->det_cap, C = e->C;
  const int maxv = e->views_per_pass;
  UploadPipe pipe(e, n_images, images, heights, widths, maxv);
  struct CopyFence { cudaStream_t s; ~CopyFence() { cudaStreamSynchronize(s); } } copy_fence{e->copy_st};
  int chunk = 0;
  for (int pos = 0; pos < n_images; pos += maxv, ++chunk) {
    const int B = std::min(maxv, n_images - pos);
    DeviceImages di = pipe.get(chunk);
    std::vector<HostView> hv(B);
    for (int b = 0; b < B; ++b) hv[b] = HostView{di.ptr[b], heights[pos + b], widths[pos + b], 0, -1};
    ViewSet vs = alloc_viewset(e, B);
    detect_views(e, hv, nullptr, vs);
    std::vector<int> h_count(B), h_labels((size_t)B * dc), h_pidx((size_t)B * dc);
    std::vector<float> h_scores_all;
    cudaStream_t st = e->st;
    pipe.start(chunk + 1);   // before the (host-blocking) copies into pageable memory
    CALD_CUDA_CHECK(cudaMemcpyAsync(h_count.data(), vs.det.count, B * 4, cudaMemcpyDeviceToHost, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(h_labels.data(), vs.det.labels, (size_t)B * dc * 4, cudaMemcpyDeviceToHost, st));
    CALD_CUDA_CHECK(cudaMemcpyAsync(h_pidx.data(), vs.det.prop_idx, (size_t)B * dc * 4, cudaMemcpyDeviceToHost, st));
    if (boxes) CALD_CUDA_CHECK(cudaMemcpyAsync(boxes + (size_t)pos * dc * 4, vs.det.boxes, (size_t)B * dc * 16, cudaMemcpyDeviceToHost, st));
    if (props) CALD_CUDA_CHECK(cudaMemcpyAsync(props + (size_t)pos * dc * 4, vs.det.props, (size_t)B * dc * 16, cudaMemcpyDeviceToHost, st));
    if (scores) CALD_CUDA_CHECK(cudaMemcpyAsync(scores + (size_t)pos * dc, vs.det.scores, (size_t)B * dc * 4, cudaMemcpyDeviceToHost, st));
    if (prob_max) CALD_CUDA_CHECK(cudaMemcpyAsync(prob_max + (size_t)pos * dc, vs.det.prob_max, (size_t)B * dc * 4, cudaMemcpyDeviceToHost, st));
    if (scores_cls) {
      h_scores_all.resize((size_t)B * e->cap * C);
      CALD_CUDA_CHECK(cudaMemcpyAsync(h_scores_all.data(), vs.scores, h_scores_all.size() * 4, cudaMemcpyDeviceToHost, st));
    }
    CALD_CUDA_CHECK(cudaStreamSynchronize(st));
    check_overflow(e);
    for (int b = 0; b < B; ++b) {
      if (counts) counts[pos + b] = h_count[b];
      for (int i = 0; i < dc; ++i) {
        if (labels) labels[(size_t)(pos + b) * dc + i] = h_labels[(size_t)b * dc + i];
        if (scores_cls) {
          float* dst = scores_cls + ((size_t)(pos + b) * dc + i) * C;
          if (i < h_count[b]) memcpy(dst, &h_scores_all[((size_t)b * e->cap + h_pidx[(size_t)b * dc + i]) * C], (size_t)C * 4);
          else memset(dst, 0, (size_t)C * 4);
        }
      }
    }
    free_viewset(e, vs);
  }
  API_CATCH(e)
}

int cald_last_ref_counts(cald_engine* e, int* out, int capacity) {
  if (!e || !out) return -1;
  int n = (int)e->last_ref_counts.size();
  if (n > capacity) n = capacity;
  memcpy(out, e->last_ref_counts.data(), (size_t)n * 4);
  return n;
}

int cald_last_per_view(cald_engine* e, float* out, int capacity) {
  if (!e || !out) return -1;
  int n = (int)e->last_per_view.size();
  if (n > capacity) n = capacity;
  memcpy(out, e->last_per_view.data(), (size_t)n * 4);
  return n;
}

long long cald_debug_views(cald_engine* e, int* counts, int counts_capacity, float* boxes, float* scores, int* labels,
                           float* prob_max, long long rows_capacity) {
  if (!e) return -1;
  const cald_engine::DbgViews& d = e->dbg_views;
  const long long rows = (long long)d.scores.size();
  if (counts) memcpy(counts, d.counts.data(), (size_t)std::min<long long>(counts_capacity, (long long)d.counts.size()) * 4);
  const size_t n = (size_t)std::min(rows, rows_capacity);
  if (boxes) memcpy(boxes, d.boxes.data(), n * 16);
  if (scores) memcpy(scores, d.scores.data(), n * 4);
  if (labels) memcpy(labels, d.labels.data(), n * 4);
  if (prob_max) memcpy(prob_max, d.prob_max.data(), n * 4);
  return rows;
}

long long cald_arena_peak(cald_engine* e) { return e ? (long long)e->arena.peak : -1; }

int cald_views_per_pass(cald_engine* e) { return e ? e->views_per_pass : -1; }

long long cald_debug_fetch(cald_engine* e, const char* name, float* buf, long long capacity) {
  if (!e) return -1;
  auto it = e->dbg.find(name);
  if (it == e->dbg.end()) { e->err = std::string("no debug tensor named ") + name; return -1; }
  long long n = (long long)it->second.size();
  if (buf) memcpy(buf, it->second.data(), (size_t)std::min(n, capacity) * 4);
  return n;
}

int cald_profile(cald_engine* e, int enable) {
  if (!e) return -1;
  e->conv.profiling = enable != 0;
  e->conv.ev_used = 0; e->conv.prof_flops = 0; e->conv.prof_launches = 0;
  return 0;
}

int cald_profile_read(cald_engine* e, double* conv_ms, long long* conv_launches, double* conv_flops) {
  API_TRY(e)
  CALD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
  CALD_CUDA_CHECK(cudaStreamSynchronize(e->st));
  double fl = 0;
  long long n = 0;
  double ms = e->conv.drain_profile(&fl, &n);
  if (conv_ms) *conv_ms = ms;
  if (conv_launches) *conv_launches = n;
  if (conv_flops) *conv_flops = fl;
  API_CATCH(e)
}

long long cald_profile_layers(cald_engine* e, char* buf, long long capacity) {
  if (!e) return -1;
  std::string s = "layer\tcount\tms\tgflop\tmbyte\n";
  char line[256];
  for (auto& kv : e->conv.layer_agg) {
    snprintf(line, sizeof(line), "%s\t%lld\t%.4f\t%.3f\t%.3f\n", kv.first.c_str(), kv.second.count, kv.second.ms,
             kv.second.flops / 1e9, kv.second.bytes / 1e6);
    s += line;
  }
  if (buf && capacity > 0) {
    size_t n = std::min<size_t>(s.size(), (size_t)capacity - 1);
    memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (long long)s.size() + 1;
}

int cald_event_record(cald_engine* e, int slot) {
  API_TRY(e)
  if (slot < 0 || slot >= 8) throw std::runtime_error("event slot out of range");
  CALD_CUDA_CHECK(cudaSetDevice(e->cfg.device));
  if (!e->user_ev[slot]) CALD_CUDA_CHECK(cudaEventCreate(&e->user_ev[slot]));
  CALD_CUDA_CHECK(cudaEventRecord(e->user_ev[slot], e->st));
  API_CATCH(e)
}

int cald_event_elapsed_ms(cald_engine* e, int slot_a, int slot_b, float* ms) {
  API_TRY(e)
  CALD_CUDA_CHECK(cudaEventSynchronize(e->user_ev[slot_b]));
  CALD_CUDA_CHECK(cudaEventElapsedTime(ms, e->user_ev[slot_a], e->user_ev[slot_b]));
  API_CATCH(e)
}

int cald_counters(cald_engine* e, long long* kernel_launches, double* conv_flops) {
  if (!e) return -1;
  if (kernel_launches) *kernel_launches = e->launches;
  if (conv_flops) *conv_flops = e->conv.flops;
  return 0;
}

}  // extern "C"
