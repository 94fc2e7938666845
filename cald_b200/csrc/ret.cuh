// RetinaNet post-processing on device (detection/retinanet_cal.py:402-490, anchors 347-351 and
// tv:models/detection/anchor_utils.py:58-133, decode tv:models/detection/_utils.py:183-224,
// box rescale tv:models/detection/transform.py:306-319).
//
// The reference loops over the K classes in Python (K host syncs, K NMS launches).  Here the whole stage is three
// launches with device-side counts:
//   ret_candidates_kernel : sigmoid + score threshold over all [anchors][K] logits -> per-(view, class) candidate keys
//   ret_class_nms_kernel  : one CTA per (view, class): sort by score, decode + clip, drop small boxes, greedy NMS,
//                           first `per_class` survivors
//   ret_gather_kernel     : concatenate the classes in class order (the reference's torch.cat order), rescale the
//                           boxes to the original image, emit score rows / prob_max per detection
// fp32 arithmetic mirrors the ATen CPU ops one to one (this TU is built with -fmad=false).
#pragma once
#include "det.cuh"

namespace cald {

constexpr int RET_LEVELS = 5;
constexpr int RET_A = 9;            // anchors per location (3 scales x 3 aspect ratios)
constexpr int RET_CAND = 4096;      // candidates per (view, class) that enter the NMS

struct RetLevel {
  const float* cls;   // [V][h][w][ld_cls] fp32: channel a*K + k
  const float* reg;   // [V][h][w][ld_reg] fp32: channel a*4 + j
  int h, w;
  int ld_cls, ld_reg;
  int stride_h, stride_w;
  float base[RET_A][4];
  int n;              // h*w*9
  int off;            // offset in the concatenated anchor index space
};
struct RetLevels {
  RetLevel lv[RET_LEVELS];
  int total;
  int K;
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ int ret_level_of(const RetLevels& L, int idx) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < RET_LEVELS; ++i) if (idx >= L.lv[i].off) l = i;
  return l;
}

// ---------------------------------------------------------------- candidates: score > thresh
// grid = (ceil(total_pixels / PIX_PER_BLOCK), V).  Each block walks whole pixels (9K contiguous logits each).
// keys[v][c][slot] = desc(score) << 32 | global anchor index; counts may exceed RET_CAND (the list then holds an
// arbitrary subset and the consumer falls back to its windowed scan).
__global__ void ret_candidates_kernel(RetLevels L, float thresh, unsigned long long* __restrict__ keys,
                                      int* __restrict__ counts) {
  const int v = blockIdx.y;
  const int K = L.K, AK = RET_A * L.K;
  const int total_pix = L.total / RET_A;
  for (int pix = blockIdx.x; pix < total_pix; pix += gridDim.x) {
    const int l = ret_level_of(L, pix * RET_A);
    const RetLevel& lv = L.lv[l];
    const int lp = pix - lv.off / RET_A;
    const float* row = lv.cls + ((long long)v * lv.h * lv.w + lp) * lv.ld_cls;
    for (int e = threadIdx.x; e < AK; e += blockDim.x) {
      const float s = sigmoidf_ref(row[e]);
      if (s > thresh) {
        const int a = e / K, c = e - a * K;
        const int slot = atomicAdd(&counts[v * K + c], 1);
        // a class that overflows its list is re-scanned in windows by ret_class_nms_kernel (its count says so)
        if (slot < RET_CAND)
          keys[((long long)v * K + c) * RET_CAND + slot] =
              ((unsigned long long)desc_key(s) << 32) | (unsigned)(lv.off + lp * RET_A + a);
      }
    }
  }
}

__device__ __forceinline__ void ret_anchor(const RetLevels& L, int idx, float* a /*x1,y1,x2,y2*/, int& l, int& lp,
                                           int& an) {
  l = ret_level_of(L, idx);
  const RetLevel& lv = L.lv[l];
  const int r = idx - lv.off;
  lp = r / RET_A;
  an = r - lp * RET_A;
  const int x = lp % lv.w, y = lp / lv.w;
  const float sx = (float)(x * lv.stride_w), sy = (float)(y * lv.stride_h);
  a[0] = sx + lv.base[an][0]; a[1] = sy + lv.base[an][1];
  a[2] = sx + lv.base[an][2]; a[3] = sy + lv.base[an][3];
}

// ---------------------------------------------------------------- per-class NMS
// grid = (K, V), block = 1024.  dynamic smem: keys u64[RET_CAND] | boxes float4[RET_CAND] | anchor int[RET_CAND] |
// suppressed u8[RET_CAND].
//
// A class with at most RET_CAND candidates above the score threshold (every class of a trained detector) takes the
// candidate list ret_candidates_kernel wrote.  A class with MORE candidates -- the reference has no limit: a lightly
// trained RetinaNet at the 0.05 threshold can put tens of thousands of anchors of one class above it
// (retinanet_cal.py:436-463) -- is processed exactly all the same, in WINDOWS of the next RET_CAND best keys: a
// radix select over the class's logits finds the window, the window's boxes are first suppressed by the survivors of
// the earlier windows and then run through the greedy scan.  Greedy NMS only ever compares a box with better-scored
// survivors, so the windowed form keeps exactly the boxes torchvision's nms keeps; it stops at per_class survivors
// like the reference's [:detections_per_img] cut.
constexpr int RET_NMS_SMEM = RET_CAND * 8 + RET_CAND * 16 + RET_CAND * 4 + RET_CAND;
constexpr int RET_SEL_BITS = 11;
struct RetKept {                // per (view, class), capacity per_class
  float4* boxes;                // [V][K][per_class]  (detector-input coordinates)
  float* scores;
  int* anchor;
  int* count;                   // [V][K]
};

// key of anchor `idx` for class c of view v, or 0 when its score is not above the threshold (0 is never a valid key:
// desc_key(s) > 0 for every finite s)
__device__ __forceinline__ unsigned long long ret_class_key(const RetLevels& L, int v, int c, int idx, float thresh) {
  const int l = ret_level_of(L, idx);
  const RetLevel& lv = L.lv[l];
  const int r = idx - lv.off;
  const int lp = r / RET_A, an = r - lp * RET_A;
  const float s = sigmoidf_ref(lv.cls[((long long)v * lv.h * lv.w + lp) * lv.ld_cls + an * L.K + c]);
  return s > thresh ? (((unsigned long long)desc_key(s) << 32) | (unsigned)idx) : 0ull;
}

// The (at most RET_CAND) smallest keys greater than `prev` of class c -> skey[0..n); returns n (block-uniform).
// MSB-first radix select of the RET_CAND-th smallest key (keys are distinct), then one gathering sweep.
__device__ int ret_select_window(const RetLevels& L, int v, int c, float thresh, unsigned long long prev,
                                 unsigned long long* skey, unsigned* hist /* smem [1 << RET_SEL_BITS] */) {
  __shared__ unsigned long long s_prefix, s_mask;
  __shared__ int s_need, s_all, s_n;
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = RET_CAND; s_all = 0; s_n = 0; }
  __syncthreads();
  for (int shift = 64 - RET_SEL_BITS; !s_all; shift -= RET_SEL_BITS) {
    const int bits = shift >= 0 ? RET_SEL_BITS : RET_SEL_BITS + shift;
    const int sh = shift >= 0 ? shift : 0;
    const unsigned long long dmask = (1ull << bits) - 1;
    for (int i = tid; i < (1 << RET_SEL_BITS); i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix, mask = s_mask;
    for (int idx = tid; idx < L.total; idx += blockDim.x) {
      const unsigned long long key = ret_class_key(L, v, c, idx, thresh);
      if (key > prev && (key & mask) == prefix) atomicAdd(&hist[(unsigned)((key >> sh) & dmask)], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, need = s_need, digit = -1;
      for (int b = 0; b <= (int)dmask; ++b) {
        if (cum + (int)hist[b] >= need) { digit = b; break; }
        cum += (int)hist[b];
      }
      if (digit < 0) {
        s_all = 1;  // fewer than `need` keys remain under this prefix: the window takes every key > prev
      } else {
        s_need = need - cum;
        s_prefix = prefix | ((unsigned long long)digit << sh);
        s_mask = mask | (dmask << sh);
      }
    }
    __syncthreads();
    if (sh == 0) break;
  }
  // s_all: everything > prev (only possible on the first digit, where the prefix is still empty); else keys <= s_prefix
  const unsigned long long hi = s_all ? ~0ull : s_prefix;
  for (int idx = tid; idx < L.total; idx += blockDim.x) {
    const unsigned long long key = ret_class_key(L, v, c, idx, thresh);
    if (key > prev && key <= hi) {
      const int slot = atomicAdd(&s_n, 1);
      if (slot < RET_CAND) skey[slot] = key;
    }
  }
  __syncthreads();
  return s_n < RET_CAND ? s_n : RET_CAND;
}

__global__ void __launch_bounds__(1024) ret_class_nms_kernel(RetLevels L, const unsigned long long* __restrict__ keys,
                                                             const int* __restrict__ counts,
                                                             const int* __restrict__ image_hw, float thresh,
                                                             float min_size, double nms_thresh, int per_class,
                                                             RetKept out) {
  extern __shared__ __align__(16) unsigned char rsm[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(rsm);
  float4* sbox = reinterpret_cast<float4*>(rsm + RET_CAND * 8);
  int* sanc = reinterpret_cast<int*>(rsm + RET_CAND * 8 + RET_CAND * 16);
  unsigned char* ssup = rsm + RET_CAND * 8 + RET_CAND * 16 + RET_CAND * 4;
  __shared__ int s_m, s_tot;
  const int c = blockIdx.x, v = blockIdx.y, K = L.K;
  const int tid = threadIdx.x;
  const int n_total = counts[v * K + c];
  const bool windowed = n_total > RET_CAND;
  const long long obase = ((long long)v * K + c) * per_class;
  if (n_total == 0) {
    if (tid == 0) out.count[v * K + c] = 0;
    return;
  }
  const float img_h = (float)image_hw[v * 2], img_w = (float)image_hw[v * 2 + 1];
  int kept = 0;
  unsigned long long prev = 0;
  for (;;) {
    int n;
    if (windowed) {
      // sbox is free between windows: its first 8 KB serve as the select histogram
      n = ret_select_window(L, v, c, thresh, prev, skey, reinterpret_cast<unsigned*>(sbox));
      if (n == 0) break;
    } else {
      n = n_total;
      const unsigned long long* gk = keys + ((long long)v * K + c) * RET_CAND;
      for (int i = tid; i < n; i += blockDim.x) skey[i] = gk[i];
    }
    int P = 64;
    while (P < n) P <<= 1;
    for (int i = n + tid; i < P; i += blockDim.x) skey[i] = ~0ull;
    __syncthreads();
    block_bitonic_sort(skey, P);   // score descending, anchor index ascending (= torch's stable order)
    // ---- decode + clip; small boxes are dropped BEFORE the NMS (remove_small_boxes, retinanet_cal.py:453)
    if (tid == 0) s_m = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
      const int i = base + tid;
      int ok = 0;
      float b[4] = {0, 0, 0, 0};
      if (i < n) {
        const int idx = (int)(skey[i] & 0xffffffffu);
        float a[4];
        int l, lp, an;
        ret_anchor(L, idx, a, l, lp, an);
        const RetLevel& lv = L.lv[l];
        const float* d = lv.reg + ((long long)v * lv.h * lv.w + lp) * lv.ld_reg + an * 4;
        float dd[4] = {d[0], d[1], d[2], d[3]};
        decode_box(dd, a[0], a[1], a[2], a[3], 1.f, 1.f, 1.f, 1.f, b);
        clip_box(b, img_h, img_w);
        ok = ((b[2] - b[0]) >= min_size) && ((b[3] - b[1]) >= min_size);
      }
      const int pos = block_excl_scan_1024(ok, &s_tot);
      const int m0 = s_m, tot = s_tot;
      if (ok) {
        sbox[m0 + pos] = make_float4(b[0], b[1], b[2], b[3]);
        sanc[m0 + pos] = i;  // position in the sorted key list (score + anchor recoverable)
      }
      __syncthreads();
      if (tid == 0) s_m = m0 + tot;
      __syncthreads();
    }
    const int m = s_m;
    // ---- boxes of this window that a survivor of an earlier window suppresses (same IoU arithmetic as below)
    for (int j = tid; j < m; j += blockDim.x) {
      unsigned char sup = 0;
      const float4 bj = sbox[j];
      const float aj = (bj.z - bj.x) * (bj.w - bj.y);
      for (int i = 0; i < kept && !sup; ++i) {
        const float4 bi = out.boxes[obase + i];
        const float ai = (bi.z - bi.x) * (bi.w - bi.y);
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
        const float inter = w * h;
        const float ovr = inter / (ai + aj - inter);
        if ((double)ovr > nms_thresh) sup = 1;
      }
      ssup[j] = sup;
    }
    __syncthreads();
    // ---- greedy NMS in score order (torchvision CPU kernel arithmetic), stop after per_class survivors
    for (int i = 0; i < m && kept < per_class; ++i) {
      if (ssup[i]) continue;                      // uniform: every thread reads the same flag
      const float4 bi = sbox[i];
      if (tid == 0) {
        const unsigned long long key = skey[sanc[i]];
        out.boxes[obase + kept] = bi;
        out.scores[obase + kept] = key_to_float((uint32_t)(key >> 32));
        out.anchor[obase + kept] = (int)(key & 0xffffffffu);
      }
      kept++;
      const float ai = (bi.z - bi.x) * (bi.w - bi.y);
      for (int j = i + 1 + tid; j < m; j += blockDim.x) {
        if (ssup[j]) continue;
        const float4 bj = sbox[j];
        const float aj = (bj.z - bj.x) * (bj.w - bj.y);
        const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
        const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
        const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
        const float inter = w * h;
        const float ovr = inter / (ai + aj - inter);
        if ((double)ovr > nms_thresh) ssup[j] = 1;
      }
      __syncthreads();
    }
    if (!windowed || kept >= per_class || n < RET_CAND) break;
    prev = skey[n - 1];   // the window's worst key: the next window starts behind it
    __threadfence_block();
    __syncthreads();      // survivors written by thread 0 must be visible to the next window's suppression pass
  }
  if (tid == 0) out.count[v * K + c] = kept;
}

// ---------------------------------------------------------------- concatenate classes -> detections of the view
// grid = V, block = 1024.  scores_cls rows [V][det_cap][K] = sigmoid of the anchor's K logits (retinanet_cal.py:444,
// "scores_all_class"); prob_max = row max (445); labels are 0-based (414); boxes are rescaled to the original image
// with fp32 ratios (tv:transform.py:306-319).
__global__ void __launch_bounds__(1024) ret_gather_kernel(RetLevels L, RetKept kept, int per_class,
                                                          const float* __restrict__ ratio_hw, int det_cap,
                                                          DetOut out, float* __restrict__ scores_cls,
                                                          int* __restrict__ overflow) {
  extern __shared__ int s_off[];  // [K + 1]
  const int v = blockIdx.x, K = L.K, tid = threadIdx.x;
  if (tid == 0) {
    int s = 0;
    for (int c = 0; c < K; ++c) { s_off[c] = s; s += kept.count[v * K + c]; }
    s_off[K] = s;
    if (s > det_cap) atomicExch(overflow, 2);
    out.count[v] = s < det_cap ? s : det_cap;
  }
  __syncthreads();
  const int total = s_off[K] < det_cap ? s_off[K] : det_cap;
  const float rh = ratio_hw[v * 2], rw = ratio_hw[v * 2 + 1];
  const int warp = tid >> 5, lane = tid & 31;
  for (int d = warp; d < total; d += 32) {
    int lo = 0, hi = K;  // class of detection d: last c with s_off[c] <= d
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] <= d) lo = mid; else hi = mid;
    }
    const int c = lo, j = d - s_off[c];
    const long long src = ((long long)v * K + c) * per_class + j;
    const int idx = kept.anchor[src];
    float a[4];
    int l, lp, an;
    ret_anchor(L, idx, a, l, lp, an);
    const RetLevel& lv = L.lv[l];
    const float* row = lv.cls + ((long long)v * lv.h * lv.w + lp) * lv.ld_cls + an * K;
    float* orow = scores_cls + ((long long)v * det_cap + d) * K;
    float pm = -INFINITY;
    for (int k = lane; k < K; k += 32) {
      const float s = sigmoidf_ref(row[k]);
      orow[k] = s;
      pm = fmaxf(pm, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, o));
    if (lane == 0) {
      const long long o = (long long)v * det_cap + d;
      const float4 b = kept.boxes[src];
      out.boxes[o] = make_float4(b.x * rw, b.y * rh, b.z * rw, b.w * rh);
      out.props[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      out.scores[o] = kept.scores[src];
      out.prob_max[o] = pm;
      out.labels[o] = c;
      out.prop_idx[o] = d;
    }
  }
}

// relu on a split-pl16 tensor (P7's input is relu(P6), tv:ops/feature_pyramid_network.py:247)
__global__ void relu_split_kernel(const pl16* __restrict__ ihi, const pl16* __restrict__ ilo, pl16* __restrict__ ohi,
                                  pl16* __restrict__ olo, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = ilo ? join_pl(ihi[i], ilo[i]) : pl16_to_float(ihi[i]);
  v = fmaxf(v, 0.f);
  pl16 h, l;
  split_pl(v, h, l);
  ohi[i] = h;
  if (olo) olo[i] = l;
}

}  // namespace cald
