// CALD-specific stages (cald_train.py:101-228, cald/cald_helper.py): reference-box sub-sampling, class-max
// vectors, augmentation box transforms, the cutout accept/reject loop, and the paired-prediction reduction
// (pairwise IoU match + Jensen-Shannon divergence -> one consistency scalar per (image, augmentation)).
// fp32 arithmetic mirrors the torch / scipy op order of the reference (this TU is built with -fmad=false).
#pragma once
#include "det.cuh"
#include "pre.cuh"

namespace cald {

constexpr int REF_CAP = 50;     // cald_train.py:110-113: at most 50 reference boxes survive the sub-sampling
enum { AUG_FLIP = 0, AUG_CUTOUT = 1, AUG_RESIZE = 2, AUG_ROTATE = 3, AUG_IDENT = 4 };  // AUG_IDENT: boxes unchanged

struct RefSet {                 // per image, device
  int* n;                       // [B] number of reference rows (<= 50, duplicates allowed)
  int* n_det;                   // [B] detections of the reference view before sub-sampling
  float4* boxes;                // [B][50]
  float* prob_max;              // [B][50]
  int* prop_idx;                // [B][50]  row of the reference view's scores_cls
};

// grid = B.  lut[n][50] holds np.round(np.linspace(0, n-1, 50)).astype(int) for n in (40, det_cap].
__global__ void ref_prepare_kernel(DetOut det, int det_cap, const int* __restrict__ lut, RefSet ref) {
  const int b = blockIdx.x;
  const int nd = det.count[b];
  const int n = nd > 40 ? 50 : nd;
  for (int i = threadIdx.x; i < REF_CAP; i += blockDim.x) {
    float4 bx = make_float4(0, 0, 0, 0);
    float pm = 0.f;
    int pi = 0;
    if (i < n) {
      int src = nd > 40 ? lut[nd * 50 + i] : i;
      bx = det.boxes[(long long)b * det_cap + src];
      pm = det.prob_max[(long long)b * det_cap + src];
      pi = det.prop_idx[(long long)b * det_cap + src];
    }
    ref.boxes[b * REF_CAP + i] = bx;
    ref.prob_max[b * REF_CAP + i] = pm;
    ref.prop_idx[b * REF_CAP + i] = pi;
  }
  if (threadIdx.x == 0) { ref.n[b] = n; ref.n_det[b] = nd; }
}

// Per-view class-max vector (cald_train.py:114-116, 194-197): cls[l-1] = max score; python's negative index wraps
// label 0 to the last slot.  For the reference view only the sub-sampled rows count (they are a subset, and with
// duplicates the max is unchanged) -- but the rows dropped by the sub-sampling do NOT count, so the reference view
// passes its sub-sample LUT.  grid = views, block = 128.  out [views][ncls1] fp32.
__global__ void class_max_kernel(DetOut det, int det_cap, int ncls1, const int* __restrict__ lut, int use_lut,
                                 float* __restrict__ out) {
  extern __shared__ float cm[];
  const int v = blockIdx.x;
  for (int i = threadIdx.x; i < ncls1; i += blockDim.x) cm[i] = 0.f;
  __syncthreads();
  const int nd = det.count[v];
  const int n = (use_lut && nd > 40) ? 50 : nd;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int src = (use_lut && nd > 40) ? lut[nd * 50 + i] : i;
    int l = det.labels[(long long)v * det_cap + src];
    float s = det.scores[(long long)v * det_cap + src];
    int slot = l - 1;
    if (slot < 0) slot += ncls1;
    // scores are positive: integer atomicMax on the bit pattern is order preserving
    atomicMax(reinterpret_cast<int*>(&cm[slot]), __float_as_int(s));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncls1; i += blockDim.x) out[(long long)v * ncls1 + i] = cm[i];
}

// ---------------------------------------------------------------- cutout accept/reject (cald_helper.py:88-132, 226-243)
// Images are processed sequentially by ONE block because the reference consumes python's global RNG stream with a
// data-dependent number of draws: image b starts where image b-1 stopped.  `u` are raw random.random() doubles
// drawn by the host from the saved generator state; consumed[0] returns how many were used so the host can advance
// the real generator by exactly that amount.  random.uniform(a, b) = a + (b - a) * random().
__global__ void cutout_kernel(const RefSet ref, const int* __restrict__ img_hw /*[B][2]*/, int B, int n_cut,
                              const int* __restrict__ cut_nums /*[n_cut]*/, const double* __restrict__ u, int n_u,
                              CutRects* __restrict__ cuts /*[B][n_cut]*/, int* __restrict__ consumed) {
  __shared__ float s_max[32];
  int cursor = consumed[0];
  for (int b = 0; b < B; ++b) {
    const int n = ref.n[b];
    const double H = (double)img_hw[b * 2], W = (double)img_hw[b * 2 + 1];
    for (int ci = 0; ci < n_cut; ++ci) {
      CutRects* cr = &cuts[b * n_cut + ci];
      const int cut_num = cut_nums[ci];
      int count = 0;
      if (threadIdx.x == 0) cr->n = 0;
      if (n == 0) continue;  // the reference breaks out before augmenting (cald_train.py:118-121)
      for (int t = 0; t < 50 && count < cut_num; ++t) {
        if (cursor + 4 > n_u) break;
        const double ch = 0.05 * H + (0.2 * H - 0.05 * H) * u[cursor + 0];
        const double cw = 0.05 * W + (0.2 * W - 0.05 * W) * u[cursor + 1];
        const double left = 0.0 + ((W - cw) - 0.0) * u[cursor + 2];
        const double right = left + cw;
        const double top = 0.0 + ((H - ch) - 0.0) * u[cursor + 3];
        const double bottom = top + ch;
        cursor += 4;
        const float cl = (float)(int)left, ct = (float)(int)top, crr = (float)(int)right, cb = (float)(int)bottom;
        float best = -INFINITY;
        bool has_nan = false;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          float4 bx = ref.boxes[b * REF_CAP + i];
          float ix = fmaxf(fminf(crr, bx.z) - fmaxf(cl, bx.x), 0.f);
          float iy = fmaxf(fminf(cb, bx.w) - fmaxf(ct, bx.y), 0.f);
          float ratio = (ix * iy) / ((bx.z - bx.x) * (bx.w - bx.y));
          if (ratio != ratio) has_nan = true;
          best = fmaxf(best, ratio);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        has_nan = __any_sync(0xffffffffu, has_nan);
        if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = has_nan ? NAN : best;
        __syncthreads();
        float m = -INFINITY;
        bool nan = false;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { float q = s_max[w]; if (q != q) nan = true; m = fmaxf(m, q); }
        __syncthreads();
        // torch .max() propagates NaN; NaN > 0.4 and NaN < 0.1 are both False -> accepted (as in the reference)
        bool reject = !nan && (m > 0.4f || m < 0.1f);
        if (reject) continue;
        if (threadIdx.x == 0) {
          int k = cr->n;
          if (k < MAX_CUT) {
            cr->rect[k][0] = (int)left; cr->rect[k][1] = (int)top;
            cr->rect[k][2] = (int)right; cr->rect[k][3] = (int)bottom;
            cr->n = k + 1;
          }
        }
        count++;
      }
    }
  }
  if (threadIdx.x == 0) consumed[0] = cursor;
}

// ---------------------------------------------------------------- reference boxes mapped into each augmented view
struct AugGeom {       // per (image, aug), host-built
  int kind;
  float w, h;          // source image size
  float ratio;         // resize
  float m[6];          // rotate: float32 affine (after the nW/nH shift), row-major 2x3
  float sx, sy;        // rotate: new_image.width / w, new_image.height / h
};
// grid = B*A, block = 64.  out [B*A][50] boxes.
__global__ void aug_boxes_kernel(const RefSet ref, const AugGeom* __restrict__ geom, int A,
                                 float4* __restrict__ out) {
  const int ba = blockIdx.x, b = ba / A;
  const AugGeom g = geom[ba];
  const int i = threadIdx.x;
  if (i >= REF_CAP) return;
  float4 bx = ref.boxes[b * REF_CAP + i];
  float4 o = bx;
  if (g.kind == AUG_FLIP) {
    o.x = g.w - bx.z;
    o.z = g.w - bx.x;
  } else if (g.kind == AUG_RESIZE) {
    o = make_float4(bx.x * g.ratio, bx.y * g.ratio, bx.z * g.ratio, bx.w * g.ratio);
  } else if (g.kind == AUG_ROTATE) {
    float bw = bx.z - bx.x, bh = bx.w - bx.y;
    float cx[4] = {bx.x, bx.x + bw, bx.x, bx.z};
    float cy[4] = {bx.y, bx.y, bx.y + bh, bx.w};
    float xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float rx = fmaf(g.m[2], 1.f, fmaf(g.m[1], cy[k], g.m[0] * cx[k]));
      float ry = fmaf(g.m[5], 1.f, fmaf(g.m[4], cy[k], g.m[3] * cx[k]));
      xmin = fminf(xmin, rx); xmax = fmaxf(xmax, rx);
      ymin = fminf(ymin, ry); ymax = fmaxf(ymax, ry);
    }
    o.x = fminf(fmaxf(xmin / g.sx, 0.f), g.w);
    o.y = fminf(fmaxf(ymin / g.sy, 0.f), g.h);
    o.z = fminf(fmaxf(xmax / g.sx, 0.f), g.w);
    o.w = fminf(fmaxf(ymax / g.sy, 0.f), g.h);
  }
  out[(long long)ba * REF_CAP + i] = o;
}

// ---------------------------------------------------------------- the paired-prediction reduction (cald_train.py:189-223)
// grid = (A, B); block = 32 * CONS_WARPS; one warp per reference box, lanes stride over the view's detections.
// scores_cls rows: ref view row = ref_scores[(b*cap + prop_idx) * C ...], aug view likewise.
constexpr int CONS_WARPS = 8;
struct ConsArgs {
  RefSet ref;
  const float4* aug_boxes;   // [B*A][50]
  DetOut det;                // detections of the aug views, view index = b*A + a
  int det_cap;
  const float* ref_scores;   // [B][cap][C] softmax rows of the reference views
  const float* aug_scores;   // [B*A][cap][C]
  int cap, C, A;
  float bp;
  float* out;                // [B][A] consistency_img
};
__global__ void __launch_bounds__(32 * CONS_WARPS) consistency_kernel(ConsArgs a) {
  __shared__ float s_min[CONS_WARPS];
  const int aug = blockIdx.x, b = blockIdx.y, ba = b * a.A + aug;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nref = a.ref.n[b];
  const int nd = a.det.count[ba];
  float best = INFINITY;  // min over this warp's reference boxes
  if (nd > 0) {
    for (int r = warp; r < nref; r += CONS_WARPS) {
      const float4 ab = a.aug_boxes[(long long)ba * REF_CAP + r];
      const float a_area = (ab.z - ab.x) * (ab.w - ab.y);
      // ---- IoU against every detection, running (max, first arg-max); NaN counts as the maximum (torch semantics)
      float mx = -INFINITY;
      int arg = 0x7fffffff;
      for (int j = lane; j < nd; j += 32) {
        const float4 bb = a.det.boxes[(long long)ba * a.det_cap + j];
        float width = fminf(ab.z, bb.z) - fmaxf(ab.x, bb.x);
        float height = fminf(ab.w, bb.w) - fmaxf(ab.y, bb.y);
        float b_area = (bb.z - bb.x) * (bb.w - bb.y);
        float inter = width * height;
        float iou = inter / (a_area + b_area - inter);
        if (width < 0.f) iou = 0.f;
        if (height < 0.f) iou = 0.f;
        bool better = (iou != iou) ? (mx == mx) : (iou > mx);
        if (better) { mx = iou; arg = j; }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
        bool onan = (omx != omx), mnan = (mx != mx);
        bool take = onan ? (!mnan || oarg < arg) : (!mnan && (omx > mx || (omx == mx && oarg < arg)));
        if (take) { mx = omx; arg = oarg; }
      }
      // ---- JS divergence between the two full class vectors (scipy.stats.entropy semantics, fp32)
      const float* p = a.ref_scores + ((long long)b * a.cap + a.ref.prop_idx[b * REF_CAP + r]) * a.C;
      const float* q = a.aug_scores + ((long long)ba * a.cap + a.det.prop_idx[(long long)ba * a.det_cap + arg]) * a.C;
      float sp = 0.f, sq = 0.f, sm = 0.f;
      for (int c = lane; c < a.C; c += 32) {
        float pc = p[c], qc = q[c];
        sp += pc; sq += qc; sm += (pc + qc) / 2.f;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
      }
      float e1 = 0.f, e2 = 0.f;
      for (int c = lane; c < a.C; c += 32) {
        float pc = p[c], qc = q[c];
        float m = ((pc + qc) / 2.f) / sm;
        float pn = pc / sp, qn = qc / sq;
        // rel_entr evaluates x*log(x/y) in double and rounds once to fp32
        if (pn > 0.f) e1 += (float)((double)pn * log((double)pn / (double)m));
        if (qn > 0.f) e2 += (float)((double)qn * log((double)qn / (double)m));
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        e1 += __shfl_xor_sync(0xffffffffu, e1, o);
        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
      }
      float js = 0.5f * e1 + 0.5f * e2;
      if (js < 0.f) js = 0.f;
      const float pm = a.ref.prob_max[b * REF_CAP + r] + a.det.prob_max[(long long)ba * a.det_cap + arg];
      const float val = fabsf(mx + (0.5f * (1.f - js)) * pm - a.bp);
      // python: min(consistency_img, v) keeps the old value when v is NaN
      if (val < best) best = val;
    }
  }
  if (lane == 0) s_min[warp] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 1.0f;  // consistency_img starts at 1.0
    for (int w = 0; w < CONS_WARPS; ++w) if (s_min[w] < m) m = s_min[w];
    if (nd == 0) m = 0.f;  // empty augmented prediction contributes 0.0 (cald_train.py:198-201)
    a.out[b * a.A + aug] = m;
  }
}

// ================================================================ baseline scorers on the same detections
// ---------------------------------------------------------------- LT/C: localisation tightness (lt_c_train.py:89-121)
// per image: min(1.0, min over detections |calcu_iou(box, prop) + prob_max - 1|) with the script's own IoU
// (+1 on the intersection sides and on the height of each area only, lt_c_train.py:95-102).  grid = views, block = 128.
__global__ void ltc_kernel(DetOut det, int det_cap, float* __restrict__ out) {
  __shared__ float s_min[4];
  const int v = blockIdx.x;
  const int n = det.count[v];
  float best = INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 a = det.boxes[(long long)v * det_cap + i];
    const float4 b = det.props[(long long)v * det_cap + i];
    const float width = (fminf(a.z, b.z) - fmaxf(a.x, b.x)) + 1.f;
    const float height = (fminf(a.w, b.w) - fmaxf(a.y, b.y)) + 1.f;
    float iou = 0.f;
    if (!(width <= 0.f || height <= 0.f)) {
      const float aa = (a.z - a.x) * ((a.w - a.y) + 1.f);
      const float ba = (b.z - b.x) * ((b.w - b.y) + 1.f);
      const float inter = width * height;
      iou = inter / ((aa + ba) - inter);
    }
    const float u = fabsf((iou + det.prob_max[(long long)v * det_cap + i]) - 1.f);
    if (u < best) best = u;  // python min(): a NaN candidate is never taken
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 1.0f;  // uncertainty starts at 1.0
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) if (s_min[w] < m) m = s_min[w];
    out[v] = m;
  }
}

// ---------------------------------------------------------------- LS+C: localisation stability (ls_c_train.py:108-155)
// keys for torch.topk(prob_max, 30): descending prob_max, ties -> lower index.  grid = views.
__global__ void lsc_keys_kernel(DetOut det, int det_cap, unsigned long long* __restrict__ keys) {
  const int v = blockIdx.x;
  const int n = det.count[v];
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    keys[(long long)v * det_cap + i] =
        ((unsigned long long)desc_key(det.prob_max[(long long)v * det_cap + i]) << 32) | (unsigned)i;
}
// grid = B, block = 32 * CONS_WARPS.  sel: sorted top-30 keys of the reference view (topk_select_kernel output,
// stride TOPK_MAX); aug detections: view index b*A + a.  out[b] = sum(pm * mean_a maxIoU) / sum(pm) - max(1 - pm).
constexpr int LSC_REF = 30;
__global__ void __launch_bounds__(32 * CONS_WARPS) lsc_kernel(DetOut ref, DetOut aug, int det_cap, int A,
                                                              const unsigned long long* __restrict__ sel,
                                                              double* __restrict__ out) {
  __shared__ float s_iou[LSC_REF][8];  // [ref box][aug view], A <= 8
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nd = ref.count[b];
  const int nref = nd > LSC_REF ? LSC_REF : nd;
  for (int t = warp; t < nref * A; t += CONS_WARPS) {
    const int r = t / A, a = t - r * A;
    const int src = nd > LSC_REF ? (int)(sel[(long long)b * TOPK_MAX + r] & 0xffffffffu) : r;
    const float4 ab = ref.boxes[(long long)b * det_cap + src];
    const long long ba = (long long)b * A + a;
    const int na = aug.count[ba];
    const float a_area = (ab.z - ab.x) * (ab.w - ab.y);
    float mx = -INFINITY;
    for (int j = lane; j < na; j += 32) {
      const float4 bb = aug.boxes[ba * det_cap + j];
      const float width = fminf(ab.z, bb.z) - fmaxf(ab.x, bb.x);
      const float height = fminf(ab.w, bb.w) - fmaxf(ab.y, bb.y);
      const float b_area = (bb.z - bb.x) * (bb.w - bb.y);
      const float inter = width * height;
      float iou = inter / (a_area + b_area - inter);
      if (width < 0.f) iou = 0.f;
      if (height < 0.f) iou = 0.f;
      if (iou != iou || iou > mx) mx = (mx != mx) ? mx : iou;  // torch.max propagates NaN
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      if (om != om || (mx == mx && om > mx)) mx = om;
    }
    if (lane == 0) s_iou[r][a] = na > 0 ? mx : 0.f;  // an augmented view without boxes adds nothing (ls_c_train.py:136-137)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (nd == 0) { out[b] = 0.0; return; }  // ls_c_train.py:118-120
    double num = 0.0, den = 0.0;
    float u = -INFINITY;
    for (int r = 0; r < nref; ++r) {
      const int src = nd > LSC_REF ? (int)(sel[(long long)b * TOPK_MAX + r] & 0xffffffffu) : r;
      const float pm = ref.prob_max[(long long)b * det_cap + src];
      double st = 0.0;
      for (int a = 0; a < A; ++a) st += (double)s_iou[r][a];  // python float accumulation of .item() values
      st /= 6.0;                                               // the script divides by the literal 6.0
      num += (double)pm * st;
      den += (double)pm;
      u = fmaxf(u, 1.f - pm);
    }
    out[b] = num / den - (double)u;
  }
}

}  // namespace cald
