// Discrete / HBM-bound detection stages, all on device with fixed-capacity buffers and
// device-side counts (no host-visible sizes, so a whole pass is one stream of launches):
//   RPN: key build -> radix-select top-k + sort -> decode/clip/filter -> bitmask NMS -> merge top-1000
//   RoI: level map + RoIAlign (NHWC gather) ; softmax ; per-class NMS ; top-100 detections
// Arithmetic follows the reference's CPU fp32 ops one to one (unfused mul/add: this
// translation unit is compiled with -fmad=false).  Reference: tv:models/detection/rpn.py:242-297,
// tv:models/detection/_utils.py:183-224, tv:ops/boxes.py, tv:ops/poolers.py, tv:ops/roi_align.py,
// detection/frcnn_la.py:32-87,292-315.
#pragma once
#include "common.cuh"

namespace cald {

constexpr int RPN_LEVELS = 5;
constexpr int TOPK_MAX = 1024;
constexpr float BBOX_CLIP = 4.135166556742356f;  // log(1000/16)

struct RpnLevel {
  const float* out;   // [V][h][w][16] fp32: 3 objectness logits, then 12 deltas (a*4 + k)
  int h, w;
  int stride_h, stride_w;
  float base[3][4];   // cell anchors (x1,y1,x2,y2) per aspect ratio
  int n;              // h*w*3
  int off;            // offset of this level in the concatenated anchor index space
};
struct RpnLevels {
  RpnLevel lv[RPN_LEVELS];
  int total;          // sum of n
};

__device__ __forceinline__ uint32_t desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  uint32_t asc = u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
  return ~asc;
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t asc = ~k;
  uint32_t u = (asc & 0x80000000u) ? (asc ^ 0x80000000u) : ~asc;
  return __uint_as_float(u);
}

// ---------------------------------------------------------------- block bitonic sort (ascending) on smem u64
__device__ inline void block_bitonic_sort(unsigned long long* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------- RPN objectness keys
// key = (descending-order bits of logit) << 32 | index within level  (all keys distinct)
__global__ void rpn_keys_kernel(RpnLevels L, int V, unsigned long long* __restrict__ keys) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)V * L.total) return;
  int v = (int)(gid / L.total);
  int e = (int)(gid % L.total);
  int l = 0;
#pragma unroll
  for (int i = 1; i < RPN_LEVELS; ++i) if (e >= L.lv[i].off) l = i;
  const RpnLevel& lv = L.lv[l];
  int r = e - lv.off;
  int a = r % 3;
  int pix = r / 3;
  float logit = lv.out[((long long)v * lv.h * lv.w + pix) * 16 + a];
  keys[gid] = ((unsigned long long)desc_key(logit) << 32) | (unsigned)r;
}

// ---------------------------------------------------------------- generic top-k (k smallest distinct u64 keys), sorted
// One CTA per group.  keys + group_off[g] .. +group_n[g]; writes min(n,k) sorted keys to out[g*TOPK_MAX ..].
struct TopkGroups {
  const unsigned long long* keys;
  long long stride_outer;  // group g = (outer, inner): offset = outer*stride_outer + inner_off[inner]
  int inner;               // number of inner groups
  int inner_off[RPN_LEVELS];
  int inner_n[RPN_LEVELS];
  const int* dyn_n;        // optional device array of per-group n (overrides inner_n)
  int k;
};

__global__ void __launch_bounds__(1024) topk_select_kernel(TopkGroups G, unsigned long long* __restrict__ out,
                                                           int* __restrict__ out_count) {
  __shared__ unsigned long long buf[TOPK_MAX];
  __shared__ int hist[256];
  __shared__ unsigned long long s_prefix, s_mask;
  __shared__ int s_k, s_cnt, s_done;
  const int g = blockIdx.x;
  const int outer = g / G.inner, inner = g % G.inner;
  const unsigned long long* keys = G.keys + (long long)outer * G.stride_outer + G.inner_off[inner];
  int n = G.dyn_n ? G.dyn_n[g] : G.inner_n[inner];
  const int k = n < G.k ? n : G.k;
  const int tid = threadIdx.x;
  for (int i = tid; i < TOPK_MAX; i += blockDim.x) buf[i] = ~0ull;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_k = k; s_cnt = 0; s_done = 0; }
  __syncthreads();
  if (n <= TOPK_MAX) {
    for (int i = tid; i < n; i += blockDim.x) buf[i] = keys[i];
  } else {
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix, mask = s_mask;
      // Objectness keys of one level share their leading bytes, so one atomicAdd per key would serialise the whole
      // CTA on one or two bins: each thread run-length-merges equal consecutive bins and flushes once per run.
      int last = -1, run = 0;
      for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long key = keys[i];
        if ((key & mask) == prefix) {
          const int bin = (int)((key >> shift) & 255);
          if (bin == last) {
            ++run;
          } else {
            if (run) atomicAdd(&hist[last], run);
            last = bin;
            run = 1;
          }
        }
      }
      if (run) atomicAdd(&hist[last], run);
      __syncthreads();
      if (tid == 0) {
        int kk = s_k, cum = 0, b = 0;
        for (; b < 256; ++b) {
          if (cum + hist[b] >= kk) break;
          cum += hist[b];
        }
        s_k = kk - cum;
        s_prefix = prefix | ((unsigned long long)b << shift);
        s_mask = mask | (255ull << shift);
        // After the four score bytes the selected bin holds the keys that TIE with the k-th score (normally one).
        // If everything below them plus the whole tie group fits the sort buffer, the four index-byte passes are
        // unnecessary: collect all of them and let the sort (score, then index) pick the first k.
        if (pass == 3 && (k - s_k) + hist[b] <= TOPK_MAX) {
          s_prefix |= 0xffffffffull;
          s_done = 1;
        }
      }
      __syncthreads();
      if (s_done) break;
    }
    const unsigned long long T = s_prefix;  // the k-th smallest key (or the last key of its tie group)
    for (int i = tid; i < n; i += blockDim.x) {
      unsigned long long key = keys[i];
      if (key <= T) {
        int slot = atomicAdd(&s_cnt, 1);
        if (slot < TOPK_MAX) buf[slot] = key;
      }
    }
  }
  __syncthreads();
  block_bitonic_sort(buf, TOPK_MAX);
  for (int i = tid; i < TOPK_MAX; i += blockDim.x) out[(long long)g * TOPK_MAX + i] = buf[i];
  if (tid == 0) out_count[g] = k;
}

// ---------------------------------------------------------------- block exclusive scan of 0/1 flags (blockDim = 1024)
__device__ inline int block_excl_scan_1024(int flag, int* total) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned bal = __ballot_sync(0xffffffffu, flag);
  int inwarp = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) wsum[wid] = __popc(bal);
  __syncthreads();
  if (wid == 0) {
    int v = wsum[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    wsum[lane] = s - v;
    if (lane == 31) *total = s;
  }
  __syncthreads();
  int r = wsum[wid] + inwarp;
  __syncthreads();
  return r;
}

__device__ __forceinline__ void decode_box(const float* d, float ax1, float ay1, float ax2, float ay2, float wx,
                                           float wy, float ww, float wh, float* o) {
  float widths = ax2 - ax1, heights = ay2 - ay1;
  float cx = ax1 + 0.5f * widths, cy = ay1 + 0.5f * heights;
  float dx = d[0] / wx, dy = d[1] / wy, dw = d[2] / ww, dh = d[3] / wh;
  dw = fminf(dw, BBOX_CLIP);
  dh = fminf(dh, BBOX_CLIP);
  float pcx = dx * widths + cx, pcy = dy * heights + cy;
  float pw = expf(dw) * widths, ph = expf(dh) * heights;
  float hw = 0.5f * pw, hh = 0.5f * ph;
  o[0] = pcx - hw; o[1] = pcy - hh; o[2] = pcx + hw; o[3] = pcy + hh;
}
__device__ __forceinline__ void clip_box(float* b, float img_h, float img_w) {
  b[0] = fminf(fmaxf(b[0], 0.f), img_w);
  b[2] = fminf(fmaxf(b[2], 0.f), img_w);
  b[1] = fminf(fmaxf(b[1], 0.f), img_h);
  b[3] = fminf(fmaxf(b[3], 0.f), img_h);
}

// ---------------------------------------------------------------- RPN decode of the selected anchors
// grid = V * 5, block = 1024.  In: sorted keys.  Out (per group, capacity TOPK_MAX): boxes, scores, count.
__global__ void __launch_bounds__(1024) rpn_decode_kernel(RpnLevels L, const unsigned long long* __restrict__ sel,
                                                          const int* __restrict__ sel_count,
                                                          const int* __restrict__ image_hw /*[V][2]*/,
                                                          float min_size, float4* __restrict__ boxes,
                                                          float* __restrict__ scores, int* __restrict__ count) {
  __shared__ int s_total;
  const int g = blockIdx.x, v = g / RPN_LEVELS, l = g % RPN_LEVELS;
  const RpnLevel& lv = L.lv[l];
  const int i = threadIdx.x;
  const int n = sel_count[g];
  int ok = 0;
  float b[4] = {0, 0, 0, 0};
  float sc = 0.f;
  if (i < n) {
    unsigned long long key = sel[(long long)g * TOPK_MAX + i];
    int r = (int)(key & 0xffffffffu);
    float logit = key_to_float((uint32_t)(key >> 32));
    int a = r % 3, pix = r / 3;
    int x = pix % lv.w, y = pix / lv.w;
    const float* o = lv.out + ((long long)v * lv.h * lv.w + pix) * 16;
    float sx = (float)(x * lv.stride_w), sy = (float)(y * lv.stride_h);
    float d[4] = {o[3 + a * 4 + 0], o[3 + a * 4 + 1], o[3 + a * 4 + 2], o[3 + a * 4 + 3]};
    decode_box(d, sx + lv.base[a][0], sy + lv.base[a][1], sx + lv.base[a][2], sy + lv.base[a][3], 1.f, 1.f, 1.f, 1.f,
               b);
    clip_box(b, (float)image_hw[v * 2], (float)image_hw[v * 2 + 1]);
    sc = 1.f / (1.f + expf(-logit));
    ok = ((b[2] - b[0]) >= min_size) && ((b[3] - b[1]) >= min_size) && (sc >= 0.f);
  }
  int pos = block_excl_scan_1024(ok, &s_total);
  if (ok) {
    boxes[(long long)g * TOPK_MAX + pos] = make_float4(b[0], b[1], b[2], b[3]);
    scores[(long long)g * TOPK_MAX + pos] = sc;
  }
  if (i == 0) count[g] = s_total;
}

// ---------------------------------------------------------------- bitmask NMS for <= 1024 score-sorted boxes
// torchvision CPU kernel arithmetic: area = (x2-x1)*(y2-y1); ovr = inter / (a_i + a_j - inter); suppress if ovr > thresh
// (comparison in double, as the C++ kernel compares a float with a double threshold).
// dynamic smem: mask[1024][16] u64 (128 KB) + boxes float4[1024] (16 KB).
constexpr int NMS_SMEM = TOPK_MAX * 16 * 8 + TOPK_MAX * 16;

__device__ inline void block_nms_1024(const float4* sbox, unsigned long long* mask, int n, double thresh,
                                      int* keep_idx /*smem or global, cap 1024*/, int* keep_count) {
  const int words = (n + 63) >> 6;
  // item -> (column word cw, row i) with i fastest: the lanes of a warp share cw, so every sbox[j] read below is a
  // shared-memory broadcast (with cw fastest the 32 lanes would read float4s 1 KB apart: a 32-way bank conflict)
  for (int item = threadIdx.x; item < n * words; item += blockDim.x) {
    int cw = item / n, i = item - cw * n;
    unsigned long long m = 0;
    if (cw * 64 + 63 > i) {
      float4 bi = sbox[i];
      float ai = (bi.z - bi.x) * (bi.w - bi.y);
      int j0 = cw * 64;
      for (int jj = 0; jj < 64; ++jj) {
        int j = j0 + jj;
        if (j > i && j < n) {
          float4 bj = sbox[j];
          float aj = (bj.z - bj.x) * (bj.w - bj.y);
          float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
          float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
          float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
          float inter = w * h;
          float ovr = inter / (ai + aj - inter);
          if ((double)ovr > thresh) m |= (1ull << jj);
        }
      }
    }
    mask[i * 16 + cw] = m;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    unsigned long long removed = 0;
    int kept = 0;
    for (int i = 0; i < n; ++i) {
      unsigned long long r = __shfl_sync(0xffffffffu, removed, i >> 6);
      if (!((r >> (i & 63)) & 1ull)) {
        if (lane == 0) keep_idx[kept] = i;
        kept++;
        if (lane < words) removed |= mask[i * 16 + lane];
      }
    }
    if (lane == 0) *keep_count = kept;
  }
  __syncthreads();
}

// grid = groups (V*5); boxes/scores already score-sorted; writes ordered keep list + count.
__global__ void __launch_bounds__(1024) nms_groups_kernel(const float4* __restrict__ boxes,
                                                          const int* __restrict__ count, double thresh,
                                                          int* __restrict__ keep_idx, int* __restrict__ keep_count) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(nms_smem);
  float4* sbox = reinterpret_cast<float4*>(nms_smem + TOPK_MAX * 16 * 8);
  const int g = blockIdx.x;
  const int n = count[g];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sbox[i] = boxes[(long long)g * TOPK_MAX + i];
  __syncthreads();
  block_nms_1024(sbox, mask, n, thresh, keep_idx + (long long)g * TOPK_MAX, keep_count + g);
}

// ---------------------------------------------------------------- merge levels: top post_n by score (stable)
// grid = V, block = 1024, smem 8192 u64.
constexpr int MERGE_CAP = 8192;
__global__ void __launch_bounds__(1024) rpn_merge_kernel(const float4* __restrict__ boxes,
                                                         const float* __restrict__ scores,
                                                         const int* __restrict__ keep_idx,
                                                         const int* __restrict__ keep_count, int post_n,
                                                         float4* __restrict__ props, float* __restrict__ prop_scores,
                                                         int* __restrict__ prop_count, int prop_cap) {
  extern __shared__ __align__(16) unsigned char msm[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(msm);
  __shared__ int base[RPN_LEVELS + 1];
  const int v = blockIdx.x;
  if (threadIdx.x == 0) {
    int s = 0;
    for (int l = 0; l < RPN_LEVELS; ++l) { base[l] = s; s += keep_count[v * RPN_LEVELS + l]; }
    base[RPN_LEVELS] = s;
  }
  for (int i = threadIdx.x; i < MERGE_CAP; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  for (int l = 0; l < RPN_LEVELS; ++l) {
    const int g = v * RPN_LEVELS + l;
    const int kc = keep_count[g];
    for (int j = threadIdx.x; j < kc; j += blockDim.x) {
      int idx = keep_idx[(long long)g * TOPK_MAX + j];
      float sc = scores[(long long)g * TOPK_MAX + idx];
      // low word: concat position (stable order); the source slot is recovered from it below
      keys[base[l] + j] = ((unsigned long long)desc_key(sc) << 32) | (unsigned)(base[l] + j);
    }
  }
  __syncthreads();
  block_bitonic_sort(keys, MERGE_CAP);
  const int total = base[RPN_LEVELS];
  const int m = total < post_n ? total : post_n;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    int pos = (int)(keys[i] & 0xffffffffu);
    int l = 0;
    for (int q = 1; q < RPN_LEVELS; ++q) if (pos >= base[q]) l = q;
    const int g = v * RPN_LEVELS + l;
    int idx = keep_idx[(long long)g * TOPK_MAX + (pos - base[l])];
    props[(long long)v * prop_cap + i] = boxes[(long long)g * TOPK_MAX + idx];
    prop_scores[(long long)v * prop_cap + i] = scores[(long long)g * TOPK_MAX + idx];
  }
  for (int i = m + threadIdx.x; i < prop_cap; i += blockDim.x) {
    props[(long long)v * prop_cap + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    prop_scores[(long long)v * prop_cap + i] = 0.f;
  }
  if (threadIdx.x == 0) prop_count[v] = m;
}

// ---------------------------------------------------------------- multi-scale RoIAlign (7x7, sampling 2, aligned=False)
// Features: 4 pyramid levels, split-pl16 NHWC [V][h][w][C].  Output rows [V*cap][49][C] split pl16 (the K-major A
// operand of fc6, weight columns permuted to (ph, pw, c) at load time).  One CTA per RoI, one warp per bin, 8 ch/lane.
struct RoiFeats {
  const pl16* hi[4];
  const pl16* lo[4];
  int h[4], w[4];
  float scale[4];
  int C;
};
__device__ __forceinline__ void bilinear_prep(float v, int size, int& lo, int& hi, float& l, float& h, bool& bad) {
  bad = (v < -1.0f) || (v > (float)size);
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else { hi = lo + 1; }
  l = v - (float)lo;
  h = 1.f - l;
}
// The 14 sample rows and 14 sample columns of a RoI (7 bins x 2 samples) are prepared ONCE per RoI by 28 threads and
// broadcast from shared memory; before, every lane of every warp recomputed them for each of its samples (about a
// quarter of the kernel's instructions, which is issue-bound: ncu issue-active 78 %, L1 hit rate 78 %).
// 7 warps: each owns exactly 7 of the 49 bins.
// FUSED: the four weighted taps of a sample are accumulated with FFMA (4 instructions per value instead of the 4 FMUL +
// 4 FADD of the unfused ATen expression; the kernel is instruction-bound).  The difference to the unfused form is a few
// 2^-24 relative, two orders below the 2^-17 rounding of the split-pl16 features the taps are read from.
constexpr int ROI_THREADS = 224;
template <bool FUSED>
__global__ void __launch_bounds__(ROI_THREADS) roialign_kernel(RoiFeats F, const float4* __restrict__ props,
                                                               const int* __restrict__ prop_count, int cap,
                                                               pl16* __restrict__ ohi, pl16* __restrict__ olo) {
  __shared__ int s_lo[2][14], s_hi[2][14], s_bad[2][14];   // element offsets of the low / high row (dim 0) or column (dim 1)
  __shared__ float s_l[2][14], s_h[2][14];
  const int v = blockIdx.y, r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = F.C;  // 256
  const long long orow = ((long long)v * cap + r) * 49 * C;
  const bool live = r < prop_count[v];
  float4 box = live ? props[(long long)v * cap + r] : make_float4(0, 0, 0, 0);
  // level mapper: floor(4 + log2(sqrt(area)/224) + 1e-6) clamped to [2,5]
  float area = (box.z - box.x) * (box.w - box.y);
  float s = sqrtf(area);
  float lvf = floorf(4.f + log2f(s / 224.f) + 1e-6f);
  lvf = fminf(fmaxf(lvf, 2.f), 5.f);
  const int lv = (int)lvf - 2;
  const int H = F.h[lv], W = F.w[lv];
  const float sc = F.scale[lv];
  const float x1 = box.x * sc, y1 = box.y * sc, x2 = box.z * sc, y2 = box.w * sc;
  const float rw = fmaxf(x2 - x1, 1.f), rh = fmaxf(y2 - y1, 1.f);
  const float bw = rw / 7.f, bh = rh / 7.f;
  if (threadIdx.x < 28) {
    const int dim = threadIdx.x / 14, k = threadIdx.x - dim * 14;
    const int pb = k >> 1, i = k & 1;
    const float start = dim == 0 ? y1 : x1, bsz = dim == 0 ? bh : bw;
    const float coord = start + (float)pb * bsz + ((float)i + 0.5f) * bsz / 2.f;
    int lo, hi; float l, h; bool bad;
    bilinear_prep(coord, dim == 0 ? H : W, lo, hi, l, h, bad);
    // one feature level of one view is < 2^31 elements: 32-bit offsets (row * W * C, column * C)
    const int pitch = dim == 0 ? W * C : C;
    s_lo[dim][k] = lo * pitch; s_hi[dim][k] = hi * pitch; s_l[dim][k] = l; s_h[dim][k] = h; s_bad[dim][k] = bad ? 1 : 0;
  }
  __syncthreads();
  const pl16* fhi = F.hi[lv] + (long long)v * H * W * C + lane * 8;
  const pl16* flo = F.lo[lv] ? F.lo[lv] + (long long)v * H * W * C + lane * 8 : nullptr;
  for (int bin = warp; bin < 49; bin += ROI_THREADS / 32) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // The lo plane carries (v - hi) * 2^11, i.e. 2^-12 of the value: its four-tap interpolation runs on packed half2
    // FMAs (weights and sum rounded to half: 2^-11 of 2^-12 = 2^-23 of the value, below fp32's own rounding) and is
    // joined once per bin.  Only the hi plane is converted to fp32 -- the kernel is instruction-bound (ncu: issue
    // active 78 %), and this removes a third of its instructions.
    __half2 lacc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) lacc[q] = __float2half2_rn(0.f);
    if (live) {
      const int ph = bin / 7, pw = bin - ph * 7;
#pragma unroll
      for (int iy = 0; iy < 2; ++iy) {
        const int ky = ph * 2 + iy;
        const int ylo = s_lo[0][ky], yhi = s_hi[0][ky];
        const float ly = s_l[0][ky], hy = s_h[0][ky];
        const bool ybad = s_bad[0][ky] != 0;
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
          const int kx = pw * 2 + ix;
          if (ybad || s_bad[1][kx] != 0) continue;
          const int xlo = s_lo[1][kx], xhi = s_hi[1][kx];
          const float lx = s_l[1][kx], hx = s_h[1][kx];
          const float w[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
          const int o[4] = {ylo + xlo, ylo + xhi, yhi + xlo, yhi + xhi};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            // plain (coherent-path) loads: the non-coherent LDG.CONSTANT form measured 45 % slower here
            const uint4 a = *reinterpret_cast<const uint4*>(fhi + o[t]);
            const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
            float vv[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) unpack2(pa[q], vv[2 * q], vv[2 * q + 1]);
            if (FUSED) {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[t], vv[k], acc[k]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] += w[t] * vv[k];
            }
            if (flo) {
              const uint4 b = *reinterpret_cast<const uint4*>(flo + o[t]);
              const __half2* pb2 = reinterpret_cast<const __half2*>(&b);
              const __half2 wh = __float2half2_rn(w[t]);
#pragma unroll
              for (int q = 0; q < 4; ++q) lacc[q] = __hfma2(wh, pb2[q], lacc[q]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 l2 = __half22float2(lacc[q]);
        acc[2 * q] = fmaf(l2.x, CALD_LO_INV, acc[2 * q]) / 4.f;
        acc[2 * q + 1] = fmaf(l2.y, CALD_LO_INV, acc[2 * q + 1]) / 4.f;
      }
    }
    uint32_t ph4[4], pl4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_pack2(acc[2 * q], acc[2 * q + 1], ph4[q], pl4[q]);
    const long long off = orow + (long long)bin * C + lane * 8;
    *reinterpret_cast<uint4*>(ohi + off) = make_uint4(ph4[0], ph4[1], ph4[2], ph4[3]);
    if (olo) *reinterpret_cast<uint4*>(olo + off) = make_uint4(pl4[0], pl4[1], pl4[2], pl4[3]);
  }
}

// ---------------------------------------------------------------- softmax over classes (one warp per proposal)
// head: [V*cap][ld] fp32, columns [0,C) = class logits, [C, 5C) = box deltas (class-major, 4 per class)
__global__ void softmax_rows_kernel(const float* __restrict__ head, int ld, int C, long long rows,
                                    float* __restrict__ scores /*[rows][C]*/, float* __restrict__ prob_max) {
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = head + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(x[c] - m);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float pm = -INFINITY;
  for (int c = lane; c < C; c += 32) {
    float p = expf(x[c] - m) / s;
    scores[row * C + c] = p;
    if (c >= 1) pm = fmaxf(pm, p);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, o));
  if (lane == 0) prob_max[row] = pm;
}

// ---------------------------------------------------------------- per-class NMS of the box head's candidates
// grid = (C-1, V), block 1024, dynamic smem NMS_SMEM + keys.  Appends the kept candidates' keys
// (desc(score) << 32 | p*(C-1) + (c-1)) to kept_keys[v][...] (unordered; the final top-k sorts them).
__global__ void __launch_bounds__(1024) det_class_nms_kernel(const float* __restrict__ head, int ld, int C,
                                                             const float* __restrict__ scores,
                                                             const float4* __restrict__ props,
                                                             const int* __restrict__ prop_count, int cap,
                                                             const int* __restrict__ image_hw, float score_thresh,
                                                             double nms_thresh,
                                                             unsigned long long* __restrict__ kept_keys,
                                                             int* __restrict__ kept_count, int kept_cap) {
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(dsm);
  float4* sbox = reinterpret_cast<float4*>(dsm + TOPK_MAX * 16 * 8);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dsm + NMS_SMEM);
  int* keep_idx = reinterpret_cast<int*>(dsm + NMS_SMEM + TOPK_MAX * 8);
  __shared__ int s_n, s_keep, s_base;
  const int c = blockIdx.x + 1, v = blockIdx.y;
  const int np = prop_count[v];
  const int p = threadIdx.x;
  float sc = 0.f;
  int ok = 0;
  if (p < np) {
    sc = scores[((long long)v * cap + p) * C + c];
    ok = sc > score_thresh;
  }
  keys[p] = ~0ull;
  __syncthreads();
  int pos = block_excl_scan_1024(ok, &s_n);
  if (ok) keys[pos] = ((unsigned long long)desc_key(sc) << 32) | (unsigned)p;
  __syncthreads();
  const int n = s_n;
  if (n == 0) return;
  block_bitonic_sort(keys, TOPK_MAX);
  if (p < n) {
    int pp = (int)(keys[p] & 0xffffffffu);
    const float* d = head + ((long long)v * cap + pp) * ld + C + c * 4;
    float4 pr = props[(long long)v * cap + pp];
    float b[4];
    float dd[4] = {d[0], d[1], d[2], d[3]};
    decode_box(dd, pr.x, pr.y, pr.z, pr.w, 10.f, 10.f, 5.f, 5.f, b);
    clip_box(b, (float)image_hw[v * 2], (float)image_hw[v * 2 + 1]);
    sbox[p] = make_float4(b[0], b[1], b[2], b[3]);
  }
  __syncthreads();
  block_nms_1024(sbox, mask, n, nms_thresh, keep_idx, &s_keep);
  if (threadIdx.x == 0) s_base = atomicAdd(&kept_count[v], s_keep);
  __syncthreads();
  for (int j = threadIdx.x; j < s_keep; j += blockDim.x) {
    unsigned long long key = keys[keep_idx[j]];
    unsigned pp = (unsigned)(key & 0xffffffffu);
    int slot = s_base + j;
    if (slot < kept_cap)
      kept_keys[(long long)v * kept_cap + slot] = (key & 0xffffffff00000000ull) | (pp * (unsigned)(C - 1) + (unsigned)(c - 1));
  }
}

// ---------------------------------------------------------------- final detections (<= det_cap per view)
struct DetOut {
  int* count;        // [V]
  float4* boxes;     // [V][det_cap]  original-image coordinates
  float4* props;     // [V][det_cap]
  float* scores;     // [V][det_cap]
  float* prob_max;   // [V][det_cap]
  int* labels;       // [V][det_cap]
  int* prop_idx;     // [V][det_cap]  row of scores_cls
};
__global__ void det_gather_kernel(const unsigned long long* __restrict__ top_keys, const int* __restrict__ top_count,
                                  const float* __restrict__ head, int ld, int C, const float* __restrict__ scores,
                                  const float* __restrict__ prob_max, const float4* __restrict__ props, int cap,
                                  const int* __restrict__ image_hw, const float* __restrict__ ratio_hw /*[V][2]*/,
                                  int det_cap, DetOut out) {
  const int v = blockIdx.x, i = threadIdx.x;
  int n = top_count[v];
  if (n > det_cap) n = det_cap;
  if (i == 0) out.count[v] = n;
  if (i >= det_cap) return;
  const long long o = (long long)v * det_cap + i;
  if (i >= n) {
    out.boxes[o] = make_float4(0, 0, 0, 0); out.props[o] = make_float4(0, 0, 0, 0);
    out.scores[o] = 0.f; out.prob_max[o] = 0.f; out.labels[o] = 0; out.prop_idx[o] = 0;
    return;
  }
  unsigned long long key = top_keys[(long long)v * TOPK_MAX + i];
  unsigned flat = (unsigned)(key & 0xffffffffu);
  int p = flat / (C - 1), c = flat % (C - 1) + 1;
  const float* d = head + ((long long)v * cap + p) * ld + C + c * 4;
  float4 pr = props[(long long)v * cap + p];
  float b[4];
  float dd[4] = {d[0], d[1], d[2], d[3]};
  decode_box(dd, pr.x, pr.y, pr.z, pr.w, 10.f, 10.f, 5.f, 5.f, b);
  clip_box(b, (float)image_hw[v * 2], (float)image_hw[v * 2 + 1]);
  const float rh = ratio_hw[v * 2], rw = ratio_hw[v * 2 + 1];
  out.boxes[o] = make_float4(b[0] * rw, b[1] * rh, b[2] * rw, b[3] * rh);
  out.props[o] = make_float4(pr.x * rw, pr.y * rh, pr.z * rw, pr.w * rh);
  out.scores[o] = scores[((long long)v * cap + p) * C + c];
  out.prob_max[o] = prob_max[(long long)v * cap + p];
  out.labels[o] = c;
  out.prop_idx[o] = p;
}

}  // namespace cald
