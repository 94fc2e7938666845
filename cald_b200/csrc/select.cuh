// On-device selection: the end of an active-learning cycle, cald_train.py:439-448 with
// cls_kldiv (cald_train.py:234-271), on the gathered [n][1 + (C-1)] rows without leaving the GPU.
//   arg  = np.argsort(uncertainty)                      ascending: least consistent images first
//   cand = arg[:int(mr * budget)]
//   picked = cls_kldiv(labeled_loader, [cls[i] for i in cand], budget)
//         = candidates whose class vector is all zero (all of them, in candidate order), then -- until `budget` are
//           picked -- the remaining candidates by descending JS(softmax(mean label histogram) || softmax(class vector))
//           (the reference re-evaluates an unchanging divergence in its loop: its histogram update is commented out,
//           cald_train.py:270; torch.argmax returns the first maximum, so ties go to the earlier candidate)
//   new  = subset[arg][picked]
// Three small kernels: a radix select of the int(mr * budget)-th smallest (score, index) key + sort of the candidates,
// one warp per candidate for the divergence in float64, one CTA for the final ordering.  numpy leaves the order of
// equal scores unspecified; here they resolve to the lower pool index.
#pragma once
#include "det.cuh"

namespace cald {

constexpr int SEL_MAX_CAND = 8192;   // candidates (int(mr * budget)) one call can rank; 128 KB of sort keys in smem

__device__ __forceinline__ unsigned long long asc_key_f64(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull);
}

// ---------------------------------------------------------------- candidates: the m smallest (score, index) pairs
// grid = 1, block = 1024.  out_idx[0..m) = pool positions in ascending (score, index) order.
__global__ void __launch_bounds__(1024) select_candidates_kernel(const double* __restrict__ score, int n, int m,
                                                                 int* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char ssm[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(ssm);            // [SEL_MAX_CAND] score keys
  int* sidx = reinterpret_cast<int*>(ssm + (size_t)SEL_MAX_CAND * 8);               // [SEL_MAX_CAND]
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix, s_mask;
  __shared__ int s_need, s_cnt, s_tie_need;
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = m; s_cnt = 0; }
  __syncthreads();
  // MSB-first radix select of the m-th smallest score key (8 passes of 8 bits); ties share a key
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix, mask = s_mask;
    for (int i = tid; i < n; i += blockDim.x) {
      const unsigned long long k = asc_key_f64(score[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, need = s_need, d = 255;
      for (int b = 0; b < 256; ++b) {
        if (cum + (int)hist[b] >= need) { d = b; break; }
        cum += (int)hist[b];
      }
      s_need = need - cum;
      s_prefix = prefix | ((unsigned long long)d << shift);
      s_mask = mask | (0xffull << shift);
    }
    __syncthreads();
  }
  // keys below the threshold all qualify; of the keys equal to it the s_need lowest pool indices do
  const unsigned long long thr = s_prefix;
  if (tid == 0) s_tie_need = s_need;
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    const unsigned long long k = asc_key_f64(score[i]);
    if (k < thr) {
      const int slot = atomicAdd(&s_cnt, 1);
      skey[slot] = k; sidx[slot] = i;
    }
  }
  __syncthreads();
  // ties at the threshold, lowest index first: a serial sweep by one thread keeps it simple (rare, and n is ~1e5)
  if (tid == 0) {
    int need = s_tie_need, slot = s_cnt;
    for (int i = 0; i < n && need > 0; ++i) {
      if (asc_key_f64(score[i]) == thr) { skey[slot] = thr; sidx[slot] = i; ++slot; --need; }
    }
    s_cnt = slot;
  }
  __syncthreads();
  const int cnt = s_cnt;   // == m
  // sort by (score key, index): pack the rank of each candidate by counting (cnt <= 8192: O(cnt^2 / 1024) compares)
  for (int i = tid; i < cnt; i += blockDim.x) {
    const unsigned long long k = skey[i];
    const int id = sidx[i];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) {
      const unsigned long long kj = skey[j];
      rank += (kj < k) || (kj == k && sidx[j] < id);
    }
    out_idx[rank] = id;
  }
}

// ---------------------------------------------------------------- divergence of every candidate (one warp each)
// js[c] = JS(p || q) as cald_train.py:262-269 computes it in float64; zero[c] = class vector is all zero.
// uniform (cald_train.py:254-261): p = softmax(mean_hist + corr), q = softmax(ones) = 1 / c1.
__global__ void select_js_kernel(const double* __restrict__ cls, const int* __restrict__ cand, int m, int c1,
                                 const double* __restrict__ mean_hist, int uniform, double* __restrict__ js,
                                 int* __restrict__ zero) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= m) return;
  const double* row = cls + (long long)cand[c] * c1;
  auto wsum = [](double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
  auto wmax = [](double v) { for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; };
  double rs = 0.0, mp = -INFINITY, mq = -INFINITY;
  for (int k = lane; k < c1; k += 32) {
    rs += row[k];
    mp = fmax(mp, uniform ? mean_hist[k] + row[k] : mean_hist[k]);
    mq = fmax(mq, uniform ? 1.0 : row[k]);
  }
  rs = wsum(rs); mp = wmax(mp); mq = wmax(mq);
  double sp = 0.0, sq = 0.0;
  for (int k = lane; k < c1; k += 32) {
    sp += exp((uniform ? mean_hist[k] + row[k] : mean_hist[k]) - mp);
    sq += exp((uniform ? 1.0 : row[k]) - mq);
  }
  sp = wsum(sp); sq = wsum(sq);
  double a = 0.0, b = 0.0;
  for (int k = lane; k < c1; k += 32) {
    const double p = exp((uniform ? mean_hist[k] + row[k] : mean_hist[k]) - mp) / sp;
    const double q = exp((uniform ? 1.0 : row[k]) - mq) / sq;
    const double lm = log((p + q) / 2);
    // torch.nn.KLDivLoss(reduction='none')(log_mean, t) = xlogy(t, t) - t * log_mean
    a += (p > 0.0 ? p * log(p) : 0.0) - p * lm;
    b += (q > 0.0 ? q * log(q) : 0.0) - q * lm;
  }
  a = wsum(a); b = wsum(b);
  if (lane == 0) { js[c] = a / 2 + b / 2; zero[c] = (rs == 0.0) ? 1 : 0; }
}

// ---------------------------------------------------------------- final order
// grid = 1, block = 1024.  out_pick[0..*n_pick) = positions in the CANDIDATE list (what cls_kldiv returns).
__global__ void __launch_bounds__(1024) select_pick_kernel(const double* __restrict__ js, const int* __restrict__ zero,
                                                           int m, int budget, int uniform, int* __restrict__ out_pick,
                                                           int* __restrict__ n_pick) {
  __shared__ int s_nz;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int nz = 0;
    for (int c = 0; c < m; ++c) if (zero[c]) out_pick[nz++] = c;   // np.where(sum == 0)[0]: candidate order
    s_nz = nz;
  }
  __syncthreads();
  const int nz = s_nz;
  if (nz >= budget) { if (tid == 0) *n_pick = nz; return; }   // cald_train.py:248-249 returns all of them
  // remaining picks: non-zero candidates by descending js (ascending in uniform mode), earlier candidate first on ties
  for (int c = tid; c < m; c += blockDim.x) {
    if (zero[c]) continue;
    const double v = js[c];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      if (zero[j]) continue;
      const double w = js[j];
      const bool before = uniform ? (w < v || (w == v && j < c)) : (w > v || (w == v && j < c));
      rank += before;
    }
    if (nz + rank < budget) out_pick[nz + rank] = c;
  }
  if (tid == 0) *n_pick = (m - nz) >= (budget - nz) ? budget : m;
}

}  // namespace cald
