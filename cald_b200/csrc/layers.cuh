// Elementwise / data-movement kernels around the conv engine (all HBM-bound):
// split <-> fp32 conversion, 2x subsample (LastLevelMaxPool), 3x3/2 max-pool.
// 16-byte vector accesses, channels innermost (NHWC).
#pragma once
#include "conv_host.cuh"

namespace cald {

// ---------------------------------------------------------------- device arena
// First-fit free list over one cudaMalloc'd slab.  All kernels run on one stream,
// so host-side free/reuse is stream-ordered and the address sequence is deterministic.
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0;
  struct Blk { size_t off, size; bool used; };
  std::vector<Blk> blks;
  size_t peak = 0;
  void init(size_t bytes) {
    CALD_CUDA_CHECK(cudaMalloc((void**)&base, bytes));
    cap = bytes;
    blks.clear();
    blks.push_back({0, bytes, false});
  }
  void destroy() { if (base) cudaFree(base); base = nullptr; }
  void* alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~(size_t)1023;
    if (bytes == 0) bytes = 1024;
    for (size_t i = 0; i < blks.size(); ++i) {
      if (!blks[i].used && blks[i].size >= bytes) {
        if (blks[i].size > bytes) {
          Blk rest{blks[i].off + bytes, blks[i].size - bytes, false};
          blks[i].size = bytes;
          blks.insert(blks.begin() + i + 1, rest);
        }
        blks[i].used = true;
        if (blks[i].off + bytes > peak) peak = blks[i].off + bytes;
        return base + blks[i].off;
      }
    }
    char b[128];
    snprintf(b, sizeof(b), "device arena exhausted (cap %zu MiB, request %zu MiB)", cap >> 20, bytes >> 20);
    throw std::runtime_error(b);
  }
  void free(void* p) {
    if (!p) return;
    size_t off = (uint8_t*)p - base;
    for (size_t i = 0; i < blks.size(); ++i) {
      if (blks[i].off == off && blks[i].used) {
        blks[i].used = false;
        if (i + 1 < blks.size() && !blks[i + 1].used) { blks[i].size += blks[i + 1].size; blks.erase(blks.begin() + i + 1); }
        if (i > 0 && !blks[i - 1].used) { blks[i - 1].size += blks[i].size; blks.erase(blks.begin() + i); }
        return;
      }
    }
    throw std::runtime_error("arena: bad free");
  }
  void reset() { blks.clear(); blks.push_back({0, cap, false}); }
};

inline Act alloc_act(Arena& a, int n, int h, int w, int c, bool split) {
  Act t;
  t.n = n; t.h = h; t.w = w; t.c = c; t.split = split;
  t.hi = (pl16*)a.alloc(t.bytes());
  return t;
}
inline void free_act(Arena& a, Act& t) { a.free(t.hi); t.hi = nullptr; }

// ---------------------------------------------------------------- conversions
__global__ void f32_to_split_kernel(const float* __restrict__ in, pl16* __restrict__ hi, pl16* __restrict__ lo,
                                    long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(in + i);
    pl16 h[4], l[4];
    split_pl(v.x, h[0], l[0]); split_pl(v.y, h[1], l[1]);
    split_pl(v.z, h[2], l[2]); split_pl(v.w, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi + i) = make_uint2(pack_pl16x2(h[0], h[1]), pack_pl16x2(h[2], h[3]));
    if (lo) *reinterpret_cast<uint2*>(lo + i) = make_uint2(pack_pl16x2(l[0], l[1]), pack_pl16x2(l[2], l[3]));
  } else {
    for (; i < n; ++i) {
      pl16 h, l;
      split_pl(in[i], h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  }
}
__global__ void split_to_f32_kernel(const pl16* __restrict__ hi, const pl16* __restrict__ lo, float* __restrict__ out,
                                    long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = lo ? join_pl(hi[i], lo[i]) : pl16_to_float(hi[i]);
}
inline void f32_to_split(const float* in, Act& t, cudaStream_t st) {
  long long n = (long long)t.plane_elems();
  f32_to_split_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(in, t.hi, t.lo(), n);
  CALD_CUDA_CHECK(cudaGetLastError());
}
inline void split_to_f32(const Act& t, float* out, cudaStream_t st) {
  long long n = (long long)t.plane_elems();
  split_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t.hi, t.lo(), out, n);
  CALD_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------- 2x subsample (8 channels per thread)
// out[i][y][x][c] = in[i][2y][2x][c]   (1x1 stride-2 convs, LastLevelMaxPool k=1 s=2)
__global__ void subsample2_kernel(const pl16* __restrict__ in, pl16* __restrict__ out, int n, int h, int w, int c,
                                  int h2, int w2) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = c / 8;
  long long total = (long long)n * h2 * w2 * cg;
  if (gid >= total) return;
  int g = (int)(gid % cg);
  long long r = gid / cg;
  int x = (int)(r % w2); r /= w2;
  int y = (int)(r % h2);
  int i = (int)(r / h2);
  *reinterpret_cast<uint4*>(out + gid * 8) =
      *reinterpret_cast<const uint4*>(in + (((long long)i * h + 2 * y) * w + 2 * x) * c + g * 8);
}
inline Act subsample2(Arena& a, const Act& in, cudaStream_t st) {
  Act o = alloc_act(a, in.n, (in.h + 1) / 2, (in.w + 1) / 2, in.c, in.split);
  long long total = (long long)in.n * o.h * o.w * (in.c / 8);
  for (int pl = 0; pl < (in.split ? 2 : 1); ++pl)
    subsample2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        in.hi + pl * in.plane_elems(), o.hi + pl * o.plane_elems(), in.n, in.h, in.w, in.c, o.h, o.w);
  CALD_CUDA_CHECK(cudaGetLastError());
  return o;
}

// ---------------------------------------------------------------- 3x3 / stride 2 / pad 1 max-pool (tv resnet.py maxpool)
__global__ void maxpool3x3s2_kernel(const pl16* __restrict__ ihi, const pl16* __restrict__ ilo, pl16* __restrict__ ohi,
                                    pl16* __restrict__ olo, int n, int h, int w, int c, int ho, int wo) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = c / 8;
  long long total = (long long)n * ho * wo * cg;
  if (gid >= total) return;
  int g = (int)(gid % cg);
  long long r = gid / cg;
  int x = (int)(r % wo); r /= wo;
  int y = (int)(r % ho);
  int i = (int)(r / ho);
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    int sy = 2 * y + dy;
    if (sy < 0 || sy >= h) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      int sx = 2 * x + dx;
      if (sx < 0 || sx >= w) continue;
      long long off = (((long long)i * h + sy) * w + sx) * c + g * 8;
      uint4 vh = *reinterpret_cast<const uint4*>(ihi + off);
      const pl16* ph = reinterpret_cast<const pl16*>(&vh);
      if (ilo) {
        uint4 vl = *reinterpret_cast<const uint4*>(ilo + off);
        const pl16* pl = reinterpret_cast<const pl16*>(&vl);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], join_pl(ph[k], pl[k]));
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], pl16_to_float(ph[k]));
      }
    }
  }
  pl16 hh[8], ll[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) split_pl(m[k], hh[k], ll[k]);
  *reinterpret_cast<uint4*>(ohi + gid * 8) = make_uint4(pack_pl16x2(hh[0], hh[1]), pack_pl16x2(hh[2], hh[3]),
                                                       pack_pl16x2(hh[4], hh[5]), pack_pl16x2(hh[6], hh[7]));
  if (olo)
    *reinterpret_cast<uint4*>(olo + gid * 8) = make_uint4(pack_pl16x2(ll[0], ll[1]), pack_pl16x2(ll[2], ll[3]),
                                                         pack_pl16x2(ll[4], ll[5]), pack_pl16x2(ll[6], ll[7]));
}
inline Act maxpool3x3s2(Arena& a, const Act& in, cudaStream_t st) {
  Act o = alloc_act(a, in.n, (in.h - 1) / 2 + 1, (in.w - 1) / 2 + 1, in.c, in.split);
  long long total = (long long)in.n * o.h * o.w * (in.c / 8);
  maxpool3x3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in.hi, in.lo(), o.hi, o.lo(), in.n, in.h, in.w,
                                                                      in.c, o.h, o.w);
  CALD_CUDA_CHECK(cudaGetLastError());
  return o;
}

}  // namespace cald
