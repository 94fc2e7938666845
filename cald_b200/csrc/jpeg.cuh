// Pool ingest: baseline JPEG files -> u8 HWC RGB images in HBM, entirely on the device.
//
// The reference reads its pools with PIL on DataLoader workers (detection/voc_utils.py:52-58,
// detection/coco_utils.py:209-220: Image.open(path).convert('RGB')).  Here the host only walks the marker segments
// (a few hundred bytes per file); the compressed scan travels over the link (10-20x fewer bytes than decoded pixels)
// and three kernels produce exactly the pixels Pillow's libjpeg produces -- bit for bit, so the pinned parity of the
// scoring path carries over to files:
//   jpeg_huffman_kernel : one warp per image, lane 0 walks the entropy-coded segment (sequential by nature; a chunk's
//                         images decode side by side) -> int16 coefficients, natural order            (jdhuff.c)
//   jpeg_idct_kernel    : one thread per 8x8 block: dequantise + 13-bit fixed-point LLM inverse DCT (JDCT_ISLOW,
//                         libjpeg's and Pillow's default)                                               (jidctint.c)
//   jpeg_rgb_kernel     : one thread per pixel: "fancy" triangle-filter chroma upsampling for 4:2:2 / 4:2:0 with
//                         libjpeg's edge rules, then the 16-bit fixed-point YCbCr -> RGB tables  (jdsample.c, jdcolor.c)
// Supported: 8-bit sequential Huffman (SOF0 / SOF1), one interleaved scan, grayscale or YCbCr with 4:4:4, 4:2:2 or
// 4:2:0 sampling, restart intervals.  Anything else (progressive, CMYK, ...) is refused with a message; nothing is
// decoded on the CPU.  The CPU restatement that pins this against Pillow is oracle/jpeg_oracle.py.
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace cald {

constexpr int JPEG_LOOK_BITS = 9;

struct JpegHuff {                 // one Huffman table in decoder form
  uint16_t look[1 << JPEG_LOOK_BITS];  // (length << 8) | symbol for codes of <= 9 bits, 0 = longer code
  int maxcode[18];                // largest code of each length (-1 = none); [17] = sentinel
  int valoff[17];                 // huffval index of the first code of each length minus that code
  uint8_t huffval[256];
};

struct JpegComp {
  int h, v;                       // sampling factors
  int tq, td, ta;                 // quantisation / DC / AC table selectors
  int blocks_w, blocks_h;         // coefficient blocks (MCU-padded)
  int dw, dh;                     // downsampled_width / downsampled_height: the real samples (jdmaster.c)
  long long coef_off;             // int16 offset of this component's coefficients in the chunk's coefficient buffer
  long long plane_off;            // byte offset of this component's sample plane (pitch = blocks_w * 8)
};

struct JpegImage {
  int width, height, ncomp;
  int hmax, vmax, mcux, mcuy;
  int restart;                    // MCUs per restart interval, 0 = none
  long long scan_off, scan_len;   // entropy-coded bytes in the chunk's byte buffer
  long long out_off;              // byte offset of the RGB output in the chunk's image slab
  JpegComp comp[3];
  uint16_t quant[4][64];          // natural order
  int huff_dc[2], huff_ac[2];     // index into the chunk's table array, -1 = undefined
};

// ---------------------------------------------------------------- host: marker segments
inline void jpeg_build_huff(const uint8_t* counts, const uint8_t* vals, int nvals, JpegHuff& t) {
  memset(&t, 0, sizeof(t));
  memcpy(t.huffval, vals, (size_t)nvals);
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    t.valoff[l] = k - code;
    for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code) {
      if (l <= JPEG_LOOK_BITS) {
        const int first = code << (JPEG_LOOK_BITS - l), n = 1 << (JPEG_LOOK_BITS - l);
        for (int j = 0; j < n; ++j) t.look[first + j] = (uint16_t)((l << 8) | vals[k]);
      }
    }
    t.maxcode[l] = counts[l - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t.maxcode[17] = 0x7fffffff;
}

// Parses one file.  Fills `im` (table indices refer to `tables`, which is appended to) and the byte range of the scan.
inline void jpeg_parse(const uint8_t* d, size_t n, JpegImage& im, std::vector<JpegHuff>& tables, size_t& scan_begin,
                       size_t& scan_end) {
  auto fail = [](const std::string& m) { throw std::runtime_error("JPEG: " + m); };
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) fail("not a JPEG file (no SOI marker)");
  memset(&im, 0, sizeof(im));
  int table_of[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}};
  bool have_q[4] = {false, false, false, false}, have_frame = false;
  int comp_id[3] = {0, 0, 0};
  int adobe_transform = -1;
  size_t pos = 2;
  for (;;) {
    if (pos + 4 > n) fail("truncated file (no SOS marker)");
    if (d[pos] != 0xFF) fail("marker expected");
    while (pos + 1 < n && d[pos + 1] == 0xFF) ++pos;
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (m == 0xD9) fail("EOI before SOS");
    if (pos + 2 > n) fail("truncated segment");
    const size_t len = ((size_t)d[pos] << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) fail("bad segment length");
    const uint8_t* s = d + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) fail("bad DQT");
        static const int zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                   15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        for (int k = 0; k < 64; ++k) {
          im.quant[tq][zz[k]] = pq ? (uint16_t)((s[i + 2 * k] << 8) | s[i + 2 * k + 1]) : s[i + k];
        }
        i += pq ? 128 : 64;
        have_q[tq] = true;
      }
    } else if (m == 0xC4) {
      size_t i = 0;
      while (i < sl) {
        if (i + 17 > sl) fail("bad DHT");
        const int tc = s[i] >> 4, th = s[i] & 15;
        int nv = 0;
        for (int k = 0; k < 16; ++k) nv += s[i + 1 + k];
        if (tc > 1 || th > 3 || nv > 256 || i + 17 + nv > sl) fail("bad DHT");
        JpegHuff t;
        jpeg_build_huff(s + i + 1, s + i + 17, nv, t);
        table_of[tc][th] = (int)tables.size();
        tables.push_back(t);
        i += 17 + (size_t)nv;
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || s[0] != 8) fail("only 8-bit samples are supported");
      im.height = (s[1] << 8) | s[2];
      im.width = (s[3] << 8) | s[4];
      im.ncomp = s[5];
      if (im.width <= 0 || im.height <= 0) fail("empty image");
      if (im.ncomp != 1 && im.ncomp != 3) fail("only grayscale and YCbCr files are supported (CMYK / 4-component refused)");
      if (sl < 6 + 3 * (size_t)im.ncomp) fail("bad SOF");
      for (int c = 0; c < im.ncomp; ++c) {
        comp_id[c] = s[6 + 3 * c];
        im.comp[c].h = s[7 + 3 * c] >> 4;
        im.comp[c].v = s[7 + 3 * c] & 15;
        im.comp[c].tq = s[8 + 3 * c];
        if (im.comp[c].tq > 3) fail("bad quantisation table selector");
      }
      have_frame = true;
    } else if (m == 0xC2) {
      fail("progressive JPEG (SOF2) is not supported by the device decoder; re-encode as baseline");
    } else if ((m >= 0xC3 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      fail("unsupported JPEG process (lossless / hierarchical / arithmetic)");
    } else if (m == 0xDD) {
      if (sl < 2) fail("bad DRI");
      im.restart = (s[0] << 8) | s[1];
    } else if (m == 0xEE && sl >= 12 && memcmp(s, "Adobe", 5) == 0) {
      adobe_transform = s[11];
    } else if (m == 0xDA) {
      if (!have_frame) fail("SOS before SOF");
      if (sl < 1 || s[0] != im.ncomp || sl < 1 + 2 * (size_t)im.ncomp + 3)
        fail("multi-scan (non-interleaved) files are not supported");
      for (int k = 0; k < im.ncomp; ++k) {
        const int id = s[1 + 2 * k], td = s[2 + 2 * k] >> 4, ta = s[2 + 2 * k] & 15;
        int c = -1;
        for (int q = 0; q < im.ncomp; ++q) if (comp_id[q] == id) c = q;
        if (c < 0 || td > 1 || ta > 1) fail("bad SOS component (Huffman table selectors above 1 are not supported)");
        im.comp[c].td = td; im.comp[c].ta = ta;
      }
      if (im.ncomp == 3 && adobe_transform == 0) fail("Adobe RGB-coded JPEG (transform 0) is not supported");
      // geometry
      if (im.ncomp == 1) { im.comp[0].h = im.comp[0].v = 1; }
      im.hmax = im.vmax = 1;
      for (int c = 0; c < im.ncomp; ++c) { im.hmax = std::max(im.hmax, im.comp[c].h); im.vmax = std::max(im.vmax, im.comp[c].v); }
      if (im.ncomp == 3) {
        const bool ok = im.comp[1].h == 1 && im.comp[1].v == 1 && im.comp[2].h == 1 && im.comp[2].v == 1 &&
                        ((im.comp[0].h == 1 && im.comp[0].v == 1) || (im.comp[0].h == 2 && im.comp[0].v == 1) ||
                         (im.comp[0].h == 2 && im.comp[0].v == 2));
        if (!ok) fail("unsupported chroma subsampling (4:4:4, 4:2:2 and 4:2:0 are decoded)");
      }
      im.mcux = (im.width + 8 * im.hmax - 1) / (8 * im.hmax);
      im.mcuy = (im.height + 8 * im.vmax - 1) / (8 * im.vmax);
      for (int c = 0; c < im.ncomp; ++c) {
        JpegComp& cp = im.comp[c];
        cp.blocks_w = im.mcux * cp.h;
        cp.blocks_h = im.mcuy * cp.v;
        cp.dw = (im.width * cp.h + im.hmax - 1) / im.hmax;
        cp.dh = (im.height * cp.v + im.vmax - 1) / im.vmax;
        if (!have_q[cp.tq]) fail("missing quantisation table");
        if (table_of[0][cp.td] < 0 || table_of[1][cp.ta] < 0) fail("missing Huffman table");
      }
      for (int k = 0; k < 2; ++k) { im.huff_dc[k] = table_of[0][k]; im.huff_ac[k] = table_of[1][k]; }
      scan_begin = pos + len;
      // the entropy-coded segment ends at the first marker that is neither a stuffed zero nor RSTn
      size_t q = scan_begin;
      while (q + 1 < n && !(d[q] == 0xFF && d[q + 1] != 0x00 && !(d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7))) ++q;
      scan_end = (q + 1 < n) ? q : n;
      return;
    }
    pos += len;
  }
}

// ---------------------------------------------------------------- device: entropy decoding
struct JpegBits {
  const uint8_t* p;
  const uint8_t* end;
  unsigned long long acc;   // bits are consumed from the top
  int n;                    // valid bits in acc
  // source bytes are fetched eight at a time (one aligned 64-bit load per eight bytes instead of one load per byte:
  // the walk is a single dependent instruction chain, so every load's latency is exposed)
  unsigned long long win;   // the aligned 8-byte window that contains *p
  const uint8_t* win_at;    // its address (null = none loaded)
  __host__ __device__ __forceinline__ unsigned byte_at(const uint8_t* q) {
#ifdef __CUDA_ARCH__
    const uint8_t* a = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)7);
    if (a != win_at) {
      win = *reinterpret_cast<const unsigned long long*>(a);   // the device scan buffer is padded to 16 bytes
      win_at = a;
    }
    return (unsigned)(win >> (8 * (unsigned)(q - a))) & 0xffu;
#else
    return *q;                                                   // host: the caller's file buffer, no padding
#endif
  }
  __host__ __device__ __forceinline__ void fill() {
    while (n <= 56) {
      unsigned b = 0;
      if (p < end) {
        b = byte_at(p);
        if (b == 0xFF) {
          const unsigned nx = (p + 1 < end) ? byte_at(p + 1) : 0xD9u;
          if (nx == 0) p += 2;       // stuffed zero
          else b = 0;                // a marker: feed zeros and stay in front of it (jdhuff.c)
        } else {
          ++p;
        }
      }
      acc |= (unsigned long long)b << (56 - n);
      n += 8;
    }
  }
  __host__ __device__ __forceinline__ unsigned peek(int k) { return (unsigned)(acc >> (64 - k)); }
  __host__ __device__ __forceinline__ void skip(int k) { acc <<= k; n -= k; }
  __host__ __device__ __forceinline__ int receive_extend(int s) {
    if (n < s) fill();
    const int v = (int)peek(s);
    skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
  __host__ __device__ __forceinline__ int decode(const JpegHuff& t) {
    if (n < 16) fill();
    const unsigned l9 = t.look[peek(JPEG_LOOK_BITS)];
    if (l9) { skip((int)(l9 >> 8)); return (int)(l9 & 0xff); }
    int l = JPEG_LOOK_BITS + 1;
    int code = (int)peek(l);
    while (l <= 16 && code > t.maxcode[l]) { ++l; code = (int)peek(l); }
    if (l > 16) { skip(16); return 0; }       // corrupt data: libjpeg warns and continues with a zero symbol
    skip(l);
    return t.huffval[(code + t.valoff[l]) & 0xff];
  }
  __host__ __device__ void restart() {                  // byte-align and step over the RSTn marker
    acc = 0; n = 0;
    while (p + 1 < end && !(byte_at(p) == 0xFF && byte_at(p + 1) >= 0xD0 && byte_at(p + 1) <= 0xD7)) ++p;
    if (p + 1 < end) p += 2;
  }
};

__constant__ int c_jpeg_zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                      15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__host__ __device__ inline int jpeg_max0(int v) { return v < 0 ? 0 : v; }

// The entropy-coded segment of one image -> its non-zero coefficients (natural order), jdhuff.c.  Shared by the device
// kernel and the host path (jpeg_walk_host below).
__host__ __device__ inline void jpeg_walk(const JpegImage& im, const JpegHuff* tables, const uint8_t* scan,
                                          short* coef) {
#ifdef __CUDA_ARCH__
  const int* zigzag = c_jpeg_zigzag;
#else
  static const int zz_host[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                  41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                  15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  const int* zigzag = zz_host;
#endif
  const JpegHuff* t_dc[3];
  const JpegHuff* t_ac[3];
  int s_h[3], s_v[3], s_bw[3];
  long long s_off[3];
  for (int c = 0; c < 3; ++c) {
    const JpegComp& cp = im.comp[c < im.ncomp ? c : 0];
    s_h[c] = cp.h; s_v[c] = cp.v; s_bw[c] = cp.blocks_w; s_off[c] = cp.coef_off;
    t_dc[c] = tables + jpeg_max0(im.huff_dc[cp.td & 1]);
    t_ac[c] = tables + jpeg_max0(im.huff_ac[cp.ta & 1]);
  }
  JpegBits br;
  br.p = scan;
  br.end = br.p + im.scan_len;
  br.acc = 0; br.n = 0;
  br.win = 0; br.win_at = nullptr;
  int pred[3] = {0, 0, 0};
  const int ncomp = im.ncomp, mcux = im.mcux, mcuy = im.mcuy, restart = im.restart;
  int left = restart;
  for (int my = 0; my < mcuy; ++my) {
    for (int mx = 0; mx < mcux; ++mx) {
      if (restart) {
        if (left == 0) { br.restart(); pred[0] = pred[1] = pred[2] = 0; left = restart; }
        --left;
      }
      for (int c = 0; c < ncomp; ++c) {
        const JpegHuff& dc = *t_dc[c];
        const JpegHuff& ac = *t_ac[c];
        const int ch = s_h[c], cv = s_v[c], bw = s_bw[c];
        short* cbase = coef + s_off[c];
        for (int by = 0; by < cv; ++by) {
          for (int bx = 0; bx < ch; ++bx) {
            short* blk = cbase + ((long long)(my * cv + by) * bw + (mx * ch + bx)) * 64;
            const int t = br.decode(dc) & 15;
            if (t) pred[c] += br.receive_extend(t);
            blk[0] = (short)pred[c];
            for (int k = 1; k < 64;) {
              const int rs = br.decode(ac);
              const int r = rs >> 4, s = rs & 15;
              if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
              }
              k += r;
              const int v = br.receive_extend(s);
              if (k < 64) blk[zigzag[k]] = (short)v;
              ++k;
            }
          }
        }
      }
    }
  }
}

// grid = JPEG_WALK_BLOCKS one-warp blocks that hand the chunk's images out among themselves; coef must be zeroed
// beforehand (only non-zero coefficients are written); sched = {next image, per-SM claim flags}, zeroed per launch.
//
// The kernel runs on the copy stream WHILE the persistent conv kernels hold every SM, so it has to fit beside them and
// must never keep a conv CTA from launching:
//  * no shared memory: the one-CTA conv kernel leaves exactly the 1 KB a block reserves (measured with
//    tools/overlap_probe.py: a block with 0 B co-resides, one with 4 KB delays the conv train by the block's lifetime);
//    the decoder tables (1.4 KB each) are read through L1 instead;
//  * at most ONE walking block per SM: beside a CTA-pair conv kernel (202 KB) the block scheduler can stack several
//    blocks on one SM, and the next one-CTA conv launch then finds no room there for ~50 ms -- with statically assigned
//    tiles that stalls the whole launch (measured: scoring from files 15 % slower than from pixels).  Every block
//    therefore claims its SM (%smid) first; a block that finds the SM taken exits at once, the owners take images
//    from a shared counter until none is left.
constexpr int JPEG_WALK_BLOCKS = 2 * 148;
constexpr int JPEG_SCHED_INTS = 1 + 256;
__global__ void __launch_bounds__(32) jpeg_huffman_kernel(const JpegImage* __restrict__ imgs,
                                                          const JpegHuff* __restrict__ tables,
                                                          const uint8_t* __restrict__ bytes, short* __restrict__ coef,
                                                          int n_images, int* __restrict__ sched) {
  if (threadIdx.x != 0) return;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (atomicCAS(&sched[1 + (smid & 255u)], 0, 1) != 0) return;
  for (;;) {
  const int img_i = atomicAdd(&sched[0], 1);
  if (img_i >= n_images) return;
  jpeg_walk(imgs[img_i], tables, bytes + imgs[img_i].scan_off, coef);
  }  // next image
}

// ---------------------------------------------------------------- device: inverse DCT (jidctint.c, JDCT_ISLOW)
#define JFIX_0_298631336 2446
#define JFIX_0_390180644 3196
#define JFIX_0_541196100 4433
#define JFIX_0_765366865 6270
#define JFIX_0_899976223 7373
#define JFIX_1_175875602 9633
#define JFIX_1_501321110 12299
#define JFIX_1_847759065 15137
#define JFIX_1_961570560 16069
#define JFIX_2_053119869 16819
#define JFIX_2_562915447 20995
#define JFIX_3_072711026 25172

__device__ __forceinline__ void jpeg_idct_1d(const int* in, int* out, int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * JFIX_0_541196100;
  int tmp2 = z1 + z3 * (-JFIX_1_847759065);
  int tmp3 = z1 + z2 * JFIX_0_765366865;
  z2 = in[0]; z3 = in[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * JFIX_1_175875602;
  tmp0 *= JFIX_0_298631336; tmp1 *= JFIX_2_053119869; tmp2 *= JFIX_3_072711026; tmp3 *= JFIX_1_501321110;
  z1 *= -JFIX_0_899976223; z2 *= -JFIX_2_562915447; z3 *= -JFIX_1_961570560; z4 *= -JFIX_0_390180644;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  out[0] = (tmp10 + tmp3 + rnd) >> shift; out[7] = (tmp10 - tmp3 + rnd) >> shift;
  out[1] = (tmp11 + tmp2 + rnd) >> shift; out[6] = (tmp11 - tmp2 + rnd) >> shift;
  out[2] = (tmp12 + tmp1 + rnd) >> shift; out[5] = (tmp12 - tmp1 + rnd) >> shift;
  out[3] = (tmp13 + tmp0 + rnd) >> shift; out[4] = (tmp13 - tmp0 + rnd) >> shift;
}

// grid = (ceil(max blocks / 128), images * 3), block = 128: one thread per 8x8 block of one component
__global__ void jpeg_idct_kernel(const JpegImage* __restrict__ imgs, const short* __restrict__ coef,
                                 uint8_t* __restrict__ planes) {
  const JpegImage& im = imgs[blockIdx.y / 3];
  const int c = blockIdx.y % 3;
  if (c >= im.ncomp) return;
  const JpegComp& cp = im.comp[c];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cp.blocks_w * cp.blocks_h) return;
  const short* src = coef + cp.coef_off + (long long)b * 64;
  const uint16_t* q = im.quant[cp.tq];
  int ws[64];
  // pass 1: columns
#pragma unroll
  for (int col = 0; col < 8; ++col) {
    int in[8], out[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) in[k] = (int)src[k * 8 + col] * (int)q[k * 8 + col];
    jpeg_idct_1d(in, out, 13 - 2);
#pragma unroll
    for (int k = 0; k < 8; ++k) ws[k * 8 + col] = out[k];
  }
  // pass 2: rows -> samples through the range-limit table (10-bit wrap, then clamp(x + 128))
  const int by = b / cp.blocks_w, bx = b - by * cp.blocks_w;
  uint8_t* dst = planes + cp.plane_off + ((long long)by * 8) * (cp.blocks_w * 8) + bx * 8;
#pragma unroll
  for (int row = 0; row < 8; ++row) {
    int out[8];
    jpeg_idct_1d(&ws[row * 8], out, 13 + 2 + 3);
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int x = out[k] & 1023;
      if (x >= 512) x -= 1024;
      x += 128;
      x = x < 0 ? 0 : (x > 255 ? 255 : x);
      if (k < 4) w0 |= (uint32_t)x << (8 * k); else w1 |= (uint32_t)x << (8 * (k - 4));
    }
    *reinterpret_cast<uint2*>(dst + (long long)row * (cp.blocks_w * 8)) = make_uint2(w0, w1);
  }
}

// ---------------------------------------------------------------- device: fancy upsampling + colour conversion
__device__ __forceinline__ int jpeg_chroma(const uint8_t* __restrict__ p, int pitch, int dw, int dh, int hs, int vs,
                                           int X, int Y) {
  if (hs == 1 && vs == 1) return p[(long long)Y * pitch + X];
  const int cx = X >> 1;
  if (vs == 1) {   // h2v1_fancy_upsample
    const uint8_t* r = p + (long long)Y * pitch;
    const int v = r[cx];
    if (X & 1) return cx == dw - 1 ? v : (3 * v + r[cx + 1] + 2) >> 2;
    return cx == 0 ? v : (3 * v + r[cx - 1] + 1) >> 2;
  }
  // h2v2_fancy_upsample: nearer row weight 3, further row weight 1 (the row above for even output rows, below for
  // odd ones; at the image edge the context row is the edge row itself, jdmainct.c)
  const int cy = Y >> 1;
  int fy = (Y & 1) ? cy + 1 : cy - 1;
  fy = fy < 0 ? 0 : (fy > dh - 1 ? dh - 1 : fy);
  const uint8_t* near = p + (long long)cy * pitch;
  const uint8_t* far = p + (long long)fy * pitch;
  const int s = near[cx] * 3 + far[cx];
  if (X & 1) {
    if (cx == dw - 1) return (s * 4 + 7) >> 4;
    return (s * 3 + (near[cx + 1] * 3 + far[cx + 1]) + 7) >> 4;
  }
  if (cx == 0) return (s * 4 + 8) >> 4;
  return (s * 3 + (near[cx - 1] * 3 + far[cx - 1]) + 8) >> 4;
}

// grid = (ceil(max W / 128), max H, images), block = 128
__global__ void jpeg_rgb_kernel(const JpegImage* __restrict__ imgs, const uint8_t* __restrict__ planes,
                                uint8_t* __restrict__ out) {
  const JpegImage& im = imgs[blockIdx.z];
  const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y;
  if (X >= im.width || Y >= im.height) return;
  const JpegComp& c0 = im.comp[0];
  const int y = planes[c0.plane_off + (long long)Y * (c0.blocks_w * 8) + X];
  uint8_t* o = out + im.out_off + ((long long)Y * im.width + X) * 3;
  if (im.ncomp == 1) { o[0] = o[1] = o[2] = (uint8_t)y; return; }
  const int hs = im.hmax / im.comp[1].h, vs = im.vmax / im.comp[1].v;
  const int cb = jpeg_chroma(planes + im.comp[1].plane_off, im.comp[1].blocks_w * 8, im.comp[1].dw, im.comp[1].dh, hs, vs, X, Y) - 128;
  const int cr = jpeg_chroma(planes + im.comp[2].plane_off, im.comp[2].blocks_w * 8, im.comp[2].dw, im.comp[2].dh, hs, vs, X, Y) - 128;
  // jdcolor.c build_ycc_rgb_table: FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 32768, arithmetic right shifts
  const int r = y + ((91881 * cr + 32768) >> 16);
  const int g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
  const int b = y + ((116130 * cb + 32768) >> 16);
  o[0] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
  o[1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
  o[2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
}

}  // namespace cald
