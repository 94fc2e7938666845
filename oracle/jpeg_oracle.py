"""TEST INFRASTRUCTURE ONLY -- baseline JPEG decoder restating what ``PIL.Image.open(path).convert('RGB')`` does
for the pool images (reference: detection/voc_utils.py:52-58, detection/coco_utils.py:209-220 read the pool with PIL).

Pillow decodes through libjpeg(-turbo), a third-party dependency that is not under /root/reference and that the reference
does not pin; the installed Pillow 12.2.0 (bundled libjpeg-turbo 3.x) is the version of record.  Restated here from
the published libjpeg algorithm, with the decoder defaults Pillow leaves in place:
  * sequential Huffman baseline / extended-sequential 8-bit DCT (SOF0 / SOF1), restart intervals (jdhuff.c),
  * dct_method = JDCT_ISLOW: the 13-bit fixed-point Loeffler-Ligtenberg-Moschytz inverse DCT (jidctint.c),
  * do_fancy_upsampling = TRUE: triangle-filter chroma upsampling for 4:2:2 (h2v1) and 4:2:0 (h2v2) (jdsample.c),
    with the edge context of jdmainct.c (first / last real row duplicated),
  * YCbCr -> RGB with libjpeg's 16-bit fixed-point tables (jdcolor.c); grayscale is replicated to RGB (Pillow convert).
Pinned bit-for-bit against the installed Pillow by tests/test_jpeg_oracle.py.  Pure python: small images only.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
                   7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                   39, 46, 53, 60, 61, 54, 47, 55, 62, 63])

CONST_BITS, PASS1_BITS = 13, 2
FIX_0_298631336, FIX_0_390180644, FIX_0_541196100, FIX_0_765366865 = 2446, 3196, 4433, 6270
FIX_0_899976223, FIX_1_175875602, FIX_1_501321110, FIX_1_847759065 = 7373, 9633, 12299, 15137
FIX_1_961570560, FIX_2_053119869, FIX_2_562915447, FIX_3_072711026 = 16069, 16819, 20995, 25172


class JpegError(ValueError):
    pass


def parse(data):
    """Marker segments -> dict(frame, qt, huff, scan data offset ...).  Raises JpegError for anything but 8-bit
    sequential Huffman JPEG with 1 or 3 components."""
    if data[:2] != b"\xff\xd8":
        raise JpegError("not a JPEG (no SOI)")
    pos = 2
    qt, huff = {}, {}
    frame, restart, adobe_transform = None, 0, None
    while True:
        if data[pos] != 0xFF:
            raise JpegError("marker expected at %d" % pos)
        while data[pos + 1] == 0xFF:
            pos += 1
        m = data[pos + 1]
        pos += 2
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        ln = (data[pos] << 8) | data[pos + 1]
        seg = data[pos + 2:pos + ln]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    t = [(seg[i + 2 * k] << 8) | seg[i + 2 * k + 1] for k in range(64)]
                    i += 128
                else:
                    t = list(seg[i:i + 64])
                    i += 64
                nat = np.zeros(64, dtype=np.int64)
                nat[ZIGZAG] = t
                qt[tq] = nat
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                huff[(tc, th)] = (counts, list(seg[i + 17:i + 17 + n]))
                i += 17 + n
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegError("only 8-bit samples")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = [dict(id=seg[6 + 3 * k], h=seg[7 + 3 * k] >> 4, v=seg[7 + 3 * k] & 15, tq=seg[8 + 3 * k])
                     for k in range(nc)]
            frame = dict(h=h, w=w, comps=comps)
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise JpegError("unsupported JPEG process (SOF%d): only sequential Huffman baseline is decoded" % (m - 0xC0))
        elif m == 0xDD:
            restart = (seg[0] << 8) | seg[1]
        elif m == 0xEE and seg[:5] == b"Adobe":
            adobe_transform = seg[11]
        elif m == 0xDA:
            ns = seg[0]
            sel = {seg[1 + 2 * k]: (seg[2 + 2 * k] >> 4, seg[2 + 2 * k] & 15) for k in range(ns)}
            if frame is None or ns != len(frame["comps"]):
                raise JpegError("non-interleaved or multi-scan files are not supported")
            for c in frame["comps"]:
                c["td"], c["ta"] = sel[c["id"]]
            return dict(frame=frame, qt=qt, huff=huff, restart=restart, scan=pos + ln, adobe=adobe_transform)
        elif m == 0xD9:
            raise JpegError("EOI before SOS")
        pos += ln


class _Bits:
    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        while self.n <= 24:
            b = self.d[self.p] if self.p < len(self.d) else 0
            if b == 0xFF:
                nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                if nxt == 0:
                    self.p += 2
                else:            # a marker: feed zeros, do not advance (jdhuff.c "insufficient data" path)
                    b = 0
            else:
                self.p += 1
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def get(self, k):
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self):
        """Byte-align, skip the RSTn marker."""
        self.n = 0
        self.acc = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _build(counts, symbols):
    """code -> symbol lookup by (length, code)."""
    table, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(counts[ln - 1]):
            table[(ln, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(bits, table):
    code = 0
    for ln in range(1, 17):
        code = (code << 1) | bits.get(1)
        s = table.get((ln, code))
        if s is not None:
            return s
    raise JpegError("bad Huffman code")


def _extend(v, t):
    return v if v >= (1 << (t - 1)) else v - (1 << t) + 1


def decode_coefficients(data, info):
    """-> per component int array [blocks_h][blocks_w][64] of quantised coefficients in natural order."""
    fr = info["frame"]
    comps = fr["comps"]
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mcux, mcuy = -(-fr["w"] // (8 * hmax)), -(-fr["h"] // (8 * vmax))
    if len(comps) == 1:                       # a single-component scan is never interleaved: MCU = one 8x8 block
        comps[0]["h"] = comps[0]["v"] = 1
        hmax = vmax = 1
        mcux, mcuy = -(-fr["w"] // 8), -(-fr["h"] // 8)
    tables = {k: _build(*v) for k, v in info["huff"].items()}
    coef = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), dtype=np.int64) for c in comps]
    bits = _Bits(data, info["scan"])
    pred = [0] * len(comps)
    ri, left = info["restart"], info["restart"]
    for my in range(mcuy):
        for mx in range(mcux):
            if ri and left == 0:
                bits.restart()
                pred = [0] * len(comps)
                left = ri
            for ci, c in enumerate(comps):
                dc_t, ac_t = tables[(0, c["td"])], tables[(1, c["ta"])]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = coef[ci][my * c["v"] + by, mx * c["h"] + bx]
                        t = _decode_symbol(bits, dc_t)
                        pred[ci] += _extend(bits.get(t), t) if t else 0
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(bits, ac_t)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                            k += 1
            if ri:
                left -= 1
    return coef, (hmax, vmax)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def idct_islow(coef, quant):
    """jidctint.c jpeg_idct_islow on [..., 64] coefficient blocks -> [..., 8, 8] samples (u8)."""
    ws = (coef * quant).reshape(coef.shape[:-1] + (8, 8)).astype(np.int64)   # [row][col]

    def pass_1d(v, shift_even, descale_bits):
        # v[..., k] = the 8 inputs of one 1-D transform
        z2, z3 = v[..., 2], v[..., 6]
        z1 = (z2 + z3) * FIX_0_541196100
        tmp2 = z1 + z3 * (-FIX_1_847759065)
        tmp3 = z1 + z2 * FIX_0_765366865
        z2, z3 = v[..., 0], v[..., 4]
        tmp0 = (z2 + z3) << CONST_BITS
        tmp1 = (z2 - z3) << CONST_BITS
        tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
        tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
        z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
        z5 = (z3 + z4) * FIX_1_175875602
        tmp0 = tmp0 * FIX_0_298631336
        tmp1 = tmp1 * FIX_2_053119869
        tmp2 = tmp2 * FIX_3_072711026
        tmp3 = tmp3 * FIX_1_501321110
        z1 = z1 * (-FIX_0_899976223)
        z2 = z2 * (-FIX_2_562915447)
        z3 = z3 * (-FIX_1_961570560) + z5
        z4 = z4 * (-FIX_0_390180644) + z5
        tmp0 = tmp0 + z1 + z3
        tmp1 = tmp1 + z2 + z4
        tmp2 = tmp2 + z2 + z3
        tmp3 = tmp3 + z1 + z4
        out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0,
                        tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)
        return _descale(out, descale_bits)
    # pass 1: columns (the 8 inputs of a column are ws[0..7][col])
    cols = pass_1d(np.swapaxes(ws, -1, -2), None, CONST_BITS - PASS1_BITS)     # [..., col, k]
    ws2 = np.swapaxes(cols, -1, -2)                                            # [..., row, col]
    rows = pass_1d(ws2, None, CONST_BITS + PASS1_BITS + 3)
    # range_limit[(x & RANGE_MASK)] of jdmaster.c prepare_range_limit_table: clamp(x + 128) on the masked 10-bit value
    x = rows & 1023
    x = np.where(x >= 512, x - 1024, x)
    return np.clip(x + 128, 0, 255).astype(np.uint8)


def _plane(samples):
    """[bh][bw][8][8] -> [bh*8][bw*8]"""
    bh, bw = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def _h2v1_fancy_row(r):
    """jdsample.c h2v1_fancy_upsample on one row (int array, downsampled_width columns)."""
    n = len(r)
    out = np.zeros(2 * n, dtype=np.int64)
    if n == 1:
        out[0] = out[1] = r[0]
        return out
    r = r.astype(np.int64)
    out[0] = r[0]
    out[1] = (r[0] * 3 + r[1] + 2) >> 2
    mid = r[1:-1]
    out[2:2 * n - 2:2] = (mid * 3 + r[:-2] + 1) >> 2
    out[3:2 * n - 1:2] = (mid * 3 + r[2:] + 2) >> 2
    out[2 * n - 2] = (r[-1] * 3 + r[-2] + 1) >> 2
    out[2 * n - 1] = r[-1]
    return out


def _h2v2_fancy(p):
    """jdsample.c h2v2_fancy_upsample on a [rows][cols] plane (real rows / columns only); the context row above the
    first row is the first row, below the last the last (jdmainct.c)."""
    p = p.astype(np.int64)
    rows, n = p.shape
    out = np.zeros((2 * rows, 2 * n), dtype=np.int64)
    for y in range(rows):
        for v in range(2):
            near = p[y]
            far = p[max(y - 1, 0)] if v == 0 else p[min(y + 1, rows - 1)]
            s = near * 3 + far                        # column sums
            o = out[2 * y + v]
            if n == 1:
                o[0] = (s[0] * 4 + 8) >> 4
                o[1] = (s[0] * 4 + 7) >> 4
                continue
            o[0] = (s[0] * 4 + 8) >> 4
            o[1] = (s[0] * 3 + s[1] + 7) >> 4
            o[2:2 * n - 2:2] = (s[1:-1] * 3 + s[:-2] + 8) >> 4
            o[3:2 * n - 1:2] = (s[1:-1] * 3 + s[2:] + 7) >> 4
            o[2 * n - 2] = (s[-1] * 3 + s[-2] + 8) >> 4
            o[2 * n - 1] = (s[-1] * 4 + 7) >> 4
    return out


def _fix(x):
    return int(x * 65536 + 0.5)


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c build_ycc_rgb_table + ycc_rgb_convert."""
    i = np.arange(256, dtype=np.int64) - 128
    cr_r = (_fix(1.40200) * i + 32768) >> 16
    cb_b = (_fix(1.77200) * i + 32768) >> 16
    cr_g = -_fix(0.71414) * i
    cb_g = -_fix(0.34414) * i + 32768
    y = y.astype(np.int64)
    r = np.clip(y + cr_r[cr], 0, 255)
    g = np.clip(y + ((cb_g[cb] + cr_g[cr]) >> 16), 0, 255)
    b = np.clip(y + cb_b[cb], 0, 255)
    return np.stack([r, g, b], axis=-1).astype(np.uint8)


def decode(data):
    """bytes of a baseline JPEG file -> HxWx3 u8, equal to np.asarray(PIL.Image.open(...).convert('RGB'))."""
    data = bytes(data)
    info = parse(data)
    fr = info["frame"]
    H, W = fr["h"], fr["w"]
    coef, (hmax, vmax) = decode_coefficients(data, info)
    planes = [_plane(idct_islow(c, info["qt"][comp["tq"]])) for c, comp in zip(coef, fr["comps"])]
    if len(planes) == 1:
        g = planes[0][:H, :W]
        return np.stack([g, g, g], axis=-1)
    if len(planes) != 3 or info["adobe"] == 0:
        raise JpegError("only grayscale and YCbCr JPEG files are supported")
    full = [None] * 3
    for k, comp in enumerate(fr["comps"]):
        hs, vs = hmax // comp["h"], vmax // comp["v"]
        dw, dh = -(-W * comp["h"] // hmax), -(-H * comp["v"] // vmax)     # downsampled_width / height (jdmaster.c)
        p = planes[k][:dh, :dw]
        if (hs, vs) == (1, 1):
            full[k] = p
        elif (hs, vs) == (2, 1):
            full[k] = np.stack([_h2v1_fancy_row(r) for r in p])
        elif (hs, vs) == (2, 2):
            full[k] = _h2v2_fancy(p)
        else:
            raise JpegError("unsupported chroma subsampling %dx%d" % (hs, vs))
        full[k] = full[k][:H, :W]
    return ycc_to_rgb(full[0], full[1].astype(np.int64), full[2].astype(np.int64))
