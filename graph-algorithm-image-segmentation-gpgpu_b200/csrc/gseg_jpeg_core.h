// gseg_jpeg_core.h -- the arithmetic of the in-house baseline-JPEG decoder, written once
// for both sides: under nvcc every function is __host__ __device__ and the kernels of gseg_jpeg.cuh call it with one
// thread per restart interval / 8x8 block / pixel group; under a plain C++ compiler the same functions are driven by
// loops in tests/jpeg_host.cpp (test infrastructure) so that the code the GPU runs can be checked on a machine
// without a GPU against libjpeg (cv2.imdecode).  Nothing in the product calls the host build.
//
// What is decoded: ITU-T T.81 baseline / extended-sequential Huffman, 8-bit samples, one interleaved scan (or a
// single-component image), 1 or 3 components (Y / YCbCr, JFIF), luma sampling 1x1, 2x1, 1x2, 2x2 or 4x1 over 1x1
// chroma.  The reference's batch benchmark reads exactly such files through cv::imread (README.md:26); the
// arithmetic below is the one libjpeg's default decoder applies (accurate integer IDCT "islow", "fancy" triangle
// upsampling of subsampled chroma, 16-bit fixed-point YCbCr -> RGB), so the decoded pixels are the ones cv::imread
// hands to the reference's segmentation.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GSEG_HD __host__ __device__ __forceinline__
#else
#define GSEG_HD inline
#endif

#define JPG_LOOK 9            // bits of the first-level Huffman lookup
#define JPG_MAXCOMP 3
#define JPG_ERR_CODE 1u       // a bit pattern that is no Huffman code
#define JPG_ERR_COEF 2u       // a run that leaves the block

struct JpegHuff {
    uint16_t look[1 << JPG_LOOK]; // (code length << 8 | symbol) for codes of <= JPG_LOOK bits, else 0
    int32_t maxcode[18];          // largest code of each length (-1: none), [17] = sentinel
    int32_t valoff[17];           // index of a length's first symbol minus its first code
    uint8_t vals[256];
};

// Everything the kernels need to know about one image; built by jpeg_parse (gseg_jpeg.hpp) on the host and copied to
// the device in front of the compressed bytes.
struct JpegDev {
    int32_t w, h, ncomp;
    int32_t maxh, maxv;           // largest sampling factors
    int32_t mcus_x, mcus_y, nmcu; // MCU grid of the scan
    int32_t ri, nint;             // MCUs per restart interval (the whole scan when the file has none), intervals
    int32_t hs[JPG_MAXCOMP], vs[JPG_MAXCOMP];   // blocks of a component inside one MCU
    int32_t bw[JPG_MAXCOMP], bh[JPG_MAXCOMP];   // component plane in blocks (MCU padded)
    int32_t dw[JPG_MAXCOMP], dh[JPG_MAXCOMP];   // component plane in real samples (downsampled size)
    int32_t blk_off[JPG_MAXCOMP];               // first block of the component in the coefficient array
    int32_t pix_off[JPG_MAXCOMP];               // first sample of the component in the sample array
    int32_t nblocks, nsamples;
    uint32_t data_off, data_end;  // entropy-coded bytes [data_off, data_end) of the staged file
    uint32_t error;               // JPG_ERR_* bits, set by the decoding threads
    uint16_t quant[JPG_MAXCOMP][64]; // per component, natural (row-major) order
    JpegHuff dc[JPG_MAXCOMP], ac[JPG_MAXCOMP]; // per component
};

// zigzag position -> row-major position (padded: a corrupt run may index past 63 before it is rejected)
#define JPG_ZIGZAG_INIT                                                                                                        \
    {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,             \
     6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,             \
     39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63}
static const uint8_t jpg_zigzag_h[80] = JPG_ZIGZAG_INIT;
#if defined(__CUDACC__)
static __device__ __constant__ uint8_t jpg_zigzag_d[80] = JPG_ZIGZAG_INIT;
#endif
GSEG_HD int jpg_zigzag(int k) {
#if defined(__CUDA_ARCH__)
    return jpg_zigzag_d[k];
#else
    return jpg_zigzag_h[k];
#endif
}

// ---- bit reader over the entropy-coded segment (T.81 F.2.2.5): 0xFF00 is a stuffed 0xFF, any other marker ends the
// data of the interval and the reader feeds zero bits from there on.
struct JpegBits {
    const uint8_t *data;
    uint32_t pos, end;
    uint32_t buf; // left-aligned
    int n;
};
GSEG_HD void jpg_bits_init(JpegBits &b, const uint8_t *data, uint32_t pos, uint32_t end) {
    b.data = data; b.pos = pos; b.end = end; b.buf = 0u; b.n = 0;
}
GSEG_HD void jpg_fill(JpegBits &b) { // at least 25 valid bits afterwards
    while (b.n <= 24) {
        uint32_t v = 0u;
        if (b.pos < b.end) {
            v = b.data[b.pos];
            if (v == 0xFFu) {
                const uint32_t m = b.pos + 1 < b.end ? b.data[b.pos + 1] : 0xD9u;
                if (m == 0u) b.pos += 2;        // stuffed byte
                else { v = 0u; b.end = b.pos; } // marker: stop here
            } else ++b.pos;
        }
        b.buf |= v << (24 - b.n);
        b.n += 8;
    }
}
GSEG_HD uint32_t jpg_peek(const JpegBits &b, int k) { return b.buf >> (32 - k); } // 1 <= k <= 16
GSEG_HD void jpg_skip(JpegBits &b, int k) { b.buf <<= k; b.n -= k; }

GSEG_HD int jpg_symbol(JpegBits &b, const JpegHuff &t, uint32_t &err) {
    jpg_fill(b);
    const uint32_t e = t.look[jpg_peek(b, JPG_LOOK)];
    if (e) { jpg_skip(b, (int)(e >> 8)); return (int)(e & 255u); }
    for (int l = JPG_LOOK + 1; l <= 16; ++l) {
        const int32_t code = (int32_t)jpg_peek(b, l);
        if (code <= t.maxcode[l]) { jpg_skip(b, l); return t.vals[(code + t.valoff[l]) & 255]; }
    }
    err |= JPG_ERR_CODE;
    jpg_skip(b, 16);
    return 0;
}
// s more bits as a signed value (T.81 F.2.2.1 EXTEND)
GSEG_HD int jpg_receive_extend(JpegBits &b, int s) {
    jpg_fill(b);
    const int v = (int)jpg_peek(b, s);
    jpg_skip(b, s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// One 8x8 block: DC difference + AC run/size pairs (T.81 F.2.2); non-zero coefficients go to coef[] (row-major,
// still quantised; the array was cleared before).  Returns the new DC predictor.
GSEG_HD int jpg_decode_block(JpegBits &b, const JpegHuff &dc, const JpegHuff &ac, int pred, int16_t *coef, uint32_t &err) {
    const int s = jpg_symbol(b, dc, err) & 15;
    if (s) pred += jpg_receive_extend(b, s);
    if (pred) coef[0] = (int16_t)pred;
    for (int k = 1; k < 64; ++k) {
        const int rs = jpg_symbol(b, ac, err);
        const int r = rs >> 4, sz = rs & 15;
        if (sz) {
            k += r;
            const int v = jpg_receive_extend(b, sz);
            if (k > 63) { err |= JPG_ERR_COEF; break; }
            coef[jpg_zigzag(k)] = (int16_t)v;
        } else {
            if (r != 15) break; // EOB
            k += 15;
        }
    }
    return pred;
}

// One restart interval: MCUs [first, last) of the scan, predictors start at zero.
GSEG_HD void jpg_decode_interval(const JpegDev &d, const JpegHuff *dc, const JpegHuff *ac, const uint8_t *file, uint32_t start,
                                 int first, int last, int16_t *coef, uint32_t &err) {
    JpegBits b;
    jpg_bits_init(b, file, start, d.data_end);
    int pred[JPG_MAXCOMP] = {0, 0, 0};
    int mx = first % d.mcus_x, my = first / d.mcus_x;
    for (int m = first; m < last; ++m) {
        for (int c = 0; c < d.ncomp; ++c)
            for (int v = 0; v < d.vs[c]; ++v)
                for (int hh = 0; hh < d.hs[c]; ++hh) {
                    const int blk = d.blk_off[c] + (my * d.vs[c] + v) * d.bw[c] + mx * d.hs[c] + hh;
                    pred[c] = jpg_decode_block(b, dc[c], ac[c], pred[c], coef + (size_t)blk * 64, err);
                }
        if (++mx == d.mcus_x) { mx = 0; ++my; }
    }
}

// ---- inverse DCT: libjpeg's accurate integer method (jidctint.c, "islow": Loeffler-Ligtenberg-Moschytz, 13-bit
// constants, two passes with 2 extra bits kept between them), restated.  in: 64 quantised coefficients, q: the
// quantisation table, out: 8 rows of 8 samples, `pitch` apart.
#define JPG_DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
GSEG_HD void jpg_idct_1d(int i0, int i1, int i2, int i3, int i4, int i5, int i6, int i7, int shift, int *o) {
    int z2 = i2, z3 = i6;
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137);
    int tmp3 = z1 + z2 * 6270;
    int tmp0 = (i0 + i4) * 8192;
    int tmp1 = (i0 - i4) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = i7; tmp1 = i5; tmp2 = i3; tmp3 = i1;
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    o[0] = JPG_DESCALE(tmp10 + tmp3, shift); o[7] = JPG_DESCALE(tmp10 - tmp3, shift);
    o[1] = JPG_DESCALE(tmp11 + tmp2, shift); o[6] = JPG_DESCALE(tmp11 - tmp2, shift);
    o[2] = JPG_DESCALE(tmp12 + tmp1, shift); o[5] = JPG_DESCALE(tmp12 - tmp1, shift);
    o[3] = JPG_DESCALE(tmp13 + tmp0, shift); o[4] = JPG_DESCALE(tmp13 - tmp0, shift);
}
GSEG_HD uint8_t jpg_range_limit(int v) { // libjpeg's table look-up with a 10-bit index: clamp, wrapping far out of range
    v &= 1023;
    if (v >= 512) v -= 1024;
    v += 128;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
GSEG_HD void jpg_idct_block(const int16_t *in, const uint16_t *q, uint8_t *out, int pitch) {
    int ws[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int o[8];
        jpg_idct_1d(in[c] * q[c], in[8 + c] * q[8 + c], in[16 + c] * q[16 + c], in[24 + c] * q[24 + c], in[32 + c] * q[32 + c],
                    in[40 + c] * q[40 + c], in[48 + c] * q[48 + c], in[56 + c] * q[56 + c], 11, o);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[8 * r + c] = o[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int o[8];
        jpg_idct_1d(ws[8 * r], ws[8 * r + 1], ws[8 * r + 2], ws[8 * r + 3], ws[8 * r + 4], ws[8 * r + 5], ws[8 * r + 6], ws[8 * r + 7], 18, o);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[(size_t)r * pitch + c] = jpg_range_limit(o[c]);
    }
}

// ---- chroma upsampling (jdsample.c) + colour conversion (jdcolor.c), per output pixel.
// Sample (x, y) of the full-size image from component plane `p` (pitch `pw` samples, real size dw x dh) that is
// subsampled by hx horizontally and vy vertically relative to the image.
GSEG_HD int jpg_upsample(const uint8_t *p, int pw, int dw, int dh, int hx, int vy, int x, int y) {
    if (hx == 1 && vy == 1) return p[(size_t)y * pw + x];
    if (hx == 2 && vy == 1 && dw > 2) { // h2v1 "fancy": 3/4 nearer + 1/4 further sample
        const uint8_t *r = p + (size_t)y * pw;
        const int i = x >> 1, t = r[i];
        if (x & 1) return i == dw - 1 ? t : (3 * t + r[i + 1] + 2) >> 2;
        return i == 0 ? t : (3 * t + r[i - 1] + 1) >> 2;
    }
    if (hx == 2 && vy == 2 && dw > 2) { // h2v2 "fancy": the same in both directions, 16ths
        const int j = y >> 1, i = x >> 1;
        int jn = (y & 1) ? j + 1 : j - 1;
        jn = jn < 0 ? 0 : (jn > dh - 1 ? dh - 1 : jn);
        const uint8_t *r0 = p + (size_t)j * pw, *r1 = p + (size_t)jn * pw;
        const int t = 3 * r0[i] + r1[i];
        if (x & 1) return i == dw - 1 ? (4 * t + 7) >> 4 : (3 * t + 3 * r0[i + 1] + r1[i + 1] + 7) >> 4;
        return i == 0 ? (4 * t + 8) >> 4 : (3 * t + 3 * r0[i - 1] + r1[i - 1] + 8) >> 4;
    }
    if (hx == 1 && vy == 2) { // h1v2 "fancy"
        const int j = y >> 1;
        int jn = (y & 1) ? j + 1 : j - 1;
        jn = jn < 0 ? 0 : (jn > dh - 1 ? dh - 1 : jn);
        const int t = 3 * p[(size_t)j * pw + x] + p[(size_t)jn * pw + x];
        return (t + ((y & 1) ? 2 : 1)) >> 2;
    }
    return p[(size_t)(y / vy) * pw + x / hx]; // replication
}
GSEG_HD uint8_t jpg_clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
GSEG_HD void jpg_ycc_rgb(int y, int cb, int cr, uint8_t *rgb) { // 16-bit fixed point, libjpeg's tables written out
    cb -= 128; cr -= 128;
    rgb[0] = jpg_clamp8(y + ((91881 * cr + 32768) >> 16));
    rgb[1] = jpg_clamp8(y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
    rgb[2] = jpg_clamp8(y + ((116130 * cb + 32768) >> 16));
}
// RGB of pixel (x, y) from the decoded component planes.
GSEG_HD void jpg_pixel(const JpegDev &d, const uint8_t *samples, int x, int y, uint8_t *rgb) {
    const int Y = jpg_upsample(samples + d.pix_off[0], d.bw[0] * 8, d.dw[0], d.dh[0], d.maxh / d.hs[0], d.maxv / d.vs[0], x, y);
    if (d.ncomp == 1) { rgb[0] = rgb[1] = rgb[2] = (uint8_t)Y; return; }
    const int Cb = jpg_upsample(samples + d.pix_off[1], d.bw[1] * 8, d.dw[1], d.dh[1], d.maxh / d.hs[1], d.maxv / d.vs[1], x, y);
    const int Cr = jpg_upsample(samples + d.pix_off[2], d.bw[2] * 8, d.dw[2], d.dh[2], d.maxh / d.hs[2], d.maxv / d.vs[2], x, y);
    jpg_ycc_rgb(Y, Cb, Cr, rgb);
}
