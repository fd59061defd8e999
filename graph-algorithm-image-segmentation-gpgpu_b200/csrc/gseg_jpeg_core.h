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

#define JPG_LOOK 10           // bits of the first-level Huffman lookup
#define JPG_MAXCOMP 3
#define JPG_ERR_CODE 1u       // a bit pattern that is no Huffman code
#define JPG_ERR_COEF 2u       // a run that leaves the block
#define JPG_ERR_RST 4u        // fewer restart markers in the file than its restart interval promises
#define JPG_ERR_BLOCKS 8u     // the entropy-coded data does not hold the image's number of blocks

struct alignas(16) JpegHuff {
    uint16_t look[1 << JPG_LOOK]; // (code length << 8 | symbol) for codes of <= JPG_LOOK bits, else 0
    uint32_t thr[8];              // longer codes: a 16-bit window c has a code of <= JPG_LOOK+1+i bits iff c < thr[i]
    int32_t valoff[20];           // index of a length's first symbol minus its first code
    uint8_t vals[256];
};

// Everything the kernels need to know about one image; built by jpeg_parse (gseg_jpeg.hpp) on the host and copied to
// the device in front of the compressed bytes.
struct alignas(16) JpegDev {
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
    uint32_t data_off, data_end;  // entropy-coded bytes [data_off, data_end) of the staged file (data_end: an upper bound)
    uint32_t pad0[2];
    uint16_t quant[JPG_MAXCOMP][64]; // per component, natural (row-major) order
    JpegHuff dc[JPG_MAXCOMP], ac[JPG_MAXCOMP]; // per component
};

// zigzag position -> row-major position (padded: a corrupt run may index past 63 before it is rejected).  The kernels
// copy it to shared memory: every thread of a warp asks for a different entry, which a __constant__ array serialises.
#define JPG_ZIGZAG_LEN 80
static const uint8_t jpg_zigzag_h[JPG_ZIGZAG_LEN] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

// ---- bit reader over the entropy-coded segment (T.81 F.2.2.5): 0xFF00 is a stuffed 0xFF, any other marker ends the
// data of the interval and the reader feeds zero bits from there on.
// A thread walks its interval serially, so the reader is built around latency and instruction count: the file is read
// in aligned 16-byte chunks, a chunk ahead of its use (one global load per ~20 symbols, never waited for), into a
// 32-byte ring per thread (shared memory on the device); four bytes at any offset are two ring words and a funnel
// shift, and enter the 64-bit bit buffer together unless one of them is 0xFF; one refill per symbol covers the code
// and its extra bits.  The staged file must be readable up to 64 bytes past data_end.
struct JpegChunk { uint32_t w[4]; };
GSEG_HD JpegChunk jpg_load16(const uint8_t *file, uint32_t chunk) {
    JpegChunk c;
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(file) + chunk); // the staged file is 16-byte aligned
    c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
#else
    for (int i = 0; i < 4; ++i) {
        const uint8_t *p = file + 16 * (size_t)chunk + 4 * i;
        c.w[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
#endif
    return c;
}
GSEG_HD uint32_t jpg_bswap(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0u, 0x0123);
#else
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
#endif
}
struct JpegBits {
    const uint8_t *file;
    uint32_t *ring;      // eight words, word k at ring[k * rs]: the two 16-byte chunks the stream is in (chunk c in half c & 1)
    uint32_t rs;         // ... in shared memory with rs = threads per block (every thread's word k in its own bank), 1 on the host
    JpegChunk pend;      // the chunk behind them, loaded ahead of its use
    uint32_t pos;        // file offset of the next byte to enter the bit buffer
    uint32_t end0;       // end of the stream (file offset)
    uint32_t ffhist;     // bit i: the (i+1)-th last byte that entered the bit buffer was a stuffed 0xFF (two file bytes)
    uint64_t buf;        // bit buffer, left-aligned
    int n;               // ... valid bits
    bool eof;            // a marker or the end of the data was reached: zero bits from here on
    bool rst;            // ... and it was a restart marker, at file offset `mark`
    uint32_t mark;
    int nfake;           // ... how many of the buffer's bits are such zeros (they sit behind the real ones)
};
GSEG_HD void jpg_ring_put(JpegBits &b, uint32_t half, const JpegChunk &c) {
#pragma unroll
    for (int i = 0; i < 4; ++i) b.ring[(4u * half + (uint32_t)i) * b.rs] = c.w[i];
}
GSEG_HD void jpg_bits_init(JpegBits &b, const uint8_t *file, uint32_t *ring, uint32_t rs, uint32_t pos, uint32_t end) {
    b.file = file; b.ring = ring; b.rs = rs;
    const uint32_t c = pos >> 4;
    jpg_ring_put(b, c & 1u, jpg_load16(file, c));
    jpg_ring_put(b, (c + 1u) & 1u, jpg_load16(file, c + 1u));
    b.pend = jpg_load16(file, c + 2u);
    b.pos = pos;
    b.end0 = end > pos ? end : pos; b.ffhist = 0u;
    b.buf = 0; b.n = 0; b.eof = false; b.rst = false; b.mark = 0u; b.nfake = 0;
}
GSEG_HD void jpg_advance(JpegBits &b, uint32_t nbytes) { // nbytes <= 4: at most one chunk border
    const uint32_t oldc = b.pos >> 4;
    b.pos += nbytes;
    if ((b.pos >> 4) != oldc) { // the chunk left behind makes room for the one loaded ahead; load the next
        jpg_ring_put(b, oldc & 1u, b.pend);
        b.pend = jpg_load16(b.file, oldc + 3u);
    }
}
GSEG_HD void jpg_fill(JpegBits &b) { // at least 33 valid bits afterwards: a code (<= 16) and its extra bits (<= 16)
    while (b.n <= 32) {
        const uint32_t wi = (b.pos >> 2) & 7u, sh = (b.pos & 3u) * 8u;
        const uint32_t w0 = b.ring[wi * b.rs], w1 = b.ring[((wi + 1u) & 7u) * b.rs];
#if defined(__CUDA_ARCH__)
        const uint32_t W = jpg_bswap(__funnelshift_r(w0, w1, sh)); // the stream's next four bytes, first byte on top
#else
        const uint32_t W = jpg_bswap(sh ? (w0 >> sh) | (w1 << (32u - sh)) : w0);
#endif
        const uint32_t x = ~W; // a byte of W is 0xFF <=> that byte of x is zero
        const uint32_t ff = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
        const uint32_t rem = b.end0 - b.pos;
        if (!b.eof && rem >= 4u && ff == 0u) { // four plain bytes
            b.buf |= (uint64_t)W << (32 - b.n);
            b.n += 32; b.ffhist <<= 4;
            jpg_advance(b, 4u);
        } else { // one byte at a time
            uint32_t v = 0u;
            b.ffhist <<= 1;
            if (!b.eof && rem > 0u) {
                v = W >> 24;
                if (v == 0xFFu) {
                    const uint32_t m = rem > 1u ? ((W >> 16) & 0xFFu) : 0xD9u;
                    if (m == 0u) { jpg_advance(b, 2u); b.ffhist |= 1u; } // stuffed byte
                    else { v = 0u; b.eof = true; b.rst = (m & 0xF8u) == 0xD0u; b.mark = b.pos; } // marker: stop here
                } else jpg_advance(b, 1u);
            } else b.eof = true;
            if (b.eof) b.nfake += 8;
            b.buf |= (uint64_t)v << (56 - b.n);
            b.n += 8;
        }
    }
}
GSEG_HD uint32_t jpg_peek(const JpegBits &b, int k) { return (uint32_t)(b.buf >> (64 - k)); } // 1 <= k <= 32
GSEG_HD void jpg_skip(JpegBits &b, int k) { b.buf <<= k; b.n -= k; }

// File position of the first unread bit, in bits (byte offset * 8 + bits of that byte already consumed): the unread
// real bits are the tail of the last bytes that entered the buffer, and every stuffed 0xFF among them stands for two
// file bytes.  A position never names the 0x00 of a stuffed pair.  End of the data once only padding zeros are left.
GSEG_HD uint32_t jpg_bits_pos(const JpegBits &b) {
    const int nr = b.n - b.nfake;
    if (nr <= 0) return b.eof ? b.end0 * 8u : b.pos * 8u;
    const uint32_t q = (uint32_t)(nr + 7) >> 3;
    const uint32_t m = (b.ffhist >> (b.nfake >> 3)) & ((1u << q) - 1u);
#if defined(__CUDA_ARCH__)
    const uint32_t st = (uint32_t)__popc(m);
#else
    uint32_t st = 0u;
    for (uint32_t t = m; t; t &= t - 1u) ++st;
#endif
    return (b.pos - q - st) * 8u + (8u * q - (uint32_t)nr);
}
// Reader positioned at a bit position as jpg_bits_pos reports them; the buffer is filled.
GSEG_HD void jpg_bits_init_bit(JpegBits &b, const uint8_t *file, uint32_t *ring, uint32_t rs, uint32_t bitpos, uint32_t end) {
    jpg_bits_init(b, file, ring, rs, bitpos >> 3, end);
    jpg_fill(b);
    jpg_skip(b, (int)(bitpos & 7u));
}

// One Huffman symbol; the caller has filled the bit buffer.
GSEG_HD int jpg_symbol(JpegBits &b, const JpegHuff &t, uint32_t &err) {
    const uint32_t c = jpg_peek(b, 16);
    const uint32_t e = t.look[c >> (16 - JPG_LOOK)];
    if (e) { jpg_skip(b, (int)(e >> 8)); return (int)(e & 255u); }
    int l = JPG_LOOK + 1; // the thresholds grow with the length: count the lengths the window is too large for
#pragma unroll
    for (int i = 0; i < 16 - JPG_LOOK; ++i) l += c >= t.thr[i] ? 1 : 0;
    if (l > 16) { err |= JPG_ERR_CODE; jpg_skip(b, 16); return 0; }
    jpg_skip(b, l);
    return t.vals[((int)(c >> (16 - l)) + t.valoff[l]) & 255];
}
// s more bits as a signed value (T.81 F.2.2.1 EXTEND), s = 0 included (no bits, value 0); the caller has filled the
// bit buffer.  Branch-free: every lane of a warp executes it on every symbol.
GSEG_HD int jpg_receive_extend(JpegBits &b, int s) {
    const int v = (int)((b.buf >> 1) >> (63 - s));
    b.buf <<= s; b.n -= s;
    return v < (int)((1u << s) >> 1) ? v - (1 << s) + 1 : v;
}

// One restart interval: MCUs [first, last) of the scan, predictors start at zero (T.81 F.2.2: per block a DC
// difference, then AC run/size pairs until 63 coefficients or EOB).  Non-zero coefficients go to coef[] (row-major
// inside a block, still quantised; the array was cleared before).
// Written as ONE loop over symbols with the position (MCU, component, block, coefficient) as state, not as nested
// loops over blocks: the 32 intervals a warp decodes then advance symbol by symbol in the same instruction stream,
// whatever block each of them is in -- with nested loops every block costs the warp its slowest lane's symbols.
GSEG_HD void jpg_decode_interval(const JpegDev &d, const JpegHuff *dc, const JpegHuff *ac, const uint8_t *zz, const uint8_t *file,
                                 uint32_t *ring, uint32_t rs, uint32_t start, int first, int last, int16_t *coef, uint32_t &err) {
    JpegBits b;
    jpg_bits_init(b, file, ring, rs, start, d.data_end);
    int p0 = 0, p1 = 0, p2 = 0;                         // DC predictors
    int m = first, mx = first % d.mcus_x, my = first / d.mcus_x;
    int c = 0, bi = 0, k = 0;                           // component, block of the component inside the MCU, coefficient
    int16_t *cb = coef + ((size_t)d.blk_off[0] + (size_t)(my * d.vs[0]) * d.bw[0] + mx * d.hs[0]) * 64;
    const JpegHuff *tdc = dc, *tac = ac;
    while (m < last) {
        jpg_fill(b);
        const bool isdc = k == 0;
        const int rs = jpg_symbol(b, isdc ? *tdc : *tac, err);
        const int sz = rs & 15;
        const int v = jpg_receive_extend(b, sz);
        if (isdc) {
            const int val = (c == 0 ? p0 : (c == 1 ? p1 : p2)) + v;
            if (c == 0) p0 = val; else if (c == 1) p1 = val; else p2 = val;
            if (val) cb[0] = (int16_t)val;
            k = 1;
        } else if (sz) {
            k += rs >> 4;
            if (k > 63) { err |= JPG_ERR_COEF; k = 64; }
            else { cb[zz[k]] = (int16_t)v; ++k; }
        } else {
            k = (rs >> 4) == 15 ? k + 16 : 64;          // ZRL : EOB
        }
        if (k >= 64) {                                  // next block
            k = 0;
            if (++bi == d.hs[c] * d.vs[c]) {
                bi = 0;
                if (++c == d.ncomp) {
                    c = 0; ++m;
                    if (++mx == d.mcus_x) { mx = 0; ++my; }
                }
                tdc = dc + c; tac = ac + c;
            }
            const int hsc = d.hs[c], bv = hsc == 1 ? bi : (hsc == 2 ? bi >> 1 : bi / hsc), bh = bi - bv * hsc;
            cb = coef + ((size_t)d.blk_off[c] + (size_t)(my * d.vs[c] + bv) * d.bw[c] + mx * hsc + bh) * 64;
        }
    }
}

// ---- files WITHOUT restart markers: self-synchronising sub-sequences ---------------------------------------------
// The entropy-coded segment is cut into sub-sequences of a fixed number of bytes; thread i decodes sub-sequence i.  Only
// the first one knows where its first code starts and what it means -- the decoder's state at a bit position is (position
// of the coefficient inside its block, block inside its MCU: k, j) -- so every other thread starts from a guess and the
// threads then iterate: take the exit state of the sub-sequence before (the state at the first code that starts behind
// the own sub-sequence's first byte), decode again if it differs from the entry state used last time.  Huffman codes
// re-synchronise after a few symbols, EOBs re-align k, and wrong guesses of j die out with the tables' differences, so
// most exits are right after the first pass and the iteration ends after a few rounds; in the worst case it is a serial
// decode, it is never wrong: sub-sequence 0 is right from the start and round r makes sub-sequence r right.  A last pass
// writes the coefficients (DC as differences: k_jpeg_dcscan turns them into values).  After Weissenberger & Schmidt,
// "Massively Parallel Huffman Decoding on GPUs" (ICPP 2018) -- restated for T.81's block structure.
// State word: bits 0..31 bit position, 32..39 k, 40..47 j.
#define JPG_STATE(p, k, j) ((uint64_t)(p) | ((uint64_t)(k) << 32) | ((uint64_t)(j) << 40))
// Decodes from `entry` until the first symbol that starts at or behind bit position end_bits (or the data / the image's
// blocks end); returns the state there.  WRITE: coefficients of the blocks, the first of which is block blk (counted over
// the scan: MCU by MCU, the MCU's blocks in order), go to coef[]; *nblk = blocks completed.
template <bool WRITE>
GSEG_HD uint64_t jpg_sub_decode(const JpegDev &d, const JpegHuff *dc, const JpegHuff *ac, const uint8_t *zz, const uint8_t *file,
                                uint32_t *ring, uint32_t rs, uint64_t entry, uint32_t end_bits, int16_t *coef, uint32_t blk, uint32_t *nblk,
                                uint32_t &err) {
    JpegBits b;
    jpg_bits_init_bit(b, file, ring, rs, (uint32_t)entry, d.data_end);
    int k = (int)((entry >> 32) & 63u), j = (int)((entry >> 40) & 255u);
    int bpm = 0;
    for (int c = 0; c < d.ncomp; ++c) bpm += d.hs[c] * d.vs[c];
    if (j >= bpm) j = 0;
    int c = 0, bi = j;
    while (bi >= d.hs[c] * d.vs[c]) { bi -= d.hs[c] * d.vs[c]; ++c; }
    int m = 0, mx = 0, my = 0;
    int16_t *cb = coef;
    if (WRITE) {
        m = (int)(blk / (uint32_t)bpm); mx = m % d.mcus_x; my = m / d.mcus_x;
        j = (int)(blk - (uint32_t)m * (uint32_t)bpm); // == the entry state's j when the iteration has converged
        c = 0; bi = j;
        while (bi >= d.hs[c] * d.vs[c]) { bi -= d.hs[c] * d.vs[c]; ++c; }
        const int hsc = d.hs[c], bv = hsc == 1 ? bi : (hsc == 2 ? bi >> 1 : bi / hsc), bh = bi - bv * hsc;
        cb = coef + ((size_t)d.blk_off[c] + (size_t)(my * d.vs[c] + bv) * d.bw[c] + mx * hsc + bh) * 64;
    }
    const JpegHuff *tdc = dc + c, *tac = ac + c;
    uint32_t done = 0u, e2 = 0u;
    uint32_t pos = (uint32_t)entry;
    while (pos < end_bits && (!WRITE || m < d.nmcu)) {
        jpg_fill(b);
        const bool isdc = k == 0;
        const int rs = jpg_symbol(b, isdc ? *tdc : *tac, e2);
        const int sz = rs & 15;
        const int v = jpg_receive_extend(b, sz);
        if (isdc) {
            if (WRITE && v) cb[0] = (int16_t)v; // the difference; k_jpeg_dcscan sums them up
            k = 1;
        } else if (sz) {
            k += rs >> 4;
            if (k > 63) { e2 |= JPG_ERR_COEF; k = 64; }
            else { if (WRITE) cb[zz[k]] = (int16_t)v; ++k; }
        } else {
            k = (rs >> 4) == 15 ? k + 16 : 64;
        }
        if (k >= 64) {
            k = 0; ++done;
            if (++j == bpm) j = 0;
            if (++bi == d.hs[c] * d.vs[c]) {
                bi = 0;
                if (++c == d.ncomp) {
                    c = 0; ++m;
                    if (++mx == d.mcus_x) { mx = 0; ++my; }
                }
                tdc = dc + c; tac = ac + c;
            }
            if (WRITE) {
                const int hsc = d.hs[c], bv = hsc == 1 ? bi : (hsc == 2 ? bi >> 1 : bi / hsc), bh = bi - bv * hsc;
                cb = coef + ((size_t)d.blk_off[c] + (size_t)(my * d.vs[c] + bv) * d.bw[c] + mx * hsc + bh) * 64;
            }
        }
        // End of a restart interval: the reader stands at an RSTn marker and all that is left in front of it are the 1-bits
        // that pad the last byte (no Huffman code is all ones, T.81 Annex C, so a symbol still to come shows a zero).
        // Decoding goes on behind the marker with a block's DC code of an MCU's first block -- whatever state a guessed
        // entry had brought along: every marker re-synchronises.  (The DC predictors restart too: k_jpeg_dcscan.)
        if (b.eof && b.rst) {
            const int nr = b.n - b.nfake;
            if (nr <= 0 || (nr <= 7 && (uint32_t)(b.buf >> (64 - nr)) == (1u << nr) - 1u)) {
                if (k != 0 || j != 0) e2 |= JPG_ERR_RST;
                jpg_bits_init(b, file, b.ring, b.rs, b.mark + 2u, d.data_end); // (`rs` is the run/size symbol in here)
                k = 0; j = 0; c = 0; bi = 0;
                tdc = dc; tac = ac;
                if (WRITE) cb = coef + ((size_t)d.blk_off[0] + (size_t)(my * d.vs[0]) * d.bw[0] + mx * d.hs[0]) * 64;
            }
        }
        pos = jpg_bits_pos(b);
    }
    if (WRITE) err |= e2; // a pass from a guessed state runs into impossible codes all the time: only the last pass counts
    *nblk = done;
    return JPG_STATE(pos, k, j);
}
// First guess of the entry state of the sub-sequence that starts at file byte `start`: a code starts there, and it is a
// block's DC code of the MCU's first block.  (The 0x00 of a stuffed pair is not a position.)
GSEG_HD uint64_t jpg_sub_guess(const uint8_t *file, uint32_t start, uint32_t first) {
    if (start > first && file[start - 1] == 0xFFu && file[start] == 0x00u) ++start;
    return JPG_STATE(start * 8u, 0, 0);
}

// Block number t of component c in scan order (MCU by MCU, the component's blocks inside an MCU in order) -> its index
// in the coefficient array.  The DC predictor of a component runs over its blocks in exactly this order (T.81 F.1.1.5.1).
GSEG_HD size_t jpg_comp_block(const JpegDev &d, int c, uint32_t t) {
    const uint32_t per = (uint32_t)(d.hs[c] * d.vs[c]);
    const uint32_t m = t / per, bi = t - m * per;
    const uint32_t mx = m % (uint32_t)d.mcus_x, my = m / (uint32_t)d.mcus_x;
    const uint32_t bv = bi / (uint32_t)d.hs[c], bh = bi - bv * (uint32_t)d.hs[c];
    return (size_t)d.blk_off[c] + (size_t)(my * (uint32_t)d.vs[c] + bv) * (uint32_t)d.bw[c] + mx * (uint32_t)d.hs[c] + bh;
}

// ---- inverse DCT: libjpeg's accurate integer method (jidctint.c, "islow": Loeffler-Ligtenberg-Moschytz, 13-bit
// constants, two passes with 2 extra bits kept between them), restated.  in: 64 quantised coefficients, q: the
// quantisation table, out: 8 rows of 8 samples, `pitch` apart.
// The arithmetic is done in uint32_t: the same bits as libjpeg's for every decodable image (its intermediate values fit 32 bits
// with room to spare) and a defined wrap-around instead of a signed overflow for the garbage a corrupt file decodes to.
#define JPG_DESCALE(x, n) ((int)((x) + (1u << ((n) - 1))) >> (n))
GSEG_HD void jpg_idct_1d(int i0, int i1, int i2, int i3, int i4, int i5, int i6, int i7, int shift, int *o) {
    typedef uint32_t u;
    u z2 = (u)i2, z3 = (u)i6;
    u z1 = (z2 + z3) * 4433u;
    u tmp2 = z1 - z3 * 15137u;
    u tmp3 = z1 + z2 * 6270u;
    u tmp0 = ((u)i0 + (u)i4) * 8192u;
    u tmp1 = ((u)i0 - (u)i4) * 8192u;
    const u tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = (u)i7; tmp1 = (u)i5; tmp2 = (u)i3; tmp3 = (u)i1;
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    u z4 = tmp1 + tmp3;
    const u z5 = (z3 + z4) * 9633u;
    tmp0 *= 2446u; tmp1 *= 16819u; tmp2 *= 25172u; tmp3 *= 12299u;
    z1 = 0u - z1 * 7373u; z2 = 0u - z2 * 20995u; z3 = 0u - z3 * 16069u; z4 = 0u - z4 * 3196u;
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

// ---- the same, eight pixels at a time (x0 a multiple of 8, x0 + 8 <= w): what k_jpeg_rgb does wherever it can.  With
// the neighbour's column index clamped to the plane, the inner formulas of the fancy filters ARE the edge formulas
// ((3t + t + 8) >> 4 == (4t + 8) >> 4, (3t + t + 1) >> 2 == t, ...), so there is no edge case left.
GSEG_HD uint32_t jpg_ld32(const uint8_t *p) { // 4-byte aligned
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t *>(p);
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}
// Modes jpg_row8 handles for one component: 0 none, 1 full size, 2 h2v1 fancy, 3 h2v2 fancy
GSEG_HD int jpg_mode8(const JpegDev &d, int c) {
    const int hx = d.maxh / d.hs[c], vy = d.maxv / d.vs[c];
    if (hx == 1 && vy == 1) return 1;
    if (hx == 2 && vy == 1 && d.dw[c] > 2) return 2;
    if (hx == 2 && vy == 2 && d.dw[c] > 2) return 3;
    return 0;
}
GSEG_HD bool jpg_fast8(const JpegDev &d) {
    for (int c = 0; c < d.ncomp; ++c)
        if (!jpg_mode8(d, c)) return false;
    return true;
}
GSEG_HD void jpg_row8(const JpegDev &d, const uint8_t *samples, int c, int mode, int x0, int y, int *o) {
    const uint8_t *p = samples + d.pix_off[c];
    const int pw = d.bw[c] * 8, dw = d.dw[c], dh = d.dh[c];
    if (mode == 1) {
        const uint32_t a = jpg_ld32(p + (size_t)y * pw + x0), b = jpg_ld32(p + (size_t)y * pw + x0 + 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[k] = (int)((a >> (8 * k)) & 255u); o[4 + k] = (int)((b >> (8 * k)) & 255u); }
        return;
    }
    const int c0 = x0 >> 1, cl = c0 > 0 ? c0 - 1 : 0, cr = c0 + 4 < dw ? c0 + 4 : dw - 1;
    int t[6];
    if (mode == 2) {
        const uint8_t *r = p + (size_t)y * pw;
        const uint32_t a = jpg_ld32(r + c0);
        t[0] = r[cl]; t[5] = r[cr];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[1 + k] = (int)((a >> (8 * k)) & 255u);
#pragma unroll
        for (int m = 0; m < 4; ++m) { o[2 * m] = (3 * t[1 + m] + t[m] + 1) >> 2; o[2 * m + 1] = (3 * t[1 + m] + t[2 + m] + 2) >> 2; }
        return;
    }
    const int j = y >> 1;
    int jn = (y & 1) ? j + 1 : j - 1;
    jn = jn < 0 ? 0 : (jn > dh - 1 ? dh - 1 : jn);
    const uint8_t *r0 = p + (size_t)j * pw, *r1 = p + (size_t)jn * pw;
    const uint32_t a = jpg_ld32(r0 + c0), b = jpg_ld32(r1 + c0);
    t[0] = 3 * r0[cl] + r1[cl]; t[5] = 3 * r0[cr] + r1[cr];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[1 + k] = 3 * (int)((a >> (8 * k)) & 255u) + (int)((b >> (8 * k)) & 255u);
#pragma unroll
    for (int m = 0; m < 4; ++m) { o[2 * m] = (3 * t[1 + m] + t[m] + 8) >> 4; o[2 * m + 1] = (3 * t[1 + m] + t[2 + m] + 7) >> 4; }
}
GSEG_HD void jpg_pixels8(const JpegDev &d, const uint8_t *samples, int x0, int y, uint8_t *rgb) { // 24 bytes
    int Y[8], Cb[8], Cr[8];
    jpg_row8(d, samples, 0, jpg_mode8(d, 0), x0, y, Y);
    if (d.ncomp == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) rgb[3 * k] = rgb[3 * k + 1] = rgb[3 * k + 2] = (uint8_t)Y[k];
        return;
    }
    jpg_row8(d, samples, 1, jpg_mode8(d, 1), x0, y, Cb);
    jpg_row8(d, samples, 2, jpg_mode8(d, 2), x0, y, Cr);
#pragma unroll
    for (int k = 0; k < 8; ++k) jpg_ycc_rgb(Y[k], Cb[k], Cr[k], rgb + 3 * k);
}

