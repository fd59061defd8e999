// gseg_jpeg.hpp -- host side of the in-house JPEG decoder: marker parser (ITU-T T.81 Annex B) that turns a file's
// headers into the JpegDev descriptor the kernels read, and the table of restart-interval start offsets that makes
// the entropy-coded segment decodable in parallel (one thread per interval, T.81 E.2.4 / F.2.2.4: predictors and
// bit alignment restart after every RSTn marker).  Plain C++, no CUDA: the product calls it from gseg_api.cu, the
// CPU tests from tests/jpeg_host.cpp.
#pragma once
#include <string.h>

#include <vector>

#include "gseg_jpeg_core.h"

enum JpegParse {
    JPG_OK = 0,
    JPG_NOT_JPEG = 1,     // no SOI / truncated / malformed segment: nobody can decode this
    JPG_UNSUPPORTED = 2   // a valid JPEG this decoder does not take (progressive, arithmetic, 12-bit, CMYK, several scans ...)
};

struct JpegPlan {
    JpegDev dev;
    std::vector<uint32_t> starts; // file offset of the first entropy-coded byte of every restart interval
    const char *why;              // reason of JPG_UNSUPPORTED / JPG_NOT_JPEG
};

// Canonical Huffman table (T.81 Annex C) from the 16 length counts and the symbol list of a DHT segment.
static inline bool jpeg_build_huff(const uint8_t *counts, const uint8_t *symbols, int nsym, JpegHuff &t) {
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < 8; ++i) t.thr[i] = 0xFFFFFFFFu;
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        t.valoff[l] = k - code;
        for (int i = 0; i < counts[l - 1]; ++i, ++k, ++code) {
            if (k >= nsym || code >= (1 << l)) return false;
            t.vals[k] = symbols[k];
            if (l <= JPG_LOOK) {
                const int base = code << (JPG_LOOK - l);
                for (int j = 0; j < (1 << (JPG_LOOK - l)); ++j) t.look[base + j] = (uint16_t)((l << 8) | symbols[k]);
            }
        }
        // `code` is now the first value of this length that is NOT a code: a 16-bit window below it (left-aligned)
        // starts with a code of at most l bits
        if (l > JPG_LOOK) t.thr[l - JPG_LOOK - 1] = (uint32_t)code << (16 - l);
        code <<= 1;
    }
    return true;
}

// Width and height only (any JPEG with a frame header).
static inline int jpeg_peek_size(const uint8_t *f, size_t n, int *w, int *h) {
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8) return JPG_NOT_JPEG;
    size_t p = 2;
    while (p + 4 <= n) {
        if (f[p] != 0xFF) return JPG_NOT_JPEG;
        const int m = f[p + 1];
        if (m == 0xFF) { ++p; continue; } // fill byte
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { p += 2; continue; }
        if (m == 0xD9 || m == 0xDA) return JPG_NOT_JPEG; // scan before any frame header
        const size_t len = ((size_t)f[p + 2] << 8) | f[p + 3];
        if (len < 2 || p + 2 + len > n) return JPG_NOT_JPEG;
        if (m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            if (len < 8) return JPG_NOT_JPEG;
            *h = (f[p + 5] << 8) | f[p + 6]; *w = (f[p + 7] << 8) | f[p + 8];
            return *w > 0 && *h > 0 ? JPG_OK : JPG_NOT_JPEG;
        }
        p += 2 + len;
    }
    return JPG_NOT_JPEG;
}

// scan_restarts: also walk the entropy-coded data for the restart markers (the CPU test harness; the product leaves that
// to k_jpeg_scan and only bounds the data: data_end = the file's EOI, or its end).
static inline int jpeg_parse(const uint8_t *f, size_t n, JpegPlan &plan, bool scan_restarts = true) {
    JpegDev &d = plan.dev;
    memset(&d, 0, sizeof(d));
    plan.starts.clear();
    plan.why = "";
#define JPG_FAIL(code, msg) do { plan.why = msg; return code; } while (0)
    if (n < 4 || f[0] != 0xFF || f[1] != 0xD8) JPG_FAIL(JPG_NOT_JPEG, "no SOI marker");
    if (n >= 0xFFFFFFF0ull) JPG_FAIL(JPG_UNSUPPORTED, "file larger than 4 GB");
    uint16_t qt[4][64];
    bool have_q[4] = {false, false, false, false};
    struct RawHuff { uint8_t counts[16], syms[256]; int nsym; bool have; } rh[2][4];
    memset(rh, 0, sizeof(rh));
    int comp_id[JPG_MAXCOMP] = {0, 0, 0}, comp_h[JPG_MAXCOMP] = {1, 1, 1}, comp_v[JPG_MAXCOMP] = {1, 1, 1}, comp_q[JPG_MAXCOMP] = {0, 0, 0};
    bool have_frame = false;
    int dri = 0, adobe_transform = -1;
    size_t p = 2;
    for (;;) {
        if (p + 4 > n) JPG_FAIL(JPG_NOT_JPEG, "truncated before the scan");
        if (f[p] != 0xFF) JPG_FAIL(JPG_NOT_JPEG, "marker expected");
        const int m = f[p + 1];
        if (m == 0xFF) { ++p; continue; }
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { p += 2; continue; }
        if (m == 0xD8 || m == 0xD9) JPG_FAIL(JPG_NOT_JPEG, "unexpected SOI / EOI");
        const size_t len = ((size_t)f[p + 2] << 8) | f[p + 3];
        if (len < 2 || p + 2 + len > n) JPG_FAIL(JPG_NOT_JPEG, "segment length");
        const uint8_t *s = f + p + 4;
        const size_t sl = len - 2;
        if (m == 0xDB) { // DQT
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                if (tq > 3) JPG_FAIL(JPG_NOT_JPEG, "quantisation table id");
                if (pq > 1 || i + 1 + (pq ? 128 : 64) > sl) JPG_FAIL(JPG_NOT_JPEG, "quantisation table length");
                for (int k = 0; k < 64; ++k) {
                    const int v = pq ? ((s[i + 1 + 2 * k] << 8) | s[i + 2 + 2 * k]) : s[i + 1 + k];
                    qt[tq][jpg_zigzag_h[k]] = (uint16_t)v;
                }
                have_q[tq] = true;
                i += 1 + (pq ? 128 : 64);
            }
        } else if (m == 0xC4) { // DHT
            size_t i = 0;
            while (i < sl) {
                if (i + 17 > sl) JPG_FAIL(JPG_NOT_JPEG, "Huffman table length");
                const int tc = s[i] >> 4, th = s[i] & 15;
                if (tc > 1 || th > 3) JPG_FAIL(JPG_NOT_JPEG, "Huffman table id");
                int ns = 0;
                for (int k = 0; k < 16; ++k) ns += s[i + 1 + k];
                if (ns > 256 || i + 17 + ns > sl) JPG_FAIL(JPG_NOT_JPEG, "Huffman table length");
                RawHuff &r = rh[tc][th];
                memcpy(r.counts, s + i + 1, 16);
                memcpy(r.syms, s + i + 17, ns);
                r.nsym = ns; r.have = true;
                i += 17 + ns;
            }
        } else if (m == 0xC0 || m == 0xC1) { // SOF0 / SOF1: sequential Huffman
            if (have_frame) JPG_FAIL(JPG_NOT_JPEG, "two frame headers");
            if (sl < 6) JPG_FAIL(JPG_NOT_JPEG, "frame header length");
            if (s[0] != 8) JPG_FAIL(JPG_UNSUPPORTED, "sample precision other than 8 bits");
            d.h = (s[1] << 8) | s[2]; d.w = (s[3] << 8) | s[4]; d.ncomp = s[5];
            if (d.w < 1 || d.h < 1) JPG_FAIL(JPG_UNSUPPORTED, "image size not in the frame header");
            if (d.ncomp != 1 && d.ncomp != 3) JPG_FAIL(JPG_UNSUPPORTED, "component count other than 1 or 3");
            if (sl < (size_t)6 + 3 * d.ncomp) JPG_FAIL(JPG_NOT_JPEG, "frame header length");
            for (int c = 0; c < d.ncomp; ++c) {
                comp_id[c] = s[6 + 3 * c]; comp_h[c] = s[7 + 3 * c] >> 4; comp_v[c] = s[7 + 3 * c] & 15; comp_q[c] = s[8 + 3 * c];
                if (comp_h[c] < 1 || comp_h[c] > 4 || comp_v[c] < 1 || comp_v[c] > 4 || comp_q[c] > 3) JPG_FAIL(JPG_NOT_JPEG, "frame component");
            }
            have_frame = true;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC8) { // 0xC4 (DHT) was taken above; 0xCC = DAC belongs to arithmetic coding
            JPG_FAIL(JPG_UNSUPPORTED, "progressive, lossless, hierarchical or arithmetic-coded frame");
        } else if (m == 0xDD) { // DRI
            if (sl < 2) JPG_FAIL(JPG_NOT_JPEG, "DRI length");
            dri = (s[0] << 8) | s[1];
        } else if (m == 0xEE) { // Adobe: colour transform flag
            if (sl >= 12 && !memcmp(s, "Adobe", 5)) adobe_transform = s[11];
        } else if (m == 0xDA) { // SOS
            if (!have_frame) JPG_FAIL(JPG_NOT_JPEG, "scan before the frame header");
            if (sl < 1 || sl < (size_t)1 + 2 * s[0] + 3) JPG_FAIL(JPG_NOT_JPEG, "scan header length");
            if (s[0] != d.ncomp) JPG_FAIL(JPG_UNSUPPORTED, "components spread over several scans");
            if (d.ncomp == 3 && adobe_transform >= 0 && adobe_transform != 1) JPG_FAIL(JPG_UNSUPPORTED, "Adobe marker: components are not YCbCr");
            if (d.ncomp == 3 && !(comp_id[0] == 1 && comp_id[1] == 2 && comp_id[2] == 3) && !(comp_id[0] == 0 && comp_id[1] == 1 && comp_id[2] == 2))
                JPG_FAIL(JPG_UNSUPPORTED, "component ids are not the JFIF YCbCr ones");
            int td[JPG_MAXCOMP], ta[JPG_MAXCOMP];
            for (int c = 0; c < d.ncomp; ++c) {
                if (s[1 + 2 * c] != comp_id[c]) JPG_FAIL(JPG_UNSUPPORTED, "scan component order differs from the frame's");
                td[c] = s[2 + 2 * c] >> 4; ta[c] = s[2 + 2 * c] & 15;
                if (td[c] > 3 || ta[c] > 3 || !rh[0][td[c]].have || !rh[1][ta[c]].have) JPG_FAIL(JPG_NOT_JPEG, "scan names a missing Huffman table");
                if (!have_q[comp_q[c]]) JPG_FAIL(JPG_NOT_JPEG, "frame names a missing quantisation table");
            }
            const uint8_t *t = s + 1 + 2 * d.ncomp;
            if (t[0] != 0 || t[1] != 63 || t[2] != 0) JPG_FAIL(JPG_UNSUPPORTED, "spectral selection / successive approximation");
            // geometry
            if (d.ncomp == 1) { comp_h[0] = comp_v[0] = 1; } // a single-component scan is never interleaved (T.81 A.2.2)
            d.maxh = 1; d.maxv = 1;
            for (int c = 0; c < d.ncomp; ++c) { if (comp_h[c] > d.maxh) d.maxh = comp_h[c]; if (comp_v[c] > d.maxv) d.maxv = comp_v[c]; }
            if (d.ncomp == 3) {
                if (comp_h[1] != 1 || comp_v[1] != 1 || comp_h[2] != 1 || comp_v[2] != 1) JPG_FAIL(JPG_UNSUPPORTED, "chroma sampling factors other than 1x1");
                const int hv = comp_h[0] * 16 + comp_v[0];
                if (hv != 0x11 && hv != 0x21 && hv != 0x12 && hv != 0x22 && hv != 0x41) JPG_FAIL(JPG_UNSUPPORTED, "luma sampling factors");
            }
            d.mcus_x = (d.w + 8 * d.maxh - 1) / (8 * d.maxh); d.mcus_y = (d.h + 8 * d.maxv - 1) / (8 * d.maxv);
            const long long nm = (long long)d.mcus_x * d.mcus_y;
            if (nm > 0x3FFFFFFF) JPG_FAIL(JPG_UNSUPPORTED, "image too large");
            d.nmcu = (int)nm;
            long long blk = 0, pix = 0;
            for (int c = 0; c < d.ncomp; ++c) {
                d.hs[c] = comp_h[c]; d.vs[c] = comp_v[c];
                d.bw[c] = d.mcus_x * comp_h[c]; d.bh[c] = d.mcus_y * comp_v[c];
                d.dw[c] = (d.w * comp_h[c] + d.maxh - 1) / d.maxh; d.dh[c] = (d.h * comp_v[c] + d.maxv - 1) / d.maxv;
                d.blk_off[c] = (int)blk; d.pix_off[c] = (int)pix;
                blk += (long long)d.bw[c] * d.bh[c]; pix += (long long)d.bw[c] * d.bh[c] * 64;
                if (pix > 0x7FFFFFFF) JPG_FAIL(JPG_UNSUPPORTED, "image too large");
                memcpy(d.quant[c], qt[comp_q[c]], sizeof(d.quant[c]));
                if (!jpeg_build_huff(rh[0][td[c]].counts, rh[0][td[c]].syms, rh[0][td[c]].nsym, d.dc[c]) ||
                    !jpeg_build_huff(rh[1][ta[c]].counts, rh[1][ta[c]].syms, rh[1][ta[c]].nsym, d.ac[c]))
                    JPG_FAIL(JPG_NOT_JPEG, "Huffman table is not a prefix code");
            }
            d.nblocks = (int)blk; d.nsamples = (int)pix;
            d.ri = dri > 0 ? dri : d.nmcu;
            d.nint = (d.nmcu + d.ri - 1) / d.ri;
            d.data_off = (uint32_t)(p + 2 + len);
            break;
        }
        // every other segment (APPn, COM, DNL ...) is skipped
        p += 2 + len;
    }
    if (!scan_restarts) {
        d.data_end = (uint32_t)((n >= 2 && f[n - 2] == 0xFF && f[n - 1] == 0xD9) ? n - 2 : n);
        if (d.data_end < d.data_off) d.data_end = d.data_off;
        return JPG_OK;
    }
    // Restart markers: interval i starts right behind the i-th RSTn.  memchr keeps this at memory speed (one 0xFF in
    // ~200 bytes of entropy-coded data, each followed by 0x00 unless it is a marker).
    plan.starts.reserve((size_t)d.nint);
    plan.starts.push_back(d.data_off);
    size_t q = d.data_off, end = n;
    while (q + 1 < n) {
        const uint8_t *hit = (const uint8_t *)memchr(f + q, 0xFF, n - 1 - q);
        if (!hit) break;
        q = (size_t)(hit - f);
        const int m = f[q + 1];
        if (m == 0x00 || m == 0xFF) { q += (m == 0x00) ? 2 : 1; continue; }
        if (m >= 0xD0 && m <= 0xD7) {
            if ((int)plan.starts.size() < d.nint) plan.starts.push_back((uint32_t)(q + 2));
            q += 2;
            continue;
        }
        end = q; // EOI or any other marker: the scan's data ends here
        break;
    }
    d.data_end = (uint32_t)end;
    if ((int)plan.starts.size() != d.nint) JPG_FAIL(JPG_NOT_JPEG, "restart markers do not match the restart interval");
    return JPG_OK;
#undef JPG_FAIL
}
