// tests/jpeg_host.cpp -- TEST INFRASTRUCTURE, not product code.  Drives the __host__ __device__ functions of
// csrc/gseg_jpeg_core.h with plain loops (one iteration per GPU thread of csrc/gseg_jpeg.cuh) so that the in-house
// JPEG decoder can be checked against libjpeg (cv2.imdecode) on a machine without a GPU.  Built by tests/test_jpeg.py
// with g++ into tests/_build/; nothing under the package loads it.
#include <stdlib.h>

#include "../graph-algorithm-image-segmentation-gpgpu_b200/csrc/gseg_jpeg.hpp"

extern "C" {

// 0 ok, 1 not a JPEG, 2 unsupported, 3 output too small, 4 corrupt entropy-coded data
int jpeg_host_decode(const uint8_t *file, size_t n, uint8_t *rgb, size_t cap, int *w, int *h, int *nint) {
    JpegPlan plan;
    const int rc = jpeg_parse(file, n, plan);
    if (rc) return rc;
    const JpegDev &d = plan.dev;
    *w = d.w; *h = d.h; *nint = d.nint;
    if ((size_t)3 * d.w * d.h > cap) return 3;
    uint8_t *staged = (uint8_t *)calloc(n + 64, 1); // the reader loads 16-byte chunks, one ahead: padding like the device buffer's
    memcpy(staged, file, n);
    file = staged;
    int16_t *coef = (int16_t *)calloc((size_t)d.nblocks * 64, sizeof(int16_t));
    uint8_t *samples = (uint8_t *)malloc((size_t)d.nsamples);
    uint32_t err = 0;
    uint32_t ring[8]; // the reader's chunk ring (shared memory on the device)
    for (int i = 0; i < d.nint; ++i) { // k_jpeg_huff: one thread per restart interval
        const int first = i * d.ri, last = first + d.ri < d.nmcu ? first + d.ri : d.nmcu;
        jpg_decode_interval(d, d.dc, d.ac, jpg_zigzag_h, file, ring, 1u, plan.starts[i], first, last, coef, err);
    }
    for (int c = 0; c < d.ncomp; ++c) // k_jpeg_idct: one thread per block
        for (int by = 0; by < d.bh[c]; ++by)
            for (int bx = 0; bx < d.bw[c]; ++bx)
                jpg_idct_block(coef + ((size_t)d.blk_off[c] + (size_t)by * d.bw[c] + bx) * 64, d.quant[c],
                               samples + d.pix_off[c] + (size_t)by * 8 * d.bw[c] * 8 + bx * 8, d.bw[c] * 8);
    for (int y = 0; y < d.h; ++y) // k_jpeg_rgb: one thread per group of eight pixels
        for (int x0 = 0; x0 < d.w; x0 += 8) {
            const int n = d.w - x0 < 8 ? d.w - x0 : 8;
            uint8_t *dst = rgb + ((size_t)y * d.w + x0) * 3;
            if (n == 8 && jpg_fast8(d)) jpg_pixels8(d, samples, x0, y, dst);
            else
                for (int j = 0; j < n; ++j) jpg_pixel(d, samples, x0 + j, y, dst + 3 * j);
        }
    free(coef); free(samples); free(staged);
    return err ? 4 : 0;
}

// The same for a file without restart markers through the self-synchronising sub-sequence decode (k_jpeg_sync + k_jpeg_dcscan):
// every "thread" of a round sees the exits of the round before (what the kernel's barrier gives it at worst).
// *rounds = iterations until no entry state changed.  Files with restart markers too: every marker re-synchronises.
int jpeg_host_decode_sync(const uint8_t *file, size_t n, uint8_t *rgb, size_t cap, int *w, int *h, int sub_bytes, int *rounds) {
    JpegPlan plan;
    const int rc = jpeg_parse(file, n, plan, false);
    if (rc) return rc;
    const JpegDev &d = plan.dev;
    *w = d.w; *h = d.h;
    if ((size_t)3 * d.w * d.h > cap) return 3;
    uint8_t *staged = (uint8_t *)calloc(n + 64, 1);
    memcpy(staged, file, n);
    file = staged;
    int16_t *coef = (int16_t *)calloc((size_t)d.nblocks * 64, sizeof(int16_t));
    uint8_t *samples = (uint8_t *)malloc((size_t)d.nsamples);
    uint32_t err = 0;
    uint32_t ring[8];
    const uint32_t off = d.data_off, end = d.data_end, S = (uint32_t)sub_bytes;
    const uint32_t nsub = end > off ? (end - off + S - 1) / S : 1;
    std::vector<uint64_t> entry(nsub), exit_(nsub), prev(nsub);
    std::vector<uint32_t> nblk(nsub), blk0(nsub);
    auto endbits = [&](uint32_t i) { const uint64_t e = (uint64_t)off + (uint64_t)(i + 1) * S; return (uint32_t)((e < end ? e : end) * 8u); };
    for (uint32_t i = 0; i < nsub; ++i) { // pass 1: from the guesses
        entry[i] = i == 0 ? JPG_STATE(off * 8u, 0, 0) : jpg_sub_guess(file, off + i * S, off);
        exit_[i] = jpg_sub_decode<false>(d, d.dc, d.ac, jpg_zigzag_h, file, ring, 1u, entry[i], endbits(i), nullptr, 0u, &nblk[i], err);
    }
    int r = 0;
    for (bool changed = true; changed; ++r) { // rounds
        changed = false;
        prev = exit_;
        for (uint32_t i = 1; i < nsub; ++i)
            if (prev[i - 1] != entry[i]) {
                entry[i] = prev[i - 1];
                exit_[i] = jpg_sub_decode<false>(d, d.dc, d.ac, jpg_zigzag_h, file, ring, 1u, entry[i], endbits(i), nullptr, 0u, &nblk[i], err);
                changed = true;
            }
    }
    *rounds = r;
    uint32_t run = 0;
    for (uint32_t i = 0; i < nsub; ++i) { blk0[i] = run; run += nblk[i]; }
    if (run < (uint32_t)d.nblocks) err |= JPG_ERR_BLOCKS; // (the padding behind the last block may decode to more)
    err &= JPG_ERR_BLOCKS; // the passes above ran from guessed states
    for (uint32_t i = 0; i < nsub; ++i) { // last pass: write
        uint32_t dummy;
        jpg_sub_decode<true>(d, d.dc, d.ac, jpg_zigzag_h, file, ring, 1u, entry[i], endbits(i), coef, blk0[i], &dummy, err);
    }
    for (int c = 0; c < d.ncomp; ++c) { // k_jpeg_dcscan: differences -> values
        const uint32_t nb = (uint32_t)d.nmcu * (uint32_t)(d.hs[c] * d.vs[c]);
        int pred = 0;
        for (uint32_t t = 0; t < nb; ++t) {
            int16_t *p = coef + jpg_comp_block(d, c, t) * 64;
            if (t % ((uint32_t)d.ri * (uint32_t)(d.hs[c] * d.vs[c])) == 0) pred = 0; // the predictors restart with every interval
            pred += p[0];
            p[0] = (int16_t)pred;
        }
    }
    for (int c = 0; c < d.ncomp; ++c)
        for (int by = 0; by < d.bh[c]; ++by)
            for (int bx = 0; bx < d.bw[c]; ++bx)
                jpg_idct_block(coef + ((size_t)d.blk_off[c] + (size_t)by * d.bw[c] + bx) * 64, d.quant[c],
                               samples + d.pix_off[c] + (size_t)by * 8 * d.bw[c] * 8 + bx * 8, d.bw[c] * 8);
    for (int y = 0; y < d.h; ++y)
        for (int x0 = 0; x0 < d.w; x0 += 8) {
            const int m = d.w - x0 < 8 ? d.w - x0 : 8;
            uint8_t *dst = rgb + ((size_t)y * d.w + x0) * 3;
            if (m == 8 && jpg_fast8(d)) jpg_pixels8(d, samples, x0, y, dst);
            else
                for (int j = 0; j < m; ++j) jpg_pixel(d, samples, x0 + j, y, dst + 3 * j);
        }
    free(coef); free(samples); free(staged);
    return err ? 4 : 0;
}

int jpeg_host_info(const uint8_t *file, size_t n, int *w, int *h) { return jpeg_peek_size(file, n, w, h); }

const char *jpeg_host_why(const uint8_t *file, size_t n) {
    static thread_local JpegPlan plan;
    jpeg_parse(file, n, plan);
    return plan.why;
}
}
