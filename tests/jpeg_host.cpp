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
    for (int i = 0; i < d.nint; ++i) { // k_jpeg_huff: one thread per restart interval
        const int first = i * d.ri, last = first + d.ri < d.nmcu ? first + d.ri : d.nmcu;
        jpg_decode_interval(d, d.dc, d.ac, jpg_zigzag_h, file, plan.starts[i], first, last, coef, err);
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

int jpeg_host_info(const uint8_t *file, size_t n, int *w, int *h) { return jpeg_peek_size(file, n, w, h); }

const char *jpeg_host_why(const uint8_t *file, size_t n) {
    static thread_local JpegPlan plan;
    jpeg_parse(file, n, plan);
    return plan.why;
}
}
