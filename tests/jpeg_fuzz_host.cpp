// tests/jpeg_fuzz_host.cpp -- TEST INFRASTRUCTURE.  Mutation fuzzer for the host-side JPEG parser (csrc/gseg_jpeg.hpp: it reads
// untrusted files inside the product) and for the decoder's arithmetic (csrc/gseg_jpeg_core.h, driven the way the kernels
// drive it), meant to be built with -fsanitize=address,undefined: a wild read or write, a shift out of range or a signed
// overflow stops the program.  tests/test_jpeg.py generates the seed files, builds this and runs it for a few seconds.
//   usage: jpeg_fuzz_host <iterations> <seed> file...
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <vector>

#include "../graph-algorithm-image-segmentation-gpgpu_b200/csrc/gseg_jpeg.hpp"

static std::vector<uint8_t> read_file(const char *path) {
    std::vector<uint8_t> v;
    FILE *f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize((size_t)n);
    if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

// What the kernels do with a parsed file, on exactly sized buffers (so that the sanitizer sees every index).
static void decode(const std::vector<uint8_t> &file, const JpegPlan &plan, bool sync, uint32_t sub, long *decoded) {
    const JpegDev &d = plan.dev;
    if (d.nblocks > 40000 || d.w > 4096 || d.h > 4096) return; // a mutated frame header can ask for gigabytes
    std::vector<uint8_t> staged(file.size() + 64, 0); // the device buffer's padding: the reader loads 16-byte chunks, one ahead
    memcpy(staged.data(), file.data(), file.size());
    std::vector<int16_t> coef((size_t)d.nblocks * 64, 0);
    std::vector<uint8_t> samples((size_t)d.nsamples);
    std::vector<uint8_t> rgb((size_t)3 * d.w * d.h);
    uint32_t err = 0;
    uint32_t ring[8];
    if (sync) {
        const uint32_t off = d.data_off, end = d.data_end;
        const uint32_t nsub = end > off ? (end - off + sub - 1) / sub : 1;
        std::vector<uint64_t> entry(nsub), ex(nsub), prev;
        std::vector<uint32_t> nblk(nsub), blk0(nsub);
        auto endbits = [&](uint32_t i) { const uint64_t e = (uint64_t)off + (uint64_t)(i + 1) * sub; return (uint32_t)((e < end ? e : end) * 8u); };
        for (uint32_t i = 0; i < nsub; ++i) {
            entry[i] = i == 0 ? JPG_STATE(off * 8u, 0, 0) : jpg_sub_guess(staged.data(), off + i * sub, off);
            ex[i] = jpg_sub_decode<false>(d, d.dc, d.ac, jpg_zigzag_h, staged.data(), ring, 1u, entry[i], endbits(i), nullptr, 0u, &nblk[i], err);
        }
        uint32_t rounds = 0;
        for (bool ch = true; ch && rounds <= nsub + 2; ++rounds) {
            ch = false;
            prev = ex;
            for (uint32_t i = 1; i < nsub; ++i)
                if (prev[i - 1] != entry[i]) {
                    entry[i] = prev[i - 1];
                    ex[i] = jpg_sub_decode<false>(d, d.dc, d.ac, jpg_zigzag_h, staged.data(), ring, 1u, entry[i], endbits(i), nullptr, 0u, &nblk[i], err);
                    ch = true;
                }
        }
        if (rounds > nsub + 2) { fprintf(stderr, "the iteration did not end\n"); abort(); }
        uint32_t run = 0;
        for (uint32_t i = 0; i < nsub; ++i) { blk0[i] = run; run += nblk[i]; }
        for (uint32_t i = 0; i < nsub; ++i) {
            uint32_t dummy;
            jpg_sub_decode<true>(d, d.dc, d.ac, jpg_zigzag_h, staged.data(), ring, 1u, entry[i], endbits(i), coef.data(), blk0[i], &dummy, err);
        }
        for (int c = 0; c < d.ncomp; ++c) {
            const uint32_t nb = (uint32_t)d.nmcu * (uint32_t)(d.hs[c] * d.vs[c]);
            int pred = 0;
            for (uint32_t t = 0; t < nb; ++t) {
                int16_t *p = coef.data() + jpg_comp_block(d, c, t) * 64;
                if (t % ((uint32_t)d.ri * (uint32_t)(d.hs[c] * d.vs[c])) == 0) pred = 0;
                pred = (int16_t)(pred + p[0]);
                p[0] = (int16_t)pred;
            }
        }
    } else {
        if ((int)plan.starts.size() != d.nint) return;
        for (int i = 0; i < d.nint; ++i) {
            const int first = i * d.ri, last = first + d.ri < d.nmcu ? first + d.ri : d.nmcu;
            jpg_decode_interval(d, d.dc, d.ac, jpg_zigzag_h, staged.data(), ring, 1u, plan.starts[(size_t)i], first, last, coef.data(), err);
        }
    }
    for (int c = 0; c < d.ncomp; ++c)
        for (int by = 0; by < d.bh[c]; ++by)
            for (int bx = 0; bx < d.bw[c]; ++bx)
                jpg_idct_block(coef.data() + ((size_t)d.blk_off[c] + (size_t)by * d.bw[c] + bx) * 64, d.quant[c],
                               samples.data() + d.pix_off[c] + (size_t)by * 8 * d.bw[c] * 8 + bx * 8, d.bw[c] * 8);
    for (int y = 0; y < d.h; ++y)
        for (int x0 = 0; x0 < d.w; x0 += 8) {
            const int m = d.w - x0 < 8 ? d.w - x0 : 8;
            uint8_t *dst = rgb.data() + ((size_t)y * d.w + x0) * 3;
            if (m == 8 && jpg_fast8(d)) jpg_pixels8(d, samples.data(), x0, y, dst);
            else
                for (int j = 0; j < m; ++j) jpg_pixel(d, samples.data(), x0 + j, y, dst + 3 * j);
        }
    ++*decoded;
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const long iters = atol(argv[1]);
    std::mt19937 rng((unsigned)atol(argv[2]));
    std::vector<std::vector<uint8_t>> seeds;
    for (int i = 3; i < argc; ++i) {
        seeds.push_back(read_file(argv[i]));
        if (seeds.back().empty()) { fprintf(stderr, "cannot read %s\n", argv[i]); return 2; }
    }
    long parsed = 0, decoded = 0, rejected = 0;
    JpegPlan plan;
    for (long it = 0; it < iters; ++it) {
        std::vector<uint8_t> f = seeds[rng() % seeds.size()];
        const unsigned kind = rng() % 8;
        size_t sos = f.size();
        for (size_t i = 0; i + 1 < f.size(); ++i)
            if (f[i] == 0xFF && f[i + 1] == 0xDA) { sos = i; break; }
        const int nmut = kind == 0 ? 0 : 1 + (int)(rng() % 6);
        for (int m = 0; m < nmut; ++m) {
            // half of the mutations hit the headers (tables, frame, scan), the rest the entropy-coded data
            const size_t lim = (kind & 1) ? (sos + 14 < f.size() ? sos + 14 : f.size()) : f.size();
            const size_t p = rng() % lim;
            switch (rng() % 4) {
            case 0: f[p] = (uint8_t)rng(); break;
            case 1: f[p] ^= (uint8_t)(1u << (rng() % 8)); break;
            case 2: f[p] = 0xFF; break;
            default: f[p] = (uint8_t)(rng() % 4 ? 0 : 0xD0 + rng() % 16); break;
            }
        }
        if (kind == 7) f.resize(rng() % (f.size() + 1)); // truncation
        int w = 0, h = 0;
        jpeg_peek_size(f.data(), f.size(), &w, &h);
        const bool sync = (rng() & 1) != 0;
        const int rc = jpeg_parse(f.data(), f.size(), plan, !sync);
        if (rc != JPG_OK) { ++rejected; continue; }
        ++parsed;
        const uint32_t subs[5] = {8, 13, 32, 128, 1024};
        decode(f, plan, sync, subs[rng() % 5], &decoded);
    }
    printf("%ld files: %ld rejected by the parser, %ld parsed, %ld decoded\n", iters, rejected, parsed, decoded);
    return 0;
}
