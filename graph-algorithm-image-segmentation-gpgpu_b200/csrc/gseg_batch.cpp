// gseg_batch.cpp -- C++-only caller of the batch pipeline (gseg_pool_*): the reference's benchmark loop over a
// data set of equally sized images (Report.pdf p4 s4.1: N iterations, disk I/O excluded, transfers included;
// README.md:26-28), here with several images in flight.  No CUDA headers, no compute: it only calls libgseg.so.
//
//   gseg_batch --synth WxH [--n N] [--contexts S] [--steps K] [--warmup W] [--conn 4|8] [--variant felz|hier|superpix]
//              [--level L] [--seed S0] [--device D] [--int32] [--print-counts]
// Every step segments the same N pinned host images (synthetic, seeds S0..S0+N-1) and receives N label images in
// pinned host memory (narrowest lossless element type unless --int32).  Prints Mpixel/s end to end.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gseg.h"

int main(int argc, char **argv) {
    int w = 1920, h = 1080, n = 32, S = 8, steps = 10, warmup = 3, device = 0, level = -1, elem = 0;
    unsigned long long seed = 2000;
    bool counts = false;
    gseg_params p;
    memset(&p, 0, sizeof p);
    p.sigma = 0.8f; p.k = 300.0f; p.min_size = 20; p.connectivity = 4; p.variant = GSEG_FELZ;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "gseg_batch: %s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--synth") { if (sscanf(need(), "%dx%d", &w, &h) != 2) return 2; }
        else if (a == "--n") n = atoi(need());
        else if (a == "--contexts") S = atoi(need());
        else if (a == "--steps") steps = atoi(need());
        else if (a == "--warmup") warmup = atoi(need());
        else if (a == "--conn") p.connectivity = atoi(need());
        else if (a == "--level") level = atoi(need());
        else if (a == "--seed") seed = strtoull(need(), nullptr, 10);
        else if (a == "--device") device = atoi(need());
        else if (a == "--sigma") p.sigma = (float)atof(need());
        else if (a == "--k") p.k = (float)atof(need());
        else if (a == "--min") p.min_size = atoi(need());
        else if (a == "--int32") elem = 4;
        else if (a == "--print-counts") counts = true;
        else if (a == "--variant") {
            std::string v = need();
            p.variant = v == "hier" ? GSEG_HIER : v == "superpix" ? GSEG_SUPERPIX : GSEG_FELZ;
        } else { fprintf(stderr, "gseg_batch: unknown option %s\n", a.c_str()); return 2; }
    }
    if (w < 1 || h < 1 || n < 1 || S < 1 || steps < 1) return 2;
    gseg_pool *pool = nullptr;
    int rc = gseg_pool_create(&pool, device, w, h, p.connectivity, S, p.variant == GSEG_SUPERPIX ? GSEG_CAP_SUPERPIX : 0u);
    if (rc) { fprintf(stderr, "gseg_batch: gseg_pool_create: %s\n", gseg_strerror(rc)); return 1; }
    const size_t V = (size_t)w * h;
    uint8_t *in = (uint8_t *)gseg_host_alloc((size_t)n * V * 3);
    uint8_t *out = (uint8_t *)gseg_host_alloc((size_t)n * V * 4);
    if (!in || !out) { fprintf(stderr, "gseg_batch: pinned allocation failed\n"); return 1; }
    for (int i = 0; i < n; ++i) {
        rc = gseg_synth(gseg_pool_context(pool, 0), in + (size_t)i * V * 3, w, h, seed + (unsigned long long)i, GSEG_MEM_HOST);
        if (rc) { fprintf(stderr, "gseg_batch: gseg_synth: %s\n", gseg_strerror(rc)); return 1; }
    }
    std::vector<gseg_pool_job> jobs((size_t)n);
    std::vector<gseg_pool_result> res((size_t)n);
    for (int i = 0; i < n; ++i) {
        gseg_pool_job &j = jobs[(size_t)i];
        memset(&j, 0, sizeof j);
        j.input = in + (size_t)i * V * 3; j.w = w; j.h = h; j.stride_bytes = 3 * w; j.mem_kind = GSEG_MEM_HOST;
        j.params = p;
        j.out_mode = GSEG_OUT_LABELS; j.level = level; j.elem_bytes = elem; j.out_mem_kind = GSEG_MEM_HOST;
        j.out = out + (size_t)i * V * 4; j.out_capacity = V * 4;
    }
    for (int s = 0; s < warmup; ++s) {
        rc = gseg_pool_run(pool, jobs.data(), n, res.data());
        if (rc) { fprintf(stderr, "gseg_batch: gseg_pool_run: %s (%s)\n", gseg_strerror(rc), gseg_pool_last_error(pool)); return 1; }
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; ++s) {
        rc = gseg_pool_run(pool, jobs.data(), n, res.data());
        if (rc) { fprintf(stderr, "gseg_batch: gseg_pool_run: %s (%s)\n", gseg_strerror(rc), gseg_pool_last_error(pool)); return 1; }
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;
    long long d2h = 0;
    for (int i = 0; i < n; ++i) d2h += res[(size_t)i].out_bytes;
    double ceil_ms = 0;
    rc = gseg_pool_copy_ceiling(pool, jobs.data(), res.data(), n, 3, &ceil_ms);
    printf("e2e %.1f Mpixel/s: %d images of %dx%d per step, %d contexts, %.3f ms/step, h2d %lld B/step, d2h %lld B/step",
           (double)n * V / 1e3 / ms, n, w, h, S, ms, (long long)n * (long long)V * 3, d2h);
    if (!rc) printf(", copies alone %.3f ms/step (%.1f Mpixel/s)", ceil_ms, (double)n * V / 1e3 / ceil_ms);
    printf("\n");
    if (counts) {
        rc = gseg_pool_run(pool, jobs.data(), n, res.data());
        if (rc) return 1;
        printf("counts:");
        for (int i = 0; i < n; ++i) printf(" %d", res[(size_t)i].n_components);
        printf("\n");
    }
    gseg_pool_destroy(pool);
    gseg_host_free(in);
    gseg_host_free(out);
    return 0;
}
