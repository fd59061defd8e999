// Probe: which nvJPEG backends decode a batch of baseline JPEGs on this GPU, and how fast (nvjpegDecodeBatched, RGBI out).
// Build on the box: nvcc -O2 -o gpurun_out/nvjpeg_probe tools/micro/nvjpeg_probe.cu -lnvjpeg ; run: nvjpeg_probe dir_with_jpgs
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <dirent.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <chrono>
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::vector<std::vector<unsigned char>> files;
    DIR *d = opendir(argv[1]);
    if (!d) return 2;
    while (dirent *e = readdir(d)) {
        std::string n = e->d_name;
        if (n.size() < 4 || n.substr(n.size() - 4) != ".jpg") continue;
        FILE *f = fopen((std::string(argv[1]) + "/" + n).c_str(), "rb");
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<unsigned char> b(sz); if (fread(b.data(), 1, sz, f) != (size_t)sz) return 2; fclose(f);
        files.push_back(b);
    }
    const int N = (int)files.size();
    printf("%d files\n", N);
    const char *names[] = {"DEFAULT", "HYBRID", "GPU_HYBRID", "HARDWARE", "GPU_HYBRID_DEVICE", "HARDWARE_DEVICE"};
    for (int be : {3, 2, 0}) {
        nvjpegHandle_t h; nvjpegJpegState_t st;
        nvjpegStatus_t rc = nvjpegCreateEx((nvjpegBackend_t)be, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &h);
        if (rc != NVJPEG_STATUS_SUCCESS) { printf("backend %s: nvjpegCreateEx failed (%d)\n", names[be], (int)rc); continue; }
        if (nvjpegJpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) { printf("state create failed\n"); continue; }
        int nc, ws[4], hs[4]; nvjpegChromaSubsampling_t ss;
        nvjpegGetImageInfo(h, files[0].data(), files[0].size(), &nc, &ss, ws, hs);
        const int w = ws[0], hh = hs[0];
        rc = nvjpegDecodeBatchedInitialize(h, st, N, 1, NVJPEG_OUTPUT_RGBI);
        if (rc != NVJPEG_STATUS_SUCCESS) { printf("backend %s: DecodeBatchedInitialize failed (%d)\n", names[be], (int)rc); continue; }
        std::vector<nvjpegImage_t> outs(N);
        std::vector<const unsigned char *> ptrs(N); std::vector<size_t> lens(N);
        for (int i = 0; i < N; ++i) {
            outs[i] = nvjpegImage_t();
            cudaMalloc((void **)&outs[i].channel[0], (size_t)3 * w * hh); outs[i].pitch[0] = 3 * w;
            ptrs[i] = files[i].data(); lens[i] = files[i].size();
        }
        cudaStream_t s; cudaStreamCreate(&s);
        double best = 1e9; bool ok = true;
        for (int rep = 0; rep < 5 && ok; ++rep) {
            cudaStreamSynchronize(s);
            auto t0 = std::chrono::steady_clock::now();
            rc = nvjpegDecodeBatched(h, st, ptrs.data(), lens.data(), outs.data(), s);
            if (rc != NVJPEG_STATUS_SUCCESS) { printf("backend %s: DecodeBatched failed (%d)\n", names[be], (int)rc); ok = false; break; }
            cudaStreamSynchronize(s);
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rep) best = dt < best ? dt : best;
        }
        if (ok) printf("backend %s: %d x %dx%d (subsampling %d): %.2f ms per batch = %.0f images/s = %.0f Mpixel/s\n", names[be], N, w, hh, (int)ss, best * 1e3, N / best, N * (double)w * hh / 1e6 / best);
        for (int i = 0; i < N; ++i) cudaFree(outs[i].channel[0]);
        nvjpegJpegStateDestroy(st); nvjpegDestroy(h);
    }
    return 0;
}
