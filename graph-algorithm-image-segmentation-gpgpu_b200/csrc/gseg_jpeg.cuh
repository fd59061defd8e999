// gseg_jpeg.cuh -- kernels of the in-house baseline-JPEG decoder ("GPU-side decode feeding
// batched mode -- removes the host staging bound"; the reference reads its JPEG data set with cv::imread on the host,
// README.md:26).  The compressed file crosses PCIe (7-10x fewer bytes than the RGB image) and is decoded here into
// the context's staged RGB buffer (or a caller's), on the context's stream or the pool's copy stream, in front of the blur:
//   entropy decoding, one of two ways (gseg_api.cu chooses by the file's restart interval):
//     k_jpeg_scan    one block: offsets of the restart intervals (the RSTn markers of the entropy-coded segment)
//     k_jpeg_huff    one thread per restart interval: Huffman decoding (T.81 F.2.2) of its MCUs into the coefficient
//                    array (all zero between images; only non-zero coefficients are written)
//   or, for files without restart markers and for long intervals,
//     k_jpeg_sync    one cluster (k_jpeg_sync_grid: a grid of small blocks): self-synchronising sub-sequences
//     k_jpeg_dcscan  DC differences -> values (segmented prefix sums per component)
//   k_jpeg_idct  one thread per 8x8 block: dequantisation + libjpeg's accurate integer inverse DCT -> sample planes;
//                clears the block's coefficients behind it
//   k_jpeg_rgb   eight pixels per thread: chroma upsampling ("fancy" triangle filter) + YCbCr -> interleaved RGB
//   k_jpeg_flag  hands a decoding error to the run's control block
// The arithmetic lives in gseg_jpeg_core.h (shared with the CPU test harness) and reproduces libjpeg's default decoder
// bit for bit (tests/test_jpeg.py against cv2.imdecode).
#pragma once
#include <cooperative_groups.h>

#include "gseg_device.cuh"
#include "gseg_jpeg_core.h"

namespace cg = cooperative_groups;

#define JPG_NT_HUFF 512 // big blocks on purpose: a block lives for the decode's whole latency, and an SM with a resident block
                        // cannot take a CTA of the tail cluster (1024 threads x 64 registers = the whole register file)
#define JPG_NT 128

__constant__ uint8_t c_jpg_zigzag[JPG_ZIGZAG_LEN] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

// Restart markers of the entropy-coded segment: starts[i] = offset of the first byte of interval i (right behind the
// i-th RSTn marker; T.81 B.2.1 / E.1.4).  In entropy-coded data a 0xFF is always followed by 0x00 unless it starts a
// marker, so the byte pairs FF D0..D7 are exactly the restart markers.  ONE block: its 32 warps take 32 contiguous
// parts of the data; pass 1 counts the markers of every part, a scan of the 32 counts gives each warp its first
// interval number, pass 2 writes the offsets in file order (lane prefix by shuffles).  Sixteen bytes per lane and step;
// a word is looked at byte by byte only if it holds a 0xFF (one word in ~64).
#define JPG_NT_SCAN 1024
#define JPG_SCAN_MLP 4
__device__ __forceinline__ uint32_t jpg_ff_bytes(uint32_t w) { // 0x80 in every byte of w that is 0xFF
    const uint32_t x = ~w;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// bit j set <=> bytes j, j + 1 of the chunk (byte 16 = `next`, the first byte behind it) are a restart marker that
// lies inside [off, end)
__device__ __forceinline__ uint32_t jpg_rst_mask(const uint4 v, uint32_t next, uint32_t pos0, uint32_t off, uint32_t end) {
    const uint32_t w[5] = {v.x, v.y, v.z, v.w, next};
    uint32_t mask = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t f = jpg_ff_bytes(w[i]);
        if (f) {
            const uint64_t two = ((uint64_t)w[i + 1] << 32) | w[i]; // the word and the bytes behind it
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (f & (0x80u << (8 * k))) {
                    const uint32_t m = (uint32_t)(two >> (8 * k + 8)) & 0xFFu;
                    const uint32_t p = pos0 + 4 * i + k;
                    if ((m & 0xF8u) == 0xD0u && p >= off && p + 1 < end) mask |= 1u << (4 * i + k);
                }
        }
    }
    return mask;
}
__global__ void __launch_bounds__(JPG_NT_SCAN) k_jpeg_scan(const JpegDev *__restrict__ gd, const uint8_t *__restrict__ file,
                                                           uint32_t *__restrict__ starts, uint32_t *__restrict__ errp) {
    __shared__ uint32_t s_cnt[32], s_off[33];
    const uint32_t off = gd->data_off, end = gd->data_end;
    const int nint = gd->nint;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { starts[0] = off; *errp = 0u; }
    if (nint <= 1) return;
    const uint32_t c0 = off >> 4, c1 = (end + 15u) >> 4;                 // chunks [c0, c1)
    const uint32_t per = (((c1 - c0 + 31u) / 32u) + 31u) / 32u * 32u;    // chunks per warp: whole steps of 32
    const uint32_t wb = c0 + (uint32_t)warp * per, we = min(wb + per, c1);
    const uint4 *f4 = reinterpret_cast<const uint4 *>(file);
    uint32_t cnt = 0u;
    for (uint32_t cb = wb; cb < we; cb += 32u * JPG_SCAN_MLP) { // JPG_SCAN_MLP loads in flight per lane: the loop is latency-bound
        uint4 v[JPG_SCAN_MLP];
#pragma unroll
        for (int u = 0; u < JPG_SCAN_MLP; ++u) {
            const uint32_t c = cb + 32u * u + lane;
            v[u] = c < we ? __ldg(f4 + c) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < JPG_SCAN_MLP; ++u) {
            const uint32_t c = cb + 32u * u + lane;
            if (jpg_ff_bytes(v[u].x) | jpg_ff_bytes(v[u].y) | jpg_ff_bytes(v[u].z) | jpg_ff_bytes(v[u].w))
                cnt += __popc(jpg_rst_mask(v[u], __ldg(reinterpret_cast<const uint32_t *>(f4 + c + 1)), 16u * c, off, end));
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_cnt[lane], inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        s_off[lane] = inc - v;
        if (lane == 31) {
            s_off[32] = inc;
            if ((int)inc + 1 < nint) *errp = JPG_ERR_RST;
        }
    }
    __syncthreads();
    const uint32_t total = s_off[32];
    // intervals the file has no marker for start at the end of the data (they decode to nothing and are flagged above)
    for (uint32_t i = total + 1 + threadIdx.x; i < (uint32_t)nint; i += JPG_NT_SCAN) starts[i] = end;
    uint32_t running = s_off[warp];
    for (uint32_t cb0 = wb; cb0 < we; cb0 += 32u * JPG_SCAN_MLP) {
        uint4 v[JPG_SCAN_MLP];
#pragma unroll
        for (int u = 0; u < JPG_SCAN_MLP; ++u) {
            const uint32_t c = cb0 + 32u * u + lane;
            v[u] = c < we ? __ldg(f4 + c) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < JPG_SCAN_MLP; ++u) { // file order: step by step, lane by lane, bit by bit
            const uint32_t c = cb0 + 32u * u + lane;
            uint32_t mask = 0u;
            if (jpg_ff_bytes(v[u].x) | jpg_ff_bytes(v[u].y) | jpg_ff_bytes(v[u].z) | jpg_ff_bytes(v[u].w))
                mask = jpg_rst_mask(v[u], __ldg(reinterpret_cast<const uint32_t *>(f4 + c + 1)), 16u * c, off, end);
            if (__ballot_sync(0xFFFFFFFFu, mask != 0u) == 0u) continue;
            const uint32_t n = __popc(mask);
            uint32_t inc = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
            uint32_t idx = running + inc - n + 1u; // interval number of the lane's first marker
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1u;
                if (idx < (uint32_t)nint) starts[idx] = 16u * c + (uint32_t)j + 2u;
                ++idx;
            }
            running += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    }
}

__global__ void __launch_bounds__(JPG_NT_HUFF) k_jpeg_huff(const JpegDev *__restrict__ gd, const uint32_t *__restrict__ starts,
                                                           const uint8_t *__restrict__ file, int16_t *__restrict__ coef,
                                                           uint32_t *__restrict__ errp) {
    __shared__ JpegDev sd; // geometry + the six Huffman tables (15 KB): a symbol is one shared-memory look-up, three for a long code
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(gd);
        uint4 *dst = reinterpret_cast<uint4 *>(&sd);
        for (int i = threadIdx.x; i < (int)(sizeof(JpegDev) / 16); i += JPG_NT_HUFF) dst[i] = src[i];
    }
    __shared__ uint8_t szz[JPG_ZIGZAG_LEN];
    __shared__ uint32_t s_ring[8 * JPG_NT_HUFF]; // the readers' chunk rings: word k of thread t at [k * JPG_NT_HUFF + t], its own bank
    for (int i = threadIdx.x; i < JPG_ZIGZAG_LEN; i += JPG_NT_HUFF) szz[i] = c_jpg_zigzag[i];
    __syncthreads();
    const int i = blockIdx.x * JPG_NT_HUFF + threadIdx.x;
    if (i >= sd.nint) return;
    const int first = i * sd.ri;
    const int last = first + sd.ri < sd.nmcu ? first + sd.ri : sd.nmcu;
    uint32_t err = 0u;
    jpg_decode_interval(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_HUFF, starts[i], first, last, coef, err);
    if (err) atomicOr(errp, err);
}

// ---- files without restart markers: the self-synchronising sub-sequence decode (gseg_jpeg_core.h) in ONE launch of a
// single thread-block cluster (hardware co-scheduled, barrier.cluster between the passes -- like k_tail):
//   pass 1   every thread decodes its sub-sequence(s) from the guessed entry state and publishes the exit state;
//   rounds   a thread whose predecessor's exit differs from the entry it used decodes again; a round in which nobody
//            decoded ends the iteration (flag words rotate over three slots so that a reset never races a reader);
//   scan     blocks completed per sub-sequence -> number of the block each sub-sequence starts in (block 0 of the cluster);
//   write    every thread decodes once more, now storing the coefficients (DC as differences).
// State words are 64-bit and read with ld.cg: a reader sees the old or the new state of its neighbour, never a mix.
#define JPG_NT_SYNC 1024
__global__ void __launch_bounds__(JPG_NT_SYNC, 1) k_jpeg_sync(const JpegDev *__restrict__ gd, const uint8_t *__restrict__ file,
                                                              uint64_t *entryS, uint64_t *exitS, uint32_t *nblk, uint32_t *blk0,
                                                              uint32_t *flags, int16_t *__restrict__ coef, uint32_t *errp, uint32_t S) {
    __shared__ JpegDev sd;
    __shared__ uint8_t szz[JPG_ZIGZAG_LEN];
    __shared__ uint32_t s_ring[8 * JPG_NT_SYNC];
    __shared__ uint32_t s_part[32];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(gd);
        uint4 *dst = reinterpret_cast<uint4 *>(&sd);
        for (int i = threadIdx.x; i < (int)(sizeof(JpegDev) / 16); i += JPG_NT_SYNC) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < JPG_ZIGZAG_LEN; i += JPG_NT_SYNC) szz[i] = c_jpg_zigzag[i];
    __syncthreads();
    cg::cluster_group cl = cg::this_cluster();
    const uint32_t T = gridDim.x * JPG_NT_SYNC, tid = blockIdx.x * JPG_NT_SYNC + threadIdx.x;
    const uint32_t off = sd.data_off, end = sd.data_end;
    const uint32_t nsub = end > off ? (end - off + S - 1u) / S : 1u;
    if (tid == 0) { flags[0] = 0u; flags[1] = 0u; flags[2] = 0u; *errp = 0u; }
    uint32_t err = 0u, nb = 0u;
    for (uint32_t i = tid; i < nsub; i += T) { // pass 1
        const uint32_t e1 = min(off + (i + 1u) * S, end) * 8u;
        const uint64_t en = i == 0u ? JPG_STATE(off * 8u, 0, 0) : jpg_sub_guess(file, off + i * S, off);
        entryS[i] = en;
        exitS[i] = jpg_sub_decode<false>(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_SYNC, en, e1, nullptr, 0u, &nb, err);
        nblk[i] = nb;
    }
    __threadfence();
    cl.sync();
    for (uint32_t r = 0;; ++r) { // rounds
        if (tid == 0) flags[(r + 1u) % 3u] = 0u;
        int ch = 0;
        for (uint32_t i = tid; i < nsub; i += T) {
            if (i == 0u) continue;
            const uint64_t en = __ldcg(exitS + i - 1u);
            if (en != entryS[i]) {
                const uint32_t e1 = min(off + (i + 1u) * S, end) * 8u;
                entryS[i] = en;
                exitS[i] = jpg_sub_decode<false>(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_SYNC, en, e1, nullptr, 0u, &nb, err);
                nblk[i] = nb;
                ch = 1;
            }
        }
        if (__syncthreads_or(ch) && threadIdx.x == 0) atomicOr(&flags[r % 3u], 1u);
        __threadfence();
        cl.sync();
        if (__ldcg(&flags[r % 3u]) == 0u) break; // the same word for every thread of the cluster
        if (r > nsub + 2u) { if (tid == 0) atomicOr(errp, JPG_ERR_BLOCKS); break; } // cannot happen: round r fixes sub-sequence r
    }
    if (blockIdx.x == 0) { // exclusive scan of the block counts
        const uint32_t per = (nsub + JPG_NT_SYNC - 1u) / JPG_NT_SYNC;
        const uint32_t b = threadIdx.x * per, e = min(b + per, nsub);
        uint32_t sum = 0u;
        for (uint32_t i = b; i < e; ++i) sum += __ldcg(nblk + i);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t v = s_part[lane], w2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w2, o); if (lane >= o) w2 += t; }
            s_part[lane] = w2 - v;
            if (lane == 31 && w2 < (uint32_t)sd.nblocks) atomicOr(errp, JPG_ERR_BLOCKS); // (padding behind the last block may decode to more)
        }
        __syncthreads();
        uint32_t run = s_part[warp] + inc - sum;
        for (uint32_t i = b; i < e; ++i) { blk0[i] = run; run += __ldcg(nblk + i); }
    }
    __threadfence();
    cl.sync();
    err = 0u; // the passes above ran from guessed states: their impossible codes mean nothing
    for (uint32_t i = tid; i < nsub; i += T) { // write
        const uint32_t e1 = min(off + (i + 1u) * S, end) * 8u;
        jpg_sub_decode<true>(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_SYNC, entryS[i], e1, coef, __ldcg(blk0 + i), &nb, err);
    }
    if (err) atomicOr(errp, err);
}

// The same algorithm as a plain grid of SMALL blocks with a software grid barrier (one atomic counter; every block's
// thread 0 adds itself and spins until all have).  A cluster block of 1024 threads x 58 registers needs an SM to itself,
// and five of them at the same moment: with eight images in flight the block scheduler rarely has that, and the decode
// waits; 128-thread blocks fit anywhere.  All blocks of the grid become resident without anybody's help -- the kernels
// they share the GPU with are short and never wait for them -- so the barrier cannot deadlock; a bounded spin turns a
// scheduling surprise into an error code instead of a hung GPU (like the look-back watchdog of the segmentation).
#define JPG_NT_GRID 128
__device__ __forceinline__ bool jpg_grid_barrier(uint32_t *bar, uint32_t *abortp, uint32_t &gen, uint32_t nblocks, uint32_t *s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gen;
        __threadfence();
        atomicAdd(bar, 1u);
        const uint32_t target = gen * nblocks;
        uint32_t spins = 0u, ab = 0u;
        while (*reinterpret_cast<volatile uint32_t *>(bar) < target) {
            ab = *reinterpret_cast<volatile uint32_t *>(abortp);
            if (ab) break;
            if (++spins > (1u << 22)) { atomicExch(abortp, 1u); ab = 1u; break; } // ~1 s
            __nanosleep(spins < 64u ? 50u : 250u);
        }
        __threadfence();
        *s_flag = ab;
    }
    __syncthreads();
    return *s_flag == 0u;
}
__global__ void __launch_bounds__(JPG_NT_GRID) k_jpeg_sync_grid(const JpegDev *__restrict__ gd, const uint8_t *__restrict__ file,
                                                                uint64_t *entryS, uint64_t *exitS, uint32_t *nblk, uint32_t *blk0,
                                                                uint32_t *flags, int16_t *__restrict__ coef, uint32_t *errp, uint32_t S) {
    __shared__ JpegDev sd;
    __shared__ uint8_t szz[JPG_ZIGZAG_LEN];
    __shared__ uint32_t s_ring[8 * JPG_NT_GRID];
    __shared__ uint32_t s_part[32], s_flag;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(gd);
        uint4 *dst = reinterpret_cast<uint4 *>(&sd);
        for (int i = threadIdx.x; i < (int)(sizeof(JpegDev) / 16); i += JPG_NT_GRID) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < JPG_ZIGZAG_LEN; i += JPG_NT_GRID) szz[i] = c_jpg_zigzag[i];
    if (threadIdx.x < 32) s_part[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t *bar = flags + 3, *abortp = flags + 4; // flags[0..4] were cleared by the host in front of the launch
    uint32_t gen = 0u;
    const uint32_t T = gridDim.x * JPG_NT_GRID, tid = blockIdx.x * JPG_NT_GRID + threadIdx.x;
    const uint32_t off = sd.data_off, end = sd.data_end;
    const uint32_t nsub = end > off ? (end - off + S - 1u) / S : 1u;
    if (tid == 0) *errp = 0u;
    uint32_t err = 0u, nb = 0u;
    for (uint32_t i = tid; i < nsub; i += T) { // pass 1
        const uint32_t e1 = min(off + (i + 1u) * S, end) * 8u;
        const uint64_t en = i == 0u ? JPG_STATE(off * 8u, 0, 0) : jpg_sub_guess(file, off + i * S, off);
        entryS[i] = en;
        exitS[i] = jpg_sub_decode<false>(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_GRID, en, e1, nullptr, 0u, &nb, err);
        nblk[i] = nb;
    }
    if (!jpg_grid_barrier(bar, abortp, gen, gridDim.x, &s_flag)) { if (tid == 0) atomicOr(errp, JPG_ERR_BLOCKS); return; }
    for (uint32_t r = 0;; ++r) { // rounds
        if (tid == 0) flags[(r + 1u) % 3u] = 0u;
        int ch = 0;
        for (uint32_t i = tid; i < nsub; i += T) {
            if (i == 0u) continue;
            const uint64_t en = __ldcg(exitS + i - 1u);
            if (en != entryS[i]) {
                const uint32_t e1 = min(off + (i + 1u) * S, end) * 8u;
                entryS[i] = en;
                exitS[i] = jpg_sub_decode<false>(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_GRID, en, e1, nullptr, 0u, &nb, err);
                nblk[i] = nb;
                ch = 1;
            }
        }
        if (__syncthreads_or(ch) && threadIdx.x == 0) atomicOr(&flags[r % 3u], 1u);
        if (!jpg_grid_barrier(bar, abortp, gen, gridDim.x, &s_flag)) { if (tid == 0) atomicOr(errp, JPG_ERR_BLOCKS); return; }
        if (__ldcg(&flags[r % 3u]) == 0u) break; // the same word for every thread of the grid
        if (r > nsub + 2u) { if (tid == 0) atomicOr(errp, JPG_ERR_BLOCKS); break; } // cannot happen: round r fixes sub-sequence r
    }
    if (blockIdx.x == 0) { // exclusive scan of the block counts (one block: 128 threads, a contiguous run each)
        const uint32_t per = (nsub + JPG_NT_GRID - 1u) / JPG_NT_GRID;
        const uint32_t b = min(threadIdx.x * per, nsub), e = min(b + per, nsub);
        uint32_t sum = 0u;
        for (uint32_t i = b; i < e; ++i) sum += __ldcg(nblk + i);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t v = s_part[lane], w2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w2, o); if (lane >= o) w2 += t; }
            s_part[lane] = w2 - v;
            if (lane == 31 && w2 < (uint32_t)sd.nblocks) atomicOr(errp, JPG_ERR_BLOCKS);
        }
        __syncthreads();
        uint32_t run = s_part[warp] + inc - sum;
        for (uint32_t i = b; i < e; ++i) { blk0[i] = run; run += __ldcg(nblk + i); }
    }
    if (!jpg_grid_barrier(bar, abortp, gen, gridDim.x, &s_flag)) { if (tid == 0) atomicOr(errp, JPG_ERR_BLOCKS); return; }
    err = 0u;
    for (uint32_t i = tid; i < nsub; i += T) { // write
        const uint32_t e1 = min(off + (i + 1u) * S, end) * 8u;
        jpg_sub_decode<true>(sd, sd.dc, sd.ac, szz, file, s_ring + threadIdx.x, JPG_NT_GRID, entryS[i], e1, coef, __ldcg(blk0 + i), &nb, err);
    }
    if (err) atomicOr(errp, err);
}

// DC differences -> DC values (T.81 F.1.1.5.1: the predictor of a component runs over its blocks in scan order): one block
// per component; a thread sums a contiguous run of the component's blocks, a block-wide scan, a second walk writes the values.
// The walk keeps (MCU column, MCU row, block inside the MCU) as state -- one division per thread, not three per block (the
// first version spent 84 us of one SM on them) -- and has eight loads in flight (blocks are 128 bytes apart).
struct JpgWalk {
    uint32_t mx, my, bi, per, hs, sh, mcus_x, row, bw, base; // sh = log2(hs); row = vs * bw; base = blk_off
    __device__ __forceinline__ void init(const JpegDev &d, int c, uint32_t t) {
        per = (uint32_t)(d.hs[c] * d.vs[c]); hs = (uint32_t)d.hs[c]; sh = hs == 4u ? 2u : (hs == 2u ? 1u : 0u);
        mcus_x = (uint32_t)d.mcus_x; bw = (uint32_t)d.bw[c]; row = (uint32_t)d.vs[c] * bw; base = (uint32_t)d.blk_off[c];
        const uint32_t m = t / per;
        bi = t - m * per; mx = m % mcus_x; my = m / mcus_x;
    }
    __device__ __forceinline__ size_t index() const { // == jpg_comp_block
        const uint32_t bv = bi >> sh, bh = bi - (bv << sh);
        return (size_t)base + (size_t)my * row + (size_t)bv * bw + mx * hs + bh;
    }
    __device__ __forceinline__ void next() {
        if (++bi == per) { bi = 0; if (++mx == mcus_x) { mx = 0; ++my; } }
    }
};
__global__ void __launch_bounds__(JPG_NT_SYNC) k_jpeg_dcscan(const JpegDev *__restrict__ gd, int16_t *__restrict__ coef) {
    __shared__ __align__(16) uint32_t sgeo[offsetof(JpegDev, quant) / 4];
    __shared__ int s_val[32], s_flag[32];
    for (int i = threadIdx.x; i < (int)(offsetof(JpegDev, quant) / 4); i += JPG_NT_SYNC) sgeo[i] = reinterpret_cast<const uint32_t *>(gd)[i];
    __syncthreads();
    const JpegDev &d = *reinterpret_cast<const JpegDev *>(sgeo);
    const int c = blockIdx.x;
    if (c >= d.ncomp) return;
    const uint32_t bpc = (uint32_t)(d.hs[c] * d.vs[c]);
    const uint32_t n = (uint32_t)d.nmcu * bpc;
    const uint32_t seg = (uint32_t)d.ri * bpc; // the predictor restarts with every restart interval: a SEGMENTED prefix sum
    const uint32_t per = (n + JPG_NT_SYNC - 1u) / JPG_NT_SYNC;
    const uint32_t b = min(threadIdx.x * per, n), e = min(b + per, n);
    JpgWalk wk;
    wk.init(d, c, b);
    // the thread's run as (flag: a segment starts inside, val: sum since the run's last segment start)
    int val = 0, flag = 0;
    uint32_t left = seg - b % seg; // blocks until the next segment start (== seg: b itself is one)
    if (left == seg) left = 0u;
    for (uint32_t t0 = b; t0 < e; t0 += 8u) {
        int x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { x[u] = t0 + u < e ? (int)coef[wk.index() * 64] : 0; wk.next(); }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u < e) {
                if (left == 0u) { val = 0; flag = 1; left = seg; }
                val += x[u];
                --left;
            }
    }
    // inclusive scan over the threads with (f1, v1) + (f2, v2) = (f1 | f2, f2 ? v2 : v1 + v2)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int iv = val, ifl = flag;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int pv = __shfl_up_sync(0xFFFFFFFFu, iv, o), pf = __shfl_up_sync(0xFFFFFFFFu, ifl, o);
        if (lane >= o) { iv = ifl ? iv : iv + pv; ifl |= pf; }
    }
    if (lane == 31) { s_val[warp] = iv; s_flag[warp] = ifl; }
    __syncthreads();
    if (warp == 0) {
        int wv = s_val[lane], wf = s_flag[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int pv = __shfl_up_sync(0xFFFFFFFFu, wv, o), pf = __shfl_up_sync(0xFFFFFFFFu, wf, o);
            if (lane >= o) { wv = wf ? wv : wv + pv; wf |= pf; }
        }
        // exclusive over the warps: what the warps before this one carry into it
        const int ev = __shfl_up_sync(0xFFFFFFFFu, wv, 1), ef = __shfl_up_sync(0xFFFFFFFFu, wf, 1);
        s_val[lane] = lane ? ev : 0; s_flag[lane] = lane ? ef : 0;
    }
    __syncthreads();
    // carry into this thread = (warps before) + (lanes before in the warp), exclusive
    int cv = __shfl_up_sync(0xFFFFFFFFu, iv, 1), cf = __shfl_up_sync(0xFFFFFFFFu, ifl, 1);
    if (lane == 0) { cv = 0; cf = 0; }
    int pred = cf ? cv : cv + s_val[warp];
    wk.init(d, c, b);
    left = seg - b % seg;
    if (left == seg) left = 0u;
    for (uint32_t t0 = b; t0 < e; t0 += 8u) {
        int16_t *p[8];
        int x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { p[u] = coef + wk.index() * 64; x[u] = t0 + u < e ? (int)*p[u] : 0; wk.next(); }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u < e) {
                if (left == 0u) { pred = 0; left = seg; }
                pred += x[u];
                *p[u] = (int16_t)pred;
                --left;
            }
    }
}

// Reads a block's coefficients and clears them behind it: the coefficient array is all zero again when the kernel ends,
// which is what the next image's Huffman threads (they only write non-zero coefficients) need -- no memset per image.
__global__ void __launch_bounds__(JPG_NT) k_jpeg_idct(const JpegDev *__restrict__ gd, int16_t *__restrict__ coef,
                                                      uint8_t *__restrict__ samples) {
    __shared__ uint16_t sq[JPG_MAXCOMP][64];
    for (int i = threadIdx.x; i < JPG_MAXCOMP * 64; i += JPG_NT) sq[i >> 6][i & 63] = gd->quant[i >> 6][i & 63];
    __syncthreads();
    const int b = blockIdx.x * JPG_NT + threadIdx.x;
    if (b >= gd->nblocks) return;
    const int c = (gd->ncomp > 2 && b >= gd->blk_off[2]) ? 2 : ((gd->ncomp > 1 && b >= gd->blk_off[1]) ? 1 : 0);
    const int lb = b - gd->blk_off[c], bwc = gd->bw[c];
    const int by = lb / bwc, bx = lb - by * bwc;
    union { int4 v[8]; int16_t s[64]; } in;
    int4 *src = reinterpret_cast<int4 *>(coef + (size_t)b * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) in.v[i] = src[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) src[i] = make_int4(0, 0, 0, 0);
    union { uint2 v[8]; uint8_t s[64]; } out;
    jpg_idct_block(in.s, sq[c], out.s, 8);
    uint8_t *dst = samples + gd->pix_off[c] + ((size_t)by * 8 * bwc + bx) * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2 *>(dst + (size_t)r * bwc * 8) = out.v[r];
}

__global__ void __launch_bounds__(JPG_NT) k_jpeg_rgb(const JpegDev *__restrict__ gd, const uint8_t *__restrict__ samples,
                                                     uint8_t *__restrict__ rgb) {
    __shared__ __align__(16) uint32_t sgeo[offsetof(JpegDev, quant) / 4]; // the geometry in front of the tables is all that is read
    for (int i = threadIdx.x; i < (int)(offsetof(JpegDev, quant) / 4); i += JPG_NT) sgeo[i] = reinterpret_cast<const uint32_t *>(gd)[i];
    __syncthreads();
    const JpegDev &sdh = *reinterpret_cast<const JpegDev *>(sgeo);
    const int w = sdh.w, h = sdh.h;
    const int gpr = (w + 7) >> 3; // groups of eight pixels per row
    const size_t g = (size_t)blockIdx.x * JPG_NT + threadIdx.x;
    if (g >= (size_t)gpr * h) return;
    const int y = (int)(g / gpr), x0 = 8 * (int)(g - (size_t)y * gpr);
    const int n = w - x0 < 8 ? w - x0 : 8;
    uint8_t *dst = rgb + ((size_t)y * w + x0) * 3;
    if (n == 8 && jpg_fast8(sdh)) {
        union { uint32_t v[6]; uint8_t s[24]; } px;
        jpg_pixels8(sdh, samples, x0, y, px.s);
        if ((w & 3) == 0) { // rows start 4-byte aligned
            uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
            for (int k = 0; k < 6; ++k) d4[k] = px.v[k];
        } else {
#pragma unroll
            for (int k = 0; k < 24; ++k) dst[k] = px.s[k];
        }
    } else { // rows' tails and the sampling layouts without an eight-pixel form
        for (int j = 0; j < n; ++j) {
            uint8_t t[3];
            jpg_pixel(sdh, samples, x0 + j, y, t);
            dst[3 * j] = t[0]; dst[3 * j + 1] = t[1]; dst[3 * j + 2] = t[2];
        }
    }
}

__global__ void k_jpeg_flag(const uint32_t *__restrict__ errp, GsegCtl *ctl) {
    if (*errp && ctl->error == DERR_NONE) ctl->error = DERR_JPEG;
}
