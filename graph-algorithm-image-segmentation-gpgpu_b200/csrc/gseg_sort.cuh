// gseg_sort.cuh -- in-house onesweep radix sort (64-bit keys + 32-bit payload), LSD, 8-bit digits.
//
// Replaces the reference's thrust::sort on packed 64-bit edge keys (Report.pdf p3 s3.2.2 "bit
// concatenation ... single 64 bit integer").  One read of the keys builds
// the histograms of every digit position; then one kernel per digit does a single pass: per-tile
// digit counts -> chained (decoupled look-back) scan per digit -> stable in-tile ranking with
// __match_any_sync -> reorder through shared memory -> coalesced run-wise scatter.
// Algorithmic bytes: 8 n (histogram) + passes x 2 x 12 n.
#pragma once
#include "gseg_device.cuh"

#define SORT_NT 256
#define SORT_KPT 16
#define SORT_TILE (SORT_NT * SORT_KPT)
#define SORT_RADIX 256
#define SORT_MAXPASS 8
#define SORT_LB 8 /* predecessors fetched per round trip of the per-digit look-back */

struct SortScratch {
    u64 *keys_alt;
    u32 *vals_alt;
    u32 *hist;    // [SORT_MAXPASS][256] global digit histograms -> exclusive offsets
    u32 *status;  // [passes][ntiles][256] look-back words
    u32 *tickets; // [SORT_MAXPASS]
    size_t cap_n, cap_status;
    bool attr_set;
};

static void sort_scratch_free(SortScratch *s) {
    cudaFree(s->keys_alt); cudaFree(s->vals_alt); cudaFree(s->hist); cudaFree(s->status); cudaFree(s->tickets);
    memset(s, 0, sizeof(*s));
}

__global__ void __launch_bounds__(SORT_NT) k_sort_hist(const u64 *__restrict__ keys, size_t n, int begin_bit, int end_bit,
                                                       u32 *__restrict__ hist) {
    __shared__ u32 sh[SORT_MAXPASS * SORT_RADIX];
    const int npass = (end_bit - begin_bit + 7) / 8;
    for (int i = threadIdx.x; i < npass * SORT_RADIX; i += SORT_NT) sh[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * SORT_NT + threadIdx.x; i < n; i += (size_t)gridDim.x * SORT_NT) {
        const u64 k = keys[i];
        for (int p = 0; p < npass; ++p) {
            const int lo = begin_bit + 8 * p;
            const int bits = min(8, end_bit - lo);
            atomicAdd(&sh[p * SORT_RADIX + (u32)((k >> lo) & ((1u << bits) - 1u))], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * SORT_RADIX; i += SORT_NT)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of each pass's 256 bins; one block per pass
__global__ void __launch_bounds__(SORT_RADIX) k_sort_scan(u32 *hist) {
    __shared__ u32 s[34];
    u32 *h = hist + blockIdx.x * SORT_RADIX;
    const u32 v = h[threadIdx.x];
    const u32 ex = block_excl_scan<SORT_RADIX>(v, s);
    h[threadIdx.x] = ex;
}

#define SORT_FLAG_AGG (1u << 30)
#define SORT_FLAG_INC (2u << 30)
#define SORT_VAL_MASK ((1u << 30) - 1u)

// One LSD pass over all tiles (tiles by ticket: every lower tile belongs to a block that is already running).
template <bool HAS_VALS>
__device__ __forceinline__ void sort_pass(const u64 *__restrict__ kin, const u32 *__restrict__ vin, u64 *__restrict__ kout,
                                          u32 *__restrict__ vout, size_t n, int lo, int bits, const u32 *__restrict__ goff,
                                          u32 *status, u32 *ticket, u32 *err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_keys = reinterpret_cast<u64 *>(smem_raw);                           // SORT_TILE
    u32 *s_vals = reinterpret_cast<u32 *>(s_keys + SORT_TILE);                 // SORT_TILE
    u32 *s_whist = s_vals + SORT_TILE;                                         // 8 warps x 256
    u32 *s_dstart = s_whist + (SORT_NT / 32) * SORT_RADIX;                     // 256 tile-exclusive digit starts
    int *s_adj = reinterpret_cast<int *>(s_dstart + SORT_RADIX);               // 256 global base - tile start
    __shared__ u32 s_scan[34];
    __shared__ u32 s_tile;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 dmask = (1u << bits) - 1u;
    const u32 ntiles = (u32)((n + SORT_TILE - 1) / SORT_TILE);
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        for (int i = threadIdx.x; i < (SORT_NT / 32) * SORT_RADIX; i += SORT_NT) s_whist[i] = 0;
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const size_t tbase = (size_t)tile * SORT_TILE;
        const u32 tcount = (u32)min((size_t)SORT_TILE, n - tbase);
        // warp-striped load: in-tile order index of (wid, i, lane) is wid*512 + i*32 + lane
        u64 key[SORT_KPT];
        u32 off[SORT_KPT];
        u32 *wh = s_whist + wid * SORT_RADIX;
#pragma unroll
        for (int i = 0; i < SORT_KPT; ++i) {
            const u32 t = wid * (32 * SORT_KPT) + i * 32 + lane;
            key[i] = t < tcount ? kin[tbase + t] : ~0ull;
        }
#pragma unroll
        for (int i = 0; i < SORT_KPT; ++i) {
            const u32 t = wid * (32 * SORT_KPT) + i * 32 + lane;
            // out-of-range slots take digit 'dmask' but are ranked after every real key of the tile
            const u32 d = t < tcount ? (u32)((key[i] >> lo) & dmask) : (SORT_RADIX + 1);
            const u32 peers = __match_any_sync(0xFFFFFFFFu, d);
            const u32 lt = peers & ((1u << lane) - 1u);
            const int leader = __ffs(peers) - 1;
            u32 pre = 0;
            if (lane == leader && d <= dmask) { pre = wh[d]; wh[d] = pre + __popc(peers); }
            pre = __shfl_sync(0xFFFFFFFFu, pre, leader);
            off[i] = pre + __popc(lt);
            __syncwarp();
        }
        __syncthreads();
        // per digit: exclusive prefix over warps, tile count
        u32 dcount;
        {
            const int d = threadIdx.x;
            u32 run = 0;
#pragma unroll
            for (int w = 0; w < SORT_NT / 32; ++w) { const u32 c = s_whist[w * SORT_RADIX + d]; s_whist[w * SORT_RADIX + d] = run; run += c; }
            dcount = run;
        }
        // publish aggregate, scan digit starts inside the tile
        if (tile == 0) status[threadIdx.x] = SORT_FLAG_INC | dcount;
        else status[(size_t)tile * SORT_RADIX + threadIdx.x] = SORT_FLAG_AGG | dcount;
        const u32 dstart = block_excl_scan<SORT_NT>(dcount, s_scan);
        s_dstart[threadIdx.x] = dstart;
        // chained scan: each thread walks back for its own digit, SORT_LB predecessors per round trip (when a whole
        // generation of tiles starts together nobody has an inclusive prefix yet and tile t has to add up t aggregates:
        // one dependent load per predecessor made a pass over ~230 tiles cost ~25 us of pure latency)
        u32 excl = 0;
        if (tile > 0) {
            u32 spins = 0;
            int t = (int)tile - 1;
            bool done = false;
            while (!done && t >= 0) {
                u32 sv[SORT_LB];
#pragma unroll
                for (int k = 0; k < SORT_LB; ++k) {
                    const int idx = t - k;
                    sv[k] = SORT_FLAG_INC; // before tile 0: an inclusive prefix of zero
                    if (idx >= 0)
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(sv[k]) : "l"(status + (size_t)idx * SORT_RADIX + threadIdx.x) : "memory");
                }
                int used = 0;
#pragma unroll
                for (int k = 0; k < SORT_LB; ++k) {
                    if (done || used != k) continue;           // stop at the first predecessor that has not published yet
                    if ((sv[k] >> 30) == 0u) continue;
                    excl += sv[k] & SORT_VAL_MASK;
                    if (sv[k] & SORT_FLAG_INC) done = true;
                    used = k + 1;
                }
                t -= used;
                if (!done && used == 0 && ++spins > (1u << 24)) { *err = DERR_SCAN; break; }
            }
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(status + (size_t)tile * SORT_RADIX + threadIdx.x),
                         "r"(SORT_FLAG_INC | (excl + dcount)) : "memory");
        }
        s_adj[threadIdx.x] = (int)(goff[threadIdx.x] + excl) - (int)dstart;
        __syncthreads();
        // stable in-tile reorder through shared memory
#pragma unroll
        for (int i = 0; i < SORT_KPT; ++i) {
            const u32 t = wid * (32 * SORT_KPT) + i * 32 + lane;
            if (t < tcount) {
                const u32 d = (u32)((key[i] >> lo) & dmask);
                const u32 tp = s_dstart[d] + s_whist[wid * SORT_RADIX + d] + off[i];
                s_keys[tp] = key[i];
                if (HAS_VALS) s_vals[tp] = vin[tbase + t];
            }
        }
        __syncthreads();
        for (u32 t = threadIdx.x; t < tcount; t += SORT_NT) {
            const u64 k = s_keys[t];
            const u32 d = (u32)((k >> lo) & dmask);
            const size_t g = (size_t)((long long)s_adj[d] + (long long)t);
            kout[g] = k;
            if (HAS_VALS) vout[g] = s_vals[t];
        }
        __syncthreads();
    }
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(SORT_NT) k_sort_onesweep(const u64 *__restrict__ kin, const u32 *__restrict__ vin,
                                                           u64 *__restrict__ kout, u32 *__restrict__ vout, size_t n,
                                                           int lo, int bits, const u32 *__restrict__ goff,
                                                           u32 *status, u32 *ticket, u32 *err) {
    sort_pass<HAS_VALS>(kin, vin, kout, vout, n, lo, bits, goff, status, ticket, err);
}

// Device-driven form (duplicate elimination between rounds, gseg_dedup.cuh): size, bit range and the buffers of pass
// `pass` come from a descriptor in device memory that an earlier kernel of the same stream filled; a pass that is
// not needed exits at once, so the host can enqueue the worst case without knowing the graph.
struct SortDev {
    u32 active, n, npass, key_bits; // key_bits: total bits of the sort key
    u64 *keys[2];
    u32 *vals[2];
    u32 *hist, *status, *tickets;   // hist [SORT_MAXPASS][256]; status [npass][ntiles(n)][256]; tickets [SORT_MAXPASS + 1]
};
__global__ void __launch_bounds__(SORT_NT) k_sort_onesweep_dev(const SortDev *__restrict__ sd, int pass) {
    if (!sd->active || (u32)pass >= sd->npass) return;
    const size_t n = sd->n;
    const size_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    const int lo = 8 * pass, bits = min(8, (int)sd->key_bits - lo);
    sort_pass<true>(sd->keys[pass & 1], sd->vals[pass & 1], sd->keys[(pass & 1) ^ 1], sd->vals[(pass & 1) ^ 1], n, lo, bits,
                    sd->hist + pass * SORT_RADIX, sd->status + (size_t)pass * ntiles * SORT_RADIX, sd->tickets + pass,
                    sd->tickets + SORT_MAXPASS);
}
__global__ void __launch_bounds__(SORT_RADIX) k_sort_scan_dev(const SortDev *__restrict__ sd) {
    __shared__ u32 s[34];
    if (!sd->active || blockIdx.x >= sd->npass) return;
    u32 *h = sd->hist + blockIdx.x * SORT_RADIX;
    const u32 v = h[threadIdx.x];
    const u32 ex = block_excl_scan<SORT_RADIX>(v, s);
    h[threadIdx.x] = ex;
}

static size_t sort_smem_bytes() {
    return SORT_TILE * (sizeof(u64) + sizeof(u32)) + ((SORT_NT / 32) * SORT_RADIX + 2 * SORT_RADIX) * sizeof(u32);
}
static cudaError_t sort_set_attrs(SortScratch *s) {
    if (s->attr_set) return cudaSuccess;
    cudaError_t e;
    const int smem = (int)sort_smem_bytes();
    if ((e = cudaFuncSetAttribute(k_sort_onesweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_sort_onesweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_sort_onesweep_dev, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    s->attr_set = true;
    return cudaSuccess;
}
// Scratch for sorts of up to n keys in up to SORT_MAXPASS passes, allocated up front (gseg_create) so that neither the
// de-duplication between rounds nor the export of a strip's graph allocates.
static cudaError_t sort_scratch_reserve(SortScratch *s, size_t n) {
    cudaError_t e;
    const size_t need_status = (size_t)SORT_MAXPASS * ((n + SORT_TILE - 1) / SORT_TILE) * SORT_RADIX;
    if (s->cap_n < n) {
        cudaFree(s->keys_alt); cudaFree(s->vals_alt);
        s->keys_alt = nullptr; s->vals_alt = nullptr; s->cap_n = 0;
        if ((e = cudaMalloc((void **)&s->keys_alt, n * sizeof(u64))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&s->vals_alt, n * sizeof(u32))) != cudaSuccess) return e;
        s->cap_n = n;
    }
    if (s->cap_status < need_status) {
        cudaFree(s->status);
        s->status = nullptr; s->cap_status = 0;
        if ((e = cudaMalloc((void **)&s->status, need_status * sizeof(u32))) != cudaSuccess) return e;
        s->cap_status = need_status;
    }
    if (!s->hist) {
        if ((e = cudaMalloc((void **)&s->hist, SORT_MAXPASS * SORT_RADIX * sizeof(u32))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&s->tickets, (SORT_MAXPASS + 1) * sizeof(u32))) != cudaSuccess) return e;
    }
    return sort_set_attrs(s);
}

static cudaError_t onesweep_sort_pairs(SortScratch *s, u64 *keys, u32 *vals, size_t n, int begin_bit, int end_bit,
                                       cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (n >= (1ull << 30)) return cudaErrorInvalidValue;
    const int npass = (end_bit - begin_bit + 7) / 8;
    const size_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    const size_t need_status = (size_t)npass * ntiles * SORT_RADIX;
    cudaError_t e;
    if (s->cap_n < n) {
        cudaFree(s->keys_alt); cudaFree(s->vals_alt);
        s->keys_alt = nullptr; s->vals_alt = nullptr; s->cap_n = 0;
        if ((e = cudaMalloc((void **)&s->keys_alt, n * sizeof(u64))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&s->vals_alt, n * sizeof(u32))) != cudaSuccess) return e;
        s->cap_n = n;
    }
    if (s->cap_status < need_status) {
        cudaFree(s->status);
        s->status = nullptr; s->cap_status = 0;
        if ((e = cudaMalloc((void **)&s->status, need_status * sizeof(u32))) != cudaSuccess) return e;
        s->cap_status = need_status;
    }
    if (!s->hist) {
        if ((e = cudaMalloc((void **)&s->hist, SORT_MAXPASS * SORT_RADIX * sizeof(u32))) != cudaSuccess) return e;
        if ((e = cudaMalloc((void **)&s->tickets, (SORT_MAXPASS + 1) * sizeof(u32))) != cudaSuccess) return e;
    }
    const size_t smem = sort_smem_bytes();
    if ((e = sort_set_attrs(s)) != cudaSuccess) return e;
    cudaMemsetAsync(s->hist, 0, SORT_MAXPASS * SORT_RADIX * sizeof(u32), st);
    cudaMemsetAsync(s->tickets, 0, (SORT_MAXPASS + 1) * sizeof(u32), st);
    cudaMemsetAsync(s->status, 0, need_status * sizeof(u32), st);
    int gh = (int)((n + SORT_NT * 16 - 1) / (SORT_NT * 16));
    if (gh > 148 * 8) gh = 148 * 8;
    k_sort_hist<<<gh, SORT_NT, 0, st>>>(keys, n, begin_bit, end_bit, s->hist);
    k_sort_scan<<<npass, SORT_RADIX, 0, st>>>(s->hist);
    int gs = (int)(ntiles < 148 * 2 ? ntiles : 148 * 2);
    u64 *kin = keys, *kout = s->keys_alt;
    u32 *vin = vals, *vout = s->vals_alt;
    for (int p = 0; p < npass; ++p) {
        const int lo = begin_bit + 8 * p;
        const int bits = end_bit - lo < 8 ? end_bit - lo : 8;
        if (vals)
            k_sort_onesweep<true><<<gs, SORT_NT, smem, st>>>(kin, vin, kout, vout, n, lo, bits, s->hist + p * SORT_RADIX,
                                                             s->status + (size_t)p * ntiles * SORT_RADIX, s->tickets + p,
                                                             s->tickets + SORT_MAXPASS);
        else
            k_sort_onesweep<false><<<gs, SORT_NT, smem, st>>>(kin, nullptr, kout, nullptr, n, lo, bits, s->hist + p * SORT_RADIX,
                                                              s->status + (size_t)p * ntiles * SORT_RADIX, s->tickets + p,
                                                              s->tickets + SORT_MAXPASS);
        u64 *tk = kin; kin = kout; kout = tk;
        u32 *tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        cudaMemcpyAsync(keys, kin, n * sizeof(u64), cudaMemcpyDeviceToDevice, st);
        if (vals) cudaMemcpyAsync(vals, vin, n * sizeof(u32), cudaMemcpyDeviceToDevice, st);
    }
    return cudaGetLastError();
}
