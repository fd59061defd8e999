// gseg_kernels.cuh -- every CUDA kernel of the segmentation hot path (sm_100a).
//
// Stage map:
//   a1  k_blur_h / k_blur_v            separable Gaussian pre-filter          Report.pdf p3 s3.2 par.2
//   a2  k_sobel                        Sobel magnitude (superpixel variant)   Report.pdf p4 s3.2.4
//   a3  k_weights                      grid edge weights / strengths          Report.pdf p3 s3.2.1, p2 par.1
//   a4  k_r0_choose, k_edges           min outgoing edge per vertex/component Report.pdf p2-3 s3.1 steps 1-3
//   a6+a7 k_r0_succ, k_succ            predicate + 2-cycle removal            Report.pdf p3 steps 4-5
//   a8  k_jump, k_relabel              flatten + size / Int(C) / colour       Report.pdf p3 steps 7-8
//   a9  k_rootscan                     supervertex renumbering (flag + scan)  Report.pdf p3 s3.2.2
//   a10 k_r0_edges, k_edges            edge relabel, self-loop drop, stable compaction
//   a11 k_edges<SUPERPIX>              per-round re-weighting from component means
//   a12 k_compose, k_compose_all       hierarchy materialisation              Report.pdf p4 s3.2.3
//   a13 min-size rounds                phase PH_MINSIZE of k_succ             Report.pdf p3 step 6
//   a14 k_colorize                     random colour per component            Report.pdf p4 s3.2.3
//
// Float contract: every fp32 product, sum, quotient and square root on the weight path is a
// separately rounded IEEE operation (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn) in the
// order DESIGN.md "Semantics" states; ptxas never contracts these intrinsics into FMAs, so weights
// are bit-identical to the CPU oracle and the total edge order (weight bits, edge index) is too.
#pragma once
#include "gseg_device.cuh"

#define NT 256
#define TILE_E (NT * 4) // edges per tile of the edge compaction
#define TILE_C (NT * 8) // components per tile of the root scan

__device__ __constant__ int c_DX[4] = {1, 0, 1, 1};
__device__ __constant__ int c_DY[4] = {0, 1, 1, -1};

// ------------------------------------------------------------------------------------------------
// run set-up
// ------------------------------------------------------------------------------------------------
__global__ void k_init(GsegCtl *ctl) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const u32 V = (u32)ctl->p.w * (u32)ctl->p.h;
        ctl->Vcur = V;
        ctl->Ecur = 0;
        ctl->Vnext = V;
        ctl->Enext = 0;
        ctl->phase = PH_PRED;
        ctl->round = 0;
        ctl->levels = 0;
        ctl->error = DERR_NONE;
        ctl->ticketC = 0;
        ctl->ticketE = 0;
        ctl->map_off[0] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// a1: separable Gaussian, clamped borders.  u8 interleaved RGB -> 3 fp32 planes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_blur_h(const GsegCtl *__restrict__ ctl, float *__restrict__ tmp) {
    const int w = ctl->p.w, h = ctl->p.h, stride = ctl->p.stride, len = ctl->p.mask_len;
    const uint8_t *__restrict__ rgb = ctl->p.rgb;
    const float *m = ctl->p.mask;
    const u32 V = (u32)w * (u32)h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V; p += gridDim.x * NT) {
        const int y = p / w, x = p - y * w;
        const uint8_t *row = rgb + (size_t)y * stride;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = __fmul_rn(m[0], (float)row[3 * x + c]);
            for (int i = 1; i < len; ++i) {
                const int xl = max(x - i, 0), xr = min(x + i, w - 1);
                const float pair = __fadd_rn((float)row[3 * xl + c], (float)row[3 * xr + c]);
                s = __fadd_rn(s, __fmul_rn(m[i], pair));
            }
            tmp[(size_t)c * V + p] = s;
        }
    }
}

__global__ void __launch_bounds__(NT) k_blur_v(const GsegCtl *__restrict__ ctl, const float *__restrict__ tmp,
                                               float *__restrict__ planes) {
    const int w = ctl->p.w, h = ctl->p.h, len = ctl->p.mask_len;
    const float *m = ctl->p.mask;
    const u32 V = (u32)w * (u32)h;
    for (u32 t = blockIdx.x * NT + threadIdx.x; t < 3u * V; t += gridDim.x * NT) {
        const u32 c = t / V, p = t - c * V;
        const int y = p / w, x = p - y * w;
        const float *pl = tmp + (size_t)c * V;
        float s = __fmul_rn(m[0], pl[p]);
        for (int i = 1; i < len; ++i) {
            const int yu = max(y - i, 0), yd = min(y + i, h - 1);
            const float pair = __fadd_rn(pl[(size_t)yu * w + x], pl[(size_t)yd * w + x]);
            s = __fadd_rn(s, __fmul_rn(m[i], pair));
        }
        planes[t] = s;
    }
}

// a2: Sobel magnitude of the blurred intensity (superpixel variant).
__device__ __forceinline__ float intensity(const float *pl, u32 V, u32 p) {
    return __fmul_rn(__fadd_rn(__fadd_rn(pl[p], pl[V + p]), pl[2 * V + p]), 0.33333334f);
}
__global__ void __launch_bounds__(NT) k_sobel(const GsegCtl *__restrict__ ctl, const float *__restrict__ planes,
                                              float *__restrict__ G) {
    const int w = ctl->p.w, h = ctl->p.h;
    const u32 V = (u32)w * (u32)h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V; p += gridDim.x * NT) {
        const int y = p / w, x = p - y * w;
        const int xm = max(x - 1, 0), xp = min(x + 1, w - 1), ym = max(y - 1, 0), yp = min(y + 1, h - 1);
#define AT(xx, yy) intensity(planes, V, (u32)(yy) * w + (xx))
        const float a00 = AT(xm, ym), a10 = AT(x, ym), a20 = AT(xp, ym);
        const float a01 = AT(xm, y), a21 = AT(xp, y);
        const float a02 = AT(xm, yp), a12 = AT(x, yp), a22 = AT(xp, yp);
#undef AT
        const float r = __fadd_rn(__fadd_rn(a20, __fmul_rn(2.0f, a21)), a22);
        const float l = __fadd_rn(__fadd_rn(a00, __fmul_rn(2.0f, a01)), a02);
        const float d = __fadd_rn(__fadd_rn(a02, __fmul_rn(2.0f, a12)), a22);
        const float u = __fadd_rn(__fadd_rn(a00, __fmul_rn(2.0f, a10)), a20);
        const float gx = __fsub_rn(r, l), gy = __fsub_rn(d, u);
        G[p] = __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
    }
}

// a3: grid edge weights, plane-major wgrid[d*V + p]; +inf where the edge does not exist.
__device__ __forceinline__ float l2rgb(const float *pl, u32 V, u32 p, u32 q) {
    const float dr = __fsub_rn(pl[p], pl[q]), dg = __fsub_rn(pl[V + p], pl[V + q]),
                db = __fsub_rn(pl[2 * V + p], pl[2 * V + q]);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(dg, dg)), __fmul_rn(db, db));
    return __fsqrt_rn(s);
}
template <bool STRENGTH>
__global__ void __launch_bounds__(NT) k_weights(const GsegCtl *__restrict__ ctl, const float *__restrict__ planes,
                                                const float *__restrict__ G, float *__restrict__ wgrid) {
    const int w = ctl->p.w, h = ctl->p.h, D = ctl->p.D;
    const u32 V = (u32)w * (u32)h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V; p += gridDim.x * NT) {
        const int y = p / w, x = p - y * w;
        for (int d = 0; d < D; ++d) {
            const int xx = x + c_DX[d], yy = y + c_DY[d];
            float v = __int_as_float(GSEG_INF_BITS);
            if (xx < w && yy < h && yy >= 0) {
                const u32 q = (u32)yy * w + xx;
                v = STRENGTH ? __fmul_rn(0.5f, __fadd_rn(G[p], G[q])) : l2rgb(planes, V, p, q);
            }
            wgrid[(size_t)d * V + p] = v;
        }
    }
}

// 24.8 fixed-point colour of a pixel and the superpixel round weight.
__device__ __forceinline__ int fx8(float v) { return __float2int_rn(__fmul_rn(v, 256.0f)); }
__device__ __forceinline__ float mean_dist(const long long *ca, u32 sa, const long long *cb, u32 sb) {
    const float fa = __fmul_rn(__uint2float_rn(sa), 256.0f), fb = __fmul_rn(__uint2float_rn(sb), 256.0f);
    const float dr = __fsub_rn(__fdiv_rn(__ll2float_rn(ca[0]), fa), __fdiv_rn(__ll2float_rn(cb[0]), fb));
    const float dg = __fsub_rn(__fdiv_rn(__ll2float_rn(ca[1]), fa), __fdiv_rn(__ll2float_rn(cb[1]), fb));
    const float db = __fsub_rn(__fdiv_rn(__ll2float_rn(ca[2]), fa), __fdiv_rn(__ll2float_rn(cb[2]), fb));
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(dg, dg)), __fmul_rn(db, db));
    return __fsqrt_rn(s);
}

// ------------------------------------------------------------------------------------------------
// Round 0 on the implicit grid: every pixel is its own component, so the minimum outgoing edge is
// a pure stencil over its <= 2D incident edges (no atomics).  dir codes: 0..3 own edge (E,S,SE,NE),
// 4..7 the reverse (W,N,NW,SW); 255 = none / rejected by the predicate.
// ------------------------------------------------------------------------------------------------
template <int VARIANT>
__global__ void __launch_bounds__(NT) k_r0_choose(const GsegCtl *__restrict__ ctl, const float *__restrict__ wgrid,
                                                  const float *__restrict__ planes, uint8_t *__restrict__ dir0,
                                                  u32 *__restrict__ wsel) {
    const int w = ctl->p.w, h = ctl->p.h, D = ctl->p.D;
    const u32 V = (u32)w * (u32)h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V; p += gridDim.x * NT) {
        const int y = p / w, x = p - y * w;
        u64 best = GSEG_KEY_NONE;
        int bdir = 255;
        long long cp[3];
        if (VARIANT == GSEG_SUPERPIX) {
            cp[0] = fx8(planes[p]); cp[1] = fx8(planes[V + p]); cp[2] = fx8(planes[2 * V + p]);
        }
        for (int d = 0; d < 2 * D; ++d) {
            const int dd = d < D ? d : d - D;
            const int sgn = d < D ? 1 : -1;
            const int xx = x + sgn * c_DX[dd], yy = y + sgn * c_DY[dd];
            if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
            const u32 q = (u32)yy * w + xx;
            const u32 owner = d < D ? p : q; // the pixel whose edge list holds this edge
            float wv = wgrid[(size_t)dd * V + owner];
            if (VARIANT == GSEG_SUPERPIX) {
                long long cq[3] = {fx8(planes[q]), fx8(planes[V + q]), fx8(planes[2 * V + q])};
                wv = __fmul_rn(wv, mean_dist(cp, 1u, cq, 1u));
            }
            const u64 key = make_key(__float_as_uint(wv), owner * (u32)D + (u32)dd);
            if (key < best) { best = key; bdir = d < D ? dd : dd + 4; }
        }
        u32 wb = 0;
        if (bdir != 255) {
            wb = (u32)(best >> 32);
            if (VARIANT == GSEG_FELZ) {
                // Int = 0, |C| = 1 on both sides: thr = 0 + k/1
                const float thr = __fadd_rn(0.0f, __fdiv_rn(ctl->p.k, 1.0f));
                if (!(__uint_as_float(wb) <= thr)) bdir = 255;
            }
        }
        dir0[p] = (uint8_t)bdir;
        wsel[p] = wb;
    }
}

__device__ __forceinline__ u32 dir_neighbor(u32 p, int dir, int w) {
    const int dd = dir & 3, sgn = dir < 4 ? 1 : -1;
    return (u32)((int)p + sgn * (c_DY[dd] * w + c_DX[dd]));
}

// successor with 2-cycle removal; also clears the next round's accumulators.
template <bool SUPERPIX>
__global__ void __launch_bounds__(NT) k_r0_succ(const GsegCtl *__restrict__ ctl, const uint8_t *__restrict__ dir0,
                                                u32 *__restrict__ succ, u32 *__restrict__ size_n,
                                                u32 *__restrict__ int_n, u64 *__restrict__ best_n,
                                                long long *__restrict__ csum_n) {
    const int w = ctl->p.w;
    const u32 V = (u32)w * (u32)ctl->p.h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V; p += gridDim.x * NT) {
        const int d = dir0[p];
        u32 s = p;
        if (d != 255) {
            const u32 q = dir_neighbor(p, d, w);
            s = (dir0[q] == (d ^ 4) && p < q) ? p : q;
        }
        succ[p] = s;
        size_n[p] = 0;
        int_n[p] = 0;
        best_n[p] = GSEG_KEY_NONE;
        if (SUPERPIX) { csum_n[3 * (size_t)p] = 0; csum_n[3 * (size_t)p + 1] = 0; csum_n[3 * (size_t)p + 2] = 0; }
    }
}

// ------------------------------------------------------------------------------------------------
// a8: flatten the merge forest in place.  Every value ever stored in succ[] is an ancestor of its
// slot and roots never change, so concurrent chasing with in-place compression is race-benign.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_jump(const GsegCtl *__restrict__ ctl, u32 *succ) {
    if (ctl->phase == PH_DONE) return;
    const u32 V = ctl->Vcur;
    for (u32 c = blockIdx.x * NT + threadIdx.x; c < V; c += gridDim.x * NT) {
        u32 s = ld_relaxed_u32(succ + c);
        if (s == c) continue;
        for (;;) {
            const u32 ss = ld_relaxed_u32(succ + s);
            if (ss == s) break;
            s = ss;
        }
        succ[c] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// a9: roots -> dense new ids in root-index order (flag + single-pass look-back scan).
// rank[c] = number of roots below c.  Total -> ctl->Vnext.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_rootscan(GsegCtl *ctl, const u32 *__restrict__ succ, u32 *__restrict__ rank,
                                                 u64 *status) {
    if (ctl->phase == PH_DONE) return;
    __shared__ u32 s_scan[34];
    __shared__ u32 s_tile;
    const u32 V = ctl->Vcur;
    const u32 ntiles = (V + TILE_C - 1) / TILE_C;
    const u32 tag = ctl->p.epoch_base + ctl->round * 2u + 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ticketE = 0;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->ticketC, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0 && threadIdx.x == 0) ctl->Vnext = 0;
            break;
        }
        const u32 base = tile * TILE_C + threadIdx.x * 8;
        u32 f[8], cnt = 0;
        if (base + 7 < V) {
            const uint4 a = *reinterpret_cast<const uint4 *>(succ + base);
            const uint4 b = *reinterpret_cast<const uint4 *>(succ + base + 4);
            f[0] = a.x == base; f[1] = a.y == base + 1; f[2] = a.z == base + 2; f[3] = a.w == base + 3;
            f[4] = b.x == base + 4; f[5] = b.y == base + 5; f[6] = b.z == base + 6; f[7] = b.w == base + 7;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (base + j < V) ? (succ[base + j] == base + j) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt += f[j];
        u32 off = tile_offset<NT>(cnt, tile, tag, status, &ctl->error, s_scan);
        if (tile == ntiles - 1 && threadIdx.x == 0) ctl->Vnext = s_scan[33] + s_scan[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + j < V) rank[base + j] = off;
            off += f[j];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// a8: old id -> new id map, and accumulation of size / Int(C) / colour sums into the new ids.
// R0: components are pixels (size 1, Int 0, colour = the pixel's fixed-point colour).
// ------------------------------------------------------------------------------------------------
template <bool R0, bool SUPERPIX>
__global__ void __launch_bounds__(NT) k_relabel(const GsegCtl *__restrict__ ctl, const u32 *__restrict__ succ,
                                                const u32 *__restrict__ rank, const u32 *__restrict__ wsel,
                                                const u32 *__restrict__ size_c, const u32 *__restrict__ int_c,
                                                const long long *__restrict__ csum_c, const float *__restrict__ planes,
                                                u32 *__restrict__ arena, u32 *__restrict__ size_n,
                                                u32 *__restrict__ int_n, long long *__restrict__ csum_n) {
    if (ctl->phase == PH_DONE) return;
    const u32 V = ctl->Vcur;
    u32 *map = arena + ctl->map_off[ctl->round];
    const u32 V0 = (u32)ctl->p.w * (u32)ctl->p.h;
    for (u32 c = blockIdx.x * NT + threadIdx.x; c < V; c += gridDim.x * NT) {
        const u32 r = succ[c];
        const u32 m = rank[r];
        map[c] = m;
        atomicAdd(size_n + m, R0 ? 1u : size_c[c]);
        u32 iv = R0 ? 0u : int_c[c];
        if (r != c) iv = max(iv, wsel[c]);
        if (iv) atomicMax(int_n + m, iv);
        if (SUPERPIX) {
            long long v0, v1, v2;
            if (R0) { v0 = fx8(planes[c]); v1 = fx8(planes[V0 + c]); v2 = fx8(planes[2 * V0 + c]); }
            else { v0 = csum_c[3 * (size_t)c]; v1 = csum_c[3 * (size_t)c + 1]; v2 = csum_c[3 * (size_t)c + 2]; }
            atomicAdd((u64 *)(csum_n + 3 * (size_t)m), (u64)v0);
            atomicAdd((u64 *)(csum_n + 3 * (size_t)m + 1), (u64)v1);
            atomicAdd((u64 *)(csum_n + 3 * (size_t)m + 2), (u64)v2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// a10 (round 0): grid edges -> explicit list of inter-component edges, in edge-index order (stable),
// fused with next round's per-component minimum (key = weight bits << 32 | position in the list;
// stable compaction keeps list order == edge-index order, so position is the same tie-break).
// One thread per pixel; a tile is NT pixels.
// ------------------------------------------------------------------------------------------------
template <bool SUPERPIX>
__global__ void __launch_bounds__(NT) k_r0_edges(GsegCtl *ctl, const float *__restrict__ wgrid,
                                                 const u32 *__restrict__ arena, u32 *__restrict__ oa,
                                                 u32 *__restrict__ ob, u32 *__restrict__ ow, u64 *best_n,
                                                 const u32 *__restrict__ size_n, const long long *__restrict__ csum_n,
                                                 u64 *status) {
    if (ctl->phase == PH_DONE) return;
    __shared__ u32 s_scan[34];
    __shared__ u32 s_tile;
    const int w = ctl->p.w, h = ctl->p.h, D = ctl->p.D;
    const u32 V = (u32)w * (u32)h;
    const u32 *__restrict__ map = arena; // map_off[0] == 0
    const u32 ntiles = (V + NT - 1) / NT;
    const u32 tag = ctl->p.epoch_base + ctl->round * 2u + 2u;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ticketC = 0;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->ticketE, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0 && threadIdx.x == 0) ctl->Enext = 0;
            break;
        }
        const u32 p = tile * NT + threadIdx.x;
        u32 a = 0, b[4], wv[4], cnt = 0;
        bool keep[4] = {false, false, false, false};
        if (p < V) {
            const int y = p / w, x = p - y * w;
            a = map[p];
            for (int d = 0; d < D; ++d) {
                const int xx = x + c_DX[d], yy = y + c_DY[d];
                if (xx < w && yy < h && yy >= 0) {
                    b[d] = map[(u32)yy * w + xx];
                    if (b[d] != a) {
                        keep[d] = true;
                        wv[d] = __float_as_uint(wgrid[(size_t)d * V + p]);
                        ++cnt;
                    }
                }
            }
        }
        u32 pos = tile_offset<NT>(cnt, tile, tag, status, &ctl->error, s_scan);
        if (tile == ntiles - 1 && threadIdx.x == 0) ctl->Enext = s_scan[33] + s_scan[32];
        for (int d = 0; d < D; ++d) {
            if (!keep[d]) continue;
            oa[pos] = a; ob[pos] = b[d]; ow[pos] = wv[d];
            u32 kb = wv[d];
            if (SUPERPIX)
                kb = __float_as_uint(__fmul_rn(__uint_as_float(wv[d]),
                                               mean_dist(csum_n + 3 * (size_t)a, size_n[a], csum_n + 3 * (size_t)b[d], size_n[b[d]])));
            const u64 key = make_key(kb, pos);
            atomicMin(best_n + a, key);
            atomicMin(best_n + b[d], key);
            ++pos;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// a6+a7 (rounds >= 1): each component's choice under the predicate / min-size rule, 2-cycle removal.
// ------------------------------------------------------------------------------------------------
struct RoundView {
    const u64 *best;
    const u32 *ea, *eb, *size, *Int;
    float k;
    int min_size, variant;
    u32 phase;
};
__device__ __forceinline__ u32 comp_choice(const RoundView &v, u32 c, u32 *wbits_out) {
    const u64 key = v.best[c];
    if (key == GSEG_KEY_NONE) return c;
    const u32 pos = (u32)key, wb = (u32)(key >> 32);
    const u32 a = v.ea[pos], b = v.eb[pos];
    const u32 other = a == c ? b : a;
    bool ok;
    if (v.variant != GSEG_FELZ) ok = true;
    else if (v.phase == PH_PRED) {
        const float wt = __uint_as_float(wb);
        const float ta = __fadd_rn(__uint_as_float(v.Int[a]), __fdiv_rn(v.k, __uint2float_rn(v.size[a])));
        const float tb = __fadd_rn(__uint_as_float(v.Int[b]), __fdiv_rn(v.k, __uint2float_rn(v.size[b])));
        ok = wt <= ta && wt <= tb;
    } else ok = v.size[c] < (u32)v.min_size;
    *wbits_out = wb;
    return ok ? other : c;
}

template <bool SUPERPIX>
__global__ void __launch_bounds__(NT) k_succ(const GsegCtl *__restrict__ ctl, const u64 *__restrict__ best,
                                             const u32 *__restrict__ ea, const u32 *__restrict__ eb,
                                             const u32 *__restrict__ size_c, const u32 *__restrict__ int_c,
                                             u32 *__restrict__ succ, u32 *__restrict__ wsel, u32 *__restrict__ size_n,
                                             u32 *__restrict__ int_n, u64 *__restrict__ best_n,
                                             long long *__restrict__ csum_n) {
    if (ctl->phase == PH_DONE) return;
    const u32 V = ctl->Vcur;
    RoundView v;
    v.best = best; v.ea = ea; v.eb = eb; v.size = size_c; v.Int = int_c;
    v.k = ctl->p.k; v.min_size = ctl->p.min_size; v.variant = ctl->p.variant; v.phase = ctl->phase;
    for (u32 c = blockIdx.x * NT + threadIdx.x; c < V; c += gridDim.x * NT) {
        u32 wb = 0, wb2;
        u32 s = comp_choice(v, c, &wb);
        if (s != c) {
            const u32 t = comp_choice(v, s, &wb2);
            if (t == c && c < s) s = c;
        }
        succ[c] = s;
        wsel[c] = wb;
        size_n[c] = 0;
        int_n[c] = 0;
        best_n[c] = GSEG_KEY_NONE;
        if (SUPERPIX) { csum_n[3 * (size_t)c] = 0; csum_n[3 * (size_t)c + 1] = 0; csum_n[3 * (size_t)c + 2] = 0; }
    }
}

// ------------------------------------------------------------------------------------------------
// a10 (rounds >= 1): relabel edge ends through this round's map, drop self-loops, stable compaction,
// fused with next round's per-component minimum.  4 consecutive edges per thread.
// ------------------------------------------------------------------------------------------------
template <bool SUPERPIX>
__global__ void __launch_bounds__(NT) k_edges(GsegCtl *ctl, const u32 *__restrict__ ea, const u32 *__restrict__ eb,
                                              const u32 *__restrict__ ew, const u32 *__restrict__ arena,
                                              u32 *__restrict__ oa, u32 *__restrict__ ob, u32 *__restrict__ ow,
                                              u64 *best_n, const u32 *__restrict__ size_n,
                                              const long long *__restrict__ csum_n, u64 *status) {
    if (ctl->phase == PH_DONE) return;
    __shared__ u32 s_scan[34];
    __shared__ u32 s_tile;
    const u32 E = ctl->Ecur;
    const u32 *__restrict__ map = arena + ctl->map_off[ctl->round];
    const u32 ntiles = (E + TILE_E - 1) / TILE_E;
    const u32 tag = ctl->p.epoch_base + ctl->round * 2u + 2u;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->ticketC = 0;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->ticketE, 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) {
            if (tile == 0 && threadIdx.x == 0) ctl->Enext = 0;
            break;
        }
        const u32 base = tile * TILE_E + threadIdx.x * 4;
        u32 a[4], b[4], wv[4], cnt = 0;
        bool keep[4];
        if (base + 3 < E) {
            const uint4 va = *reinterpret_cast<const uint4 *>(ea + base);
            const uint4 vb = *reinterpret_cast<const uint4 *>(eb + base);
            const uint4 vw = *reinterpret_cast<const uint4 *>(ew + base);
            a[0] = va.x; a[1] = va.y; a[2] = va.z; a[3] = va.w;
            b[0] = vb.x; b[1] = vb.y; b[2] = vb.z; b[3] = vb.w;
            wv[0] = vw.x; wv[1] = vw.y; wv[2] = vw.z; wv[3] = vw.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = map[a[j]]; b[j] = map[b[j]]; keep[j] = a[j] != b[j]; cnt += keep[j]; }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                keep[j] = false;
                if (base + j < E) {
                    a[j] = map[ea[base + j]]; b[j] = map[eb[base + j]]; wv[j] = ew[base + j];
                    keep[j] = a[j] != b[j];
                    cnt += keep[j];
                }
            }
        }
        u32 pos = tile_offset<NT>(cnt, tile, tag, status, &ctl->error, s_scan);
        if (tile == ntiles - 1 && threadIdx.x == 0) ctl->Enext = s_scan[33] + s_scan[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!keep[j]) continue;
            oa[pos] = a[j]; ob[pos] = b[j]; ow[pos] = wv[j];
            u32 kb = wv[j];
            if (SUPERPIX)
                kb = __float_as_uint(__fmul_rn(__uint_as_float(wv[j]),
                                               mean_dist(csum_n + 3 * (size_t)a[j], size_n[a[j]], csum_n + 3 * (size_t)b[j], size_n[b[j]])));
            const u64 key = make_key(kb, pos);
            atomicMin(best_n + a[j], key);
            atomicMin(best_n + b[j], key);
            ++pos;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// End-of-round bookkeeping: statistics, phase machine, arena accounting.  One thread.
// ------------------------------------------------------------------------------------------------
__global__ void k_advance(GsegCtl *ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ctl->phase == PH_DONE) return;
    const u32 r = ctl->round;
    const u32 V = ctl->Vcur, Vn = ctl->Vnext, merged = V - Vn;
    ctl->stV[r] = V; ctl->stE[r] = ctl->Ecur; ctl->stM[r] = merged; ctl->stP[r] = ctl->phase; ctl->stVafter[r] = Vn;
    const int variant = ctl->p.variant;
    u32 phase = ctl->phase, levels = ctl->levels;
    if (merged == 0) {
        if (variant == GSEG_FELZ && phase == PH_PRED && ctl->p.min_size > 1) phase = PH_MINSIZE;
        else phase = PH_DONE;
    } else {
        ++levels;
        if (variant != GSEG_FELZ && (Vn <= 1u || (int)levels >= ctl->p.max_levels)) phase = PH_DONE;
    }
    const u32 next_off = ctl->map_off[r] + V;
    ctl->map_off[r + 1] = next_off;
    ctl->round = r + 1;
    ctl->levels = levels;
    ctl->Vcur = Vn;
    ctl->Ecur = ctl->Enext;
    if ((int)(r + 1) >= ctl->p.max_rounds) phase = PH_DONE;
    if (phase != PH_DONE && (u64)next_off + (u64)Vn > (u64)ctl->p.arena_cap) { ctl->error = DERR_ARENA; phase = PH_DONE; }
    ctl->phase = phase;
}

// ------------------------------------------------------------------------------------------------
// a12: hierarchy materialisation.  Level l = composition of the maps of rounds 0..l.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_compose(const GsegCtl *__restrict__ ctl, const u32 *__restrict__ arena,
                                                int last_round, int *__restrict__ out) {
    const u32 V0 = (u32)ctl->p.w * (u32)ctl->p.h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V0; p += gridDim.x * NT) {
        u32 l = arena[p];
        for (int r = 1; r <= last_round; ++r) l = arena[ctl->map_off[r] + l];
        out[p] = (int)l;
    }
}
// one more level from the previous one (all-levels output: V reads + V writes per level)
__global__ void __launch_bounds__(NT) k_compose_step(const GsegCtl *__restrict__ ctl, const u32 *__restrict__ arena,
                                                     int round, const int *__restrict__ prev, int *__restrict__ out) {
    const u32 V0 = (u32)ctl->p.w * (u32)ctl->p.h;
    const u32 *map = arena + ctl->map_off[round];
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V0; p += gridDim.x * NT) out[p] = (int)map[prev[p]];
}

__device__ __forceinline__ u64 d_sm64(u64 x) {
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u64 d_hash2(u64 seed, u64 a, u64 b) {
    return d_sm64(d_sm64(seed ^ (a * 0xD6E8FEB86659FD93ull)) + b);
}

// a14: random colour per component id (counter-based hash instead of cuRAND state).
__global__ void __launch_bounds__(NT) k_colorize(const int *__restrict__ labels, u32 V0, u64 seed,
                                                 uint8_t *__restrict__ out) {
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V0; p += gridDim.x * NT) {
        const u64 hsh = d_hash2(seed, (u64)(u32)labels[p], 3);
        out[3 * (size_t)p] = (uint8_t)(hsh & 255);
        out[3 * (size_t)p + 1] = (uint8_t)((hsh >> 8) & 255);
        out[3 * (size_t)p + 2] = (uint8_t)((hsh >> 16) & 255);
    }
}

// plane-major grid weights -> edge-index order (gseg_weights)
__global__ void __launch_bounds__(NT) k_weights_export(const float *__restrict__ wgrid, u32 V, int D,
                                                       float *__restrict__ out) {
    for (u32 t = blockIdx.x * NT + threadIdx.x; t < V * (u32)D; t += gridDim.x * NT) {
        const u32 p = t / D, d = t - p * D;
        out[t] = wgrid[(size_t)d * V + p];
    }
}

// Synthetic input generator; integer arithmetic only.
__global__ void __launch_bounds__(NT) k_synth(uint8_t *__restrict__ rgb, int w, int h, u64 seed) {
    const u32 V = (u32)w * (u32)h;
    for (u32 p = blockIdx.x * NT + threadIdx.x; p < V; p += gridDim.x * NT) {
        const int y = p / w, x = p - y * w;
        const int cx = x >> 6, cy = y >> 6;
        long long bestd = 0x7FFFFFFFFFFFFFFFll;
        u64 besth = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int ccx = cx + dx, ccy = cy + dy;
                const u64 cell = ((u64)(ccy + 1) << 20) | (u64)(ccx + 1);
                const u64 hs = d_hash2(seed, cell, 1);
                const long long sx = (long long)ccx * 64 + (long long)(hs & 63);
                const long long sy = (long long)ccy * 64 + (long long)((hs >> 6) & 63);
                const long long d = (x - sx) * (x - sx) + (y - sy) * (y - sy);
                if (d < bestd) { bestd = d; besth = hs; }
            }
        const u64 hn = d_hash2(seed, (u64)y * (u64)w + (u64)x, 2);
        for (int c = 0; c < 3; ++c) {
            const int base = (int)((besth >> (16 + 8 * c)) & 255);
            const int n = (int)(((hn >> (16 * c)) & 0xFFFF) % 17) - 8;
            rgb[3 * (size_t)p + c] = (uint8_t)min(max(base + n, 0), 255);
        }
    }
}
