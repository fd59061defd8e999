/*
 * gseg_oracle.c -- CPU ORACLE for the graph-segmentation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (libgseg.so) never links, loads or calls it.
 *
 * PARITY STATUS: *** parity unpinned ***.  The mounted reference (/root/reference) holds no source
 * code, no tests and no golden vectors: only README.md, installation.md and
 * Report.pdf.  This file is therefore a restatement of
 *   - Report.pdf p1-2 section 2.1  (Gaussian pre-filter, L2 RGB edge weights, sorted-edge merge with
 *                                   the adaptive criterion and parameter k),
 *   - Report.pdf p2 section 2.2    (hierarchies: one level per Boruvka round, supervertices, lightest
 *                                   duplicate edge, no merge predicate),
 *   - Report.pdf p2-3 section 3.1 + p9 Appendix A Alg.1-6 (Boruvka-with-predicate round structure:
 *                                   min edge per vertex / per component, remove 2-cycles, mark by
 *                                   predicate, update parents, flatten + size/Int update),
 *   - Report.pdf p4 section 3.2.3  (hierarchy reconstruction from per-round supervertex ids),
 *   - Report.pdf p4 section 3.2.4  (superpixel hierarchy: Sobel edge strength x mean-colour distance,
 *                                   re-evaluated every round),
 * plus the published algorithm the report names as its CPU baseline -- Felzenszwalb & Huttenlocher,
 * "Efficient graph-based image segmentation", IJCV 2004, and its public `segment` program
 * (Report.pdf ref [23]; third-party, not vendored in the reference, no pinned version) -- whose
 * smoothing / edge construction / criterion are restated from the paper and from memory of that
 * program.  Where the report leaves a choice open the choice is stated in DESIGN.md "Semantics" and
 * tagged [D] below.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; contraction MUST stay off: the float
 * arithmetic order below is the contract the CUDA path reproduces bit for bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXMASK 64
#define KEY_NONE 0xFFFFFFFFFFFFFFFFull

/* ------------------------------------------------------------------------------------------------
 * Synthetic input (deterministic piecewise-constant regions + noise; integer
 * arithmetic only so the CUDA generator, this one and the numpy one agree bit for bit).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t hash2(uint64_t seed, uint64_t a, uint64_t b) {
    return sm64(sm64(seed ^ (a * 0xD6E8FEB86659FD93ull)) + b);
}

/* One jittered site per 64x64 cell; a pixel takes the colour of the nearest site among the 3x3
 * surrounding cells (ties: scan order), then uniform integer noise in [-8, 8] per channel. */
void orc_synth(uint8_t *rgb, int w, int h, uint64_t seed) {
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            int cx = x >> 6, cy = y >> 6;
            int64_t bestd = INT64_MAX;
            uint64_t besth = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int ccx = cx + dx, ccy = cy + dy;
                    uint64_t cell = ((uint64_t)(ccy + 1) << 20) | (uint64_t)(ccx + 1);
                    uint64_t hs = hash2(seed, cell, 1);
                    int64_t sx = (int64_t)ccx * 64 + (int64_t)(hs & 63);
                    int64_t sy = (int64_t)ccy * 64 + (int64_t)((hs >> 6) & 63);
                    int64_t d = (x - sx) * (x - sx) + (y - sy) * (y - sy);
                    if (d < bestd) { bestd = d; besth = hs; }
                }
            uint64_t hn = hash2(seed, (uint64_t)y * (uint64_t)w + (uint64_t)x, 2);
            for (int c = 0; c < 3; ++c) {
                int base = (int)((besth >> (16 + 8 * c)) & 255);
                int n = (int)(((hn >> (16 * c)) & 0xFFFF) % 17) - 8;
                int v = base + n;
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                rgb[((size_t)y * w + x) * 3 + c] = (uint8_t)v;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Gaussian pre-filter (Report.pdf p2 section 2.1 "a Gaussian filter is also applied"; tap layout
 * after F&H `segment`: half-width ceil(4 sigma), normalised one-sided mask, clamped borders,
 * horizontal then vertical).  Every product and sum below is a separately rounded fp32 op.
 * ---------------------------------------------------------------------------------------------- */
int orc_gauss_mask(float sigma, float *mask) {
    if (sigma < 0.01f) sigma = 0.01f;
    int len = (int)ceilf(sigma * 4.0f) + 1;
    if (len > ORC_MAXMASK) return -1;
    double m[ORC_MAXMASK], s = 0.0;
    for (int i = 0; i < len; ++i) {
        double t = (double)i / (double)sigma;
        m[i] = exp(-0.5 * t * t);
    }
    for (int i = 1; i < len; ++i) s += m[i];
    s = 2.0 * s + m[0];
    for (int i = 0; i < len; ++i) mask[i] = (float)(m[i] / s);
    return len;
}

void orc_blur(const uint8_t *rgb, int w, int h, float sigma, float *planes) {
    float mask[ORC_MAXMASK];
    int len = orc_gauss_mask(sigma, mask);
    size_t V = (size_t)w * h;
    float *tmp = (float *)malloc(V * sizeof(float));
    for (int c = 0; c < 3; ++c) {
        float *out = planes + (size_t)c * V;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const uint8_t *row = rgb + (size_t)y * w * 3 + c;
                float s = mask[0] * (float)row[(size_t)x * 3];
                for (int i = 1; i < len; ++i) {
                    int xl = x - i < 0 ? 0 : x - i, xr = x + i > w - 1 ? w - 1 : x + i;
                    float pair = (float)row[(size_t)xl * 3] + (float)row[(size_t)xr * 3];
                    float prod = mask[i] * pair;
                    s = s + prod;
                }
                tmp[(size_t)y * w + x] = s;
            }
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float s = mask[0] * tmp[(size_t)y * w + x];
                for (int i = 1; i < len; ++i) {
                    int yu = y - i < 0 ? 0 : y - i, yd = y + i > h - 1 ? h - 1 : y + i;
                    float pair = tmp[(size_t)yu * w + x] + tmp[(size_t)yd * w + x];
                    float prod = mask[i] * pair;
                    s = s + prod;
                }
                out[(size_t)y * w + x] = s;
            }
    }
    free(tmp);
}

/* Sobel gradient magnitude of the blurred intensity, clamped borders (Report.pdf p4 section 3.2.4
 * "a simple Sobel filter"; the exact operator is [D]). */
void orc_sobel(const float *planes, int w, int h, float *G) {
    size_t V = (size_t)w * h;
    float *I = (float *)malloc(V * sizeof(float));
    for (size_t p = 0; p < V; ++p) {
        float s = planes[p] + planes[V + p];
        s = s + planes[2 * V + p];
        I[p] = s * 0.33333334f;
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
            int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
#define AT(xx, yy) I[(size_t)(yy) * w + (xx)]
            float r = (AT(xp, ym) + 2.0f * AT(xp, y)) + AT(xp, yp);
            float l = (AT(xm, ym) + 2.0f * AT(xm, y)) + AT(xm, yp);
            float d = (AT(xm, yp) + 2.0f * AT(x, yp)) + AT(xp, yp);
            float u = (AT(xm, ym) + 2.0f * AT(x, ym)) + AT(xp, ym);
#undef AT
            float gx = r - l, gy = d - u;
            float gx2 = gx * gx, gy2 = gy * gy;
            G[(size_t)y * w + x] = sqrtf(gx2 + gy2);
        }
    free(I);
}

/* ------------------------------------------------------------------------------------------------
 * Grid graph.  Edge index idx = d*V + p (direction-major), p = y*w + x, V = w*h, D = 2
 * (4-connected: E,S) or 4 (8-connected: E,S,SE,NE -- the F&H `segment` neighbour set).  Absent
 * edges carry +inf.  This index is the tie-break of every comparison ("ties broken by edge index",
 * BASELINE.json north_star); which fixed numbering is used is a free choice ([D]) because the
 * reference's own std::sort leaves the order of equal weights unspecified.
 * ---------------------------------------------------------------------------------------------- */
static const int DX[4] = {1, 0, 1, 1};
static const int DY[4] = {0, 1, 1, -1};

static inline int dirs_of(int conn) { return conn == 8 ? 4 : 2; }

static inline float l2rgb(const float *pl, size_t V, size_t p, size_t q) {
    float dr = pl[p] - pl[q], dg = pl[V + p] - pl[V + q], db = pl[2 * V + p] - pl[2 * V + q];
    float r2 = dr * dr, g2 = dg * dg, b2 = db * db;
    float s = r2 + g2;
    s = s + b2;
    return sqrtf(s);
}

/* wts[idx] = ||rgb(p) - rgb(q)||_2 on the blurred planes (Report.pdf p2 par.1). Returns #edges. */
int64_t orc_edges(const float *planes, int w, int h, int conn, float *wts) {
    int D = dirs_of(conn);
    size_t V = (size_t)w * h;
    int64_t n = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t p = (size_t)y * w + x;
            for (int d = 0; d < D; ++d) {
                int xx = x + DX[d], yy = y + DY[d];
                if (xx < 0 || xx >= w || yy < 0 || yy >= h) { wts[(size_t)d * V + p] = INFINITY; continue; }
                wts[(size_t)d * V + p] = l2rgb(planes, V, p, (size_t)yy * w + xx);
                ++n;
            }
        }
    return n;
}

/* Superpixel static edge strength: mean Sobel magnitude of the two end pixels ([D]). */
void orc_strength(const float *G, int w, int h, int conn, float *str) {
    int D = dirs_of(conn);
    size_t V = (size_t)w * h;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t p = (size_t)y * w + x;
            for (int d = 0; d < D; ++d) {
                int xx = x + DX[d], yy = y + DY[d];
                if (xx < 0 || xx >= w || yy < 0 || yy >= h) { str[(size_t)d * V + p] = INFINITY; continue; }
                float s = G[p] + G[(size_t)yy * w + xx];
                str[(size_t)d * V + p] = 0.5f * s;
            }
        }
}

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* LSD radix sort of u64 keys (4 x 16-bit digits). */
static void radix_sort_u64(uint64_t *a, size_t n) {
    uint64_t *b = (uint64_t *)malloc(n * sizeof(uint64_t));
    size_t *cnt = (size_t *)malloc(65536 * sizeof(size_t));
    for (int pass = 0; pass < 4; ++pass) {
        int sh = pass * 16;
        memset(cnt, 0, 65536 * sizeof(size_t));
        for (size_t i = 0; i < n; ++i) cnt[(a[i] >> sh) & 0xFFFF]++;
        size_t s = 0;
        for (int i = 0; i < 65536; ++i) { size_t c = cnt[i]; cnt[i] = s; s += c; }
        for (size_t i = 0; i < n; ++i) b[cnt[(a[i] >> sh) & 0xFFFF]++] = a[i];
        uint64_t *t = a; a = b; b = t;
    }
    free(b); /* 4 passes: data ends in the original buffer */
    free(cnt);
}

/* ------------------------------------------------------------------------------------------------
 * (a) Kruskal Felzenszwalb -- the report's "CPU baseline" (Report.pdf p1-2 section 2.1, p4
 * "Baseline"; BASELINE.json configs[0]).  Union by rank + path compression.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t p, rank, size; } uf_t;
static int32_t uf_find(uf_t *u, int32_t x) {
    int32_t r = x;
    while (u[r].p != r) r = u[r].p;
    while (u[x].p != r) { int32_t n = u[x].p; u[x].p = r; x = n; }
    return r;
}
static int32_t uf_join(uf_t *u, int32_t a, int32_t b) {
    if (u[a].rank > u[b].rank) { u[b].p = a; u[a].size += u[b].size; return a; }
    u[a].p = b; u[b].size += u[a].size;
    if (u[a].rank == u[b].rank) u[b].rank++;
    return b;
}

int orc_felz_kruskal(int w, int h, int conn, const float *wts, float k, int min_size, int32_t *labels) {
    int D = dirs_of(conn);
    size_t V = (size_t)w * h, ncap = V * D, n = 0;
    uint64_t *keys = (uint64_t *)malloc(ncap * sizeof(uint64_t));
    for (size_t i = 0; i < ncap; ++i)
        if (!isinf(wts[i])) keys[n++] = ((uint64_t)fbits(wts[i]) << 32) | (uint64_t)i;
    radix_sort_u64(keys, n);
    uf_t *u = (uf_t *)malloc(V * sizeof(uf_t));
    float *thr = (float *)malloc(V * sizeof(float));
    for (size_t i = 0; i < V; ++i) { u[i].p = (int32_t)i; u[i].rank = 0; u[i].size = 1; thr[i] = k / 1.0f; }
    for (size_t i = 0; i < n; ++i) {
        uint32_t idx = (uint32_t)keys[i];
        float wt = bitsf((uint32_t)(keys[i] >> 32));
        int d = (int)(idx / V); int32_t p = (int32_t)(idx % V);
        int32_t q = p + DY[d] * w + DX[d];
        int32_t a = uf_find(u, p), b = uf_find(u, q);
        if (a != b && wt <= thr[a] && wt <= thr[b]) {
            int32_t r = uf_join(u, a, b);
            float t = k / (float)u[r].size;
            thr[r] = wt + t;
        }
    }
    for (size_t i = 0; i < n; ++i) {
        uint32_t idx = (uint32_t)keys[i];
        int d = (int)(idx / V); int32_t p = (int32_t)(idx % V);
        int32_t q = p + DY[d] * w + DX[d];
        int32_t a = uf_find(u, p), b = uf_find(u, q);
        if (a != b && (u[a].size < min_size || u[b].size < min_size)) uf_join(u, a, b);
    }
    int nc = 0;
    for (size_t i = 0; i < V; ++i) { labels[i] = uf_find(u, (int32_t)i); if (labels[i] == (int32_t)i) ++nc; }
    free(keys); free(u); free(thr);
    return nc;
}

/* ------------------------------------------------------------------------------------------------
 * (b)(c)(d) Round-synchronous Boruvka segmentation.
 *
 *   variant 0 FELZ     Report.pdf p2-3 section 3.1 steps 1-9 + Alg.1: each component's minimum
 *                      outgoing edge, 2-cycle removal, predicate w <= Int(C)+k/|C| on BOTH sides
 *                      evaluated on pre-round Int/size, simultaneous contraction; when a round
 *                      marks nothing, "post-processing" = min-size rounds in which only components
 *                      smaller than min_size select (and always take) their minimum edge.
 *   variant 1 HIER     Report.pdf p2 section 2.2, p3-4 section 3.2.2-3.2.3: no predicate, every round is a
 *                      hierarchy level, until one component remains.
 *   variant 2 SUPERPIX Report.pdf p4 section 3.2.4: as HIER but the weight of an edge in round r is
 *                      strength(e) * ||mean colour(Cu) - mean colour(Cv)||_2 with the components of
 *                      round r; component colour is kept as exact integer sums of the blurred
 *                      colour in 24.8 fixed point ([D], makes the result order-independent).
 *
 * Total order on edges everywhere: key = (fp32 bits of weight) << 32 | edge index.
 *
 * stats (may be NULL): per executed round 4 x int64 {components before, live edges before,
 * components merged away, phase(0 predicate/levels, 1 min-size)}.
 * levels_out (may be NULL): for HIER/SUPERPIX level L (0-based) labels at levels_out[L*V..],
 * written for L < max_levels.  labels: final partition (FELZ) / last level reached.
 * Returns number of rounds that merged something (= number of levels for HIER/SUPERPIX).
 * ---------------------------------------------------------------------------------------------- */
/* Optional export of the final Int(C) per representative pixel (the tiled schedule joins strips by their
 * component attributes); thread-local so that concurrent runs do not interfere. */
static __thread float *g_int_out = NULL;
void orc_set_int_out(float *p) { g_int_out = p; }

int orc_boruvka(int w, int h, int conn, int variant, const float *wts, float k, int min_size, int max_rounds,
                const float *planes, int32_t *labels, int32_t *levels_out, int max_levels,
                int32_t *ncomp_levels, int64_t *stats, int stats_cap, int *final_ncomp) {
    int D = dirs_of(conn);
    size_t V = (size_t)w * h, ncap = V * D;
    int32_t *comp = (int32_t *)malloc(V * sizeof(int32_t));
    int32_t *size = (int32_t *)malloc(V * sizeof(int32_t));
    int32_t *nsize = (int32_t *)malloc(V * sizeof(int32_t));
    float *Int = (float *)malloc(V * sizeof(float));
    float *nInt = (float *)malloc(V * sizeof(float));
    uint64_t *best = (uint64_t *)malloc(V * sizeof(uint64_t));
    int32_t *choice = (int32_t *)malloc(V * sizeof(int32_t));
    int32_t *succ = (int32_t *)malloc(V * sizeof(int32_t));
    int32_t *reps = (int32_t *)malloc(V * sizeof(int32_t));
    uint32_t *live = (uint32_t *)malloc(ncap * sizeof(uint32_t));
    int64_t *csum = NULL, *ncsum = NULL;
    size_t nlive = 0, nrep = V;
    for (size_t i = 0; i < ncap; ++i) if (!isinf(wts[i])) live[nlive++] = (uint32_t)i;
    for (size_t p = 0; p < V; ++p) { comp[p] = (int32_t)p; size[p] = 1; Int[p] = 0.0f; reps[p] = (int32_t)p; }
    if (variant == 2) {
        csum = (int64_t *)malloc(3 * V * sizeof(int64_t));
        ncsum = (int64_t *)malloc(3 * V * sizeof(int64_t));
        for (size_t p = 0; p < V; ++p)
            for (int c = 0; c < 3; ++c) csum[3 * p + c] = (int64_t)lrintf(planes[(size_t)c * V + p] * 256.0f);
    }
    int phase = 0, levels = 0, nstat = 0;
    for (int round = 0; round < max_rounds; ++round) {
        for (size_t i = 0; i < nrep; ++i) best[reps[i]] = KEY_NONE;
        /* step 1+2: minimum outgoing edge per component; drop edges that became internal */
        size_t nl2 = 0;
        for (size_t i = 0; i < nlive; ++i) {
            uint32_t idx = live[i];
            int d = (int)(idx / V); int32_t p = (int32_t)(idx % V);
            int32_t q = p + DY[d] * w + DX[d];
            int32_t a = comp[p], b = comp[q];
            if (a == b) continue;
            live[nl2++] = idx;
            float wt = wts[idx];
            if (variant == 2) {
                float fa = (float)size[a] * 256.0f, fb = (float)size[b] * 256.0f;
                float dr = (float)csum[3 * a] / fa - (float)csum[3 * b] / fb;
                float dg = (float)csum[3 * a + 1] / fa - (float)csum[3 * b + 1] / fb;
                float db = (float)csum[3 * a + 2] / fa - (float)csum[3 * b + 2] / fb;
                float r2 = dr * dr, g2 = dg * dg, b2 = db * db;
                float s = r2 + g2;
                s = s + b2;
                wt = wts[idx] * sqrtf(s);
            }
            uint64_t key = ((uint64_t)fbits(wt) << 32) | idx;
            if (key < best[a]) best[a] = key;
            if (key < best[b]) best[b] = key;
        }
        nlive = nl2;
        /* step 5: predicate (or min-size activity) decides each component's own choice */
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i];
            choice[c] = c;
            if (best[c] == KEY_NONE) continue;
            uint32_t idx = (uint32_t)best[c];
            float wt = bitsf((uint32_t)(best[c] >> 32));
            int d = (int)(idx / V); int32_t p = (int32_t)(idx % V);
            int32_t q = p + DY[d] * w + DX[d];
            int32_t a = comp[p], b = comp[q];
            int32_t other = a == c ? b : a;
            int ok;
            if (variant != 0) ok = 1;
            else if (phase == 0) {
                float ta = k / (float)size[a], tb = k / (float)size[b];
                ta = Int[a] + ta; tb = Int[b] + tb;
                ok = wt <= ta && wt <= tb;
            } else ok = size[c] < min_size;
            if (ok) choice[c] = other;
        }
        /* step 4: remove 2-cycles (the lower id of a mutual pair becomes the root) */
        size_t merged = 0;
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i], s = choice[c];
            if (s != c && choice[s] == c && c < s) s = c;
            succ[c] = s;
            if (s != c) ++merged;
        }
        if (stats && nstat < stats_cap) {
            stats[4 * nstat] = (int64_t)nrep; stats[4 * nstat + 1] = (int64_t)nlive;
            stats[4 * nstat + 2] = (int64_t)merged; stats[4 * nstat + 3] = phase; ++nstat;
        }
        if (merged == 0) {
            if (variant == 0 && phase == 0 && min_size > 1) { phase = 1; continue; }
            break;
        }
        /* steps 7+8: flatten the merge forest; accumulate size, Int (and colour) into the roots */
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i];
            nsize[c] = 0; nInt[c] = 0.0f;
            if (variant == 2) ncsum[3 * c] = ncsum[3 * c + 1] = ncsum[3 * c + 2] = 0;
        }
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i], r = c;
            while (succ[r] != r) r = succ[r];
            choice[c] = r; /* reuse as root[] */
        }
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i], r = choice[c];
            nsize[r] += size[c];
            float m = Int[c];
            if (succ[c] != c) { float wt = bitsf((uint32_t)(best[c] >> 32)); if (wt > m) m = wt; }
            if (m > nInt[r]) nInt[r] = m;
            if (variant == 2) for (int ch = 0; ch < 3; ++ch) ncsum[3 * r + ch] += csum[3 * c + ch];
        }
        for (size_t p = 0; p < V; ++p) comp[p] = choice[comp[p]];
        size_t nr2 = 0;
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i];
            if (choice[c] == c) {
                reps[nr2++] = c; size[c] = nsize[c]; Int[c] = nInt[c];
                if (variant == 2) for (int ch = 0; ch < 3; ++ch) csum[3 * c + ch] = ncsum[3 * c + ch];
            }
        }
        nrep = nr2;
        if (variant != 0) {
            if (levels_out && levels < max_levels) memcpy(levels_out + (size_t)levels * V, comp, V * sizeof(int32_t));
            if (ncomp_levels && levels < max_levels) ncomp_levels[levels] = (int32_t)nrep;
        }
        ++levels;
        if (variant != 0 && (nrep <= 1 || levels >= max_levels)) break;
    }
    memcpy(labels, comp, V * sizeof(int32_t));
    if (g_int_out) for (size_t i = 0; i < nrep; ++i) g_int_out[reps[i]] = Int[reps[i]];
    if (final_ncomp) *final_ncomp = (int)nrep;
    free(comp); free(size); free(nsize); free(Int); free(nInt); free(best); free(choice); free(succ);
    free(reps); free(live); free(csum); free(ncsum);
    return levels;
}

/* ------------------------------------------------------------------------------------------------
 * The same rounds on an EXPLICIT graph: nv components with (size, Int), ne undirected edges (ea, eb, w)
 * whose position in the list is the tie-break.  This is the second phase of the tiled schedule
 * (BASELINE.json north_star: "cross-tile boundary edges exchanged ... before the final Boruvka
 * rounds"; DESIGN.md "Tiled schedule"): the strips' final component graphs joined by the cut edges.
 * variant 0 (FELZ: predicate rounds, then min-size rounds) or 1 (HIER: until one component).
 * labels_out[c] = representative of input component c.  Returns the number of final components.
 * ---------------------------------------------------------------------------------------------- */
int orc_boruvka_graph(int nv, const int32_t *size0, const float *Int0, int64_t ne, const int32_t *ea, const int32_t *eb,
                      const float *w, int variant, float k, int min_size, int max_rounds, int32_t *labels_out,
                      int64_t *stats, int stats_cap) {
    int32_t *comp = (int32_t *)malloc((size_t)nv * sizeof(int32_t));
    int32_t *size = (int32_t *)malloc((size_t)nv * sizeof(int32_t));
    int32_t *nsize = (int32_t *)malloc((size_t)nv * sizeof(int32_t));
    float *Int = (float *)malloc((size_t)nv * sizeof(float));
    float *nInt = (float *)malloc((size_t)nv * sizeof(float));
    uint64_t *best = (uint64_t *)malloc((size_t)nv * sizeof(uint64_t));
    int32_t *choice = (int32_t *)malloc((size_t)nv * sizeof(int32_t));
    int32_t *succ = (int32_t *)malloc((size_t)nv * sizeof(int32_t));
    int32_t *reps = (int32_t *)malloc((size_t)nv * sizeof(int32_t));
    uint32_t *live = (uint32_t *)malloc((size_t)(ne > 0 ? ne : 1) * sizeof(uint32_t));
    size_t nlive = 0, nrep = (size_t)nv;
    for (int64_t i = 0; i < ne; ++i) live[nlive++] = (uint32_t)i;
    for (int c = 0; c < nv; ++c) { comp[c] = c; size[c] = size0[c]; Int[c] = Int0[c]; reps[c] = c; }
    int phase = 0, nstat = 0;
    for (int round = 0; round < max_rounds; ++round) {
        for (size_t i = 0; i < nrep; ++i) best[reps[i]] = KEY_NONE;
        size_t nl2 = 0;
        for (size_t i = 0; i < nlive; ++i) {
            const uint32_t e = live[i];
            const int32_t a = comp[ea[e]], b = comp[eb[e]];
            if (a == b) continue;
            live[nl2++] = e;
            const uint64_t key = ((uint64_t)fbits(w[e]) << 32) | e;
            if (key < best[a]) best[a] = key;
            if (key < best[b]) best[b] = key;
        }
        nlive = nl2;
        for (size_t i = 0; i < nrep; ++i) {
            const int32_t c = reps[i];
            choice[c] = c;
            if (best[c] == KEY_NONE) continue;
            const uint32_t e = (uint32_t)best[c];
            const float wt = bitsf((uint32_t)(best[c] >> 32));
            const int32_t a = comp[ea[e]], b = comp[eb[e]];
            const int32_t other = a == c ? b : a;
            int ok;
            if (variant != 0) ok = 1;
            else if (phase == 0) {
                float ta = k / (float)size[a], tb = k / (float)size[b];
                ta = Int[a] + ta; tb = Int[b] + tb;
                ok = wt <= ta && wt <= tb;
            } else ok = size[c] < min_size;
            if (ok) choice[c] = other;
        }
        size_t merged = 0;
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i], s = choice[c];
            if (s != c && choice[s] == c && c < s) s = c;
            succ[c] = s;
            if (s != c) ++merged;
        }
        if (stats && nstat < stats_cap) {
            stats[4 * nstat] = (int64_t)nrep; stats[4 * nstat + 1] = (int64_t)nlive;
            stats[4 * nstat + 2] = (int64_t)merged; stats[4 * nstat + 3] = phase; ++nstat;
        }
        if (merged == 0) {
            if (variant == 0 && phase == 0 && min_size > 1) { phase = 1; continue; }
            break;
        }
        for (size_t i = 0; i < nrep; ++i) { nsize[reps[i]] = 0; nInt[reps[i]] = 0.0f; }
        for (size_t i = 0; i < nrep; ++i) {
            int32_t c = reps[i], r = c;
            while (succ[r] != r) r = succ[r];
            choice[c] = r;
        }
        for (size_t i = 0; i < nrep; ++i) {
            const int32_t c = reps[i], r = choice[c];
            nsize[r] += size[c];
            float m = Int[c];
            if (succ[c] != c) { const float wt = bitsf((uint32_t)(best[c] >> 32)); if (wt > m) m = wt; }
            if (m > nInt[r]) nInt[r] = m;
        }
        for (int c = 0; c < nv; ++c) comp[c] = choice[comp[c]];
        size_t nr2 = 0;
        for (size_t i = 0; i < nrep; ++i) {
            const int32_t c = reps[i];
            if (choice[c] == c) { reps[nr2++] = c; size[c] = nsize[c]; Int[c] = nInt[c]; }
        }
        nrep = nr2;
        if (variant != 0 && nrep <= 1) break;
    }
    memcpy(labels_out, comp, (size_t)nv * sizeof(int32_t));
    free(comp); free(size); free(nsize); free(Int); free(nInt); free(best); free(choice); free(succ); free(reps); free(live);
    return (int)nrep;
}

/* Canonical relabelling: ids 0..n-1 in order of first appearance; two partitions are equal iff
 * their canonical label images are equal.  Labels must lie in [0, V). Returns n. */
int orc_canon(int32_t *labels, int64_t V) {
    int32_t *map = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    memset(map, 0xFF, (size_t)V * sizeof(int32_t));
    int32_t n = 0;
    for (int64_t i = 0; i < V; ++i) {
        int32_t l = labels[i];
        if (map[l] < 0) map[l] = n++;
        labels[i] = map[l];
    }
    free(map);
    return n;
}

/* Whole-pipeline convenience used by the CPU-baseline timers: u8 image in, labels out. */
int orc_segment(const uint8_t *rgb, int w, int h, float sigma, float k, int min_size, int conn, int variant,
                int max_rounds, int32_t *labels) {
    size_t V = (size_t)w * h;
    int D = dirs_of(conn);
    float *planes = (float *)malloc(3 * V * sizeof(float));
    float *wts = (float *)malloc(V * D * sizeof(float));
    orc_blur(rgb, w, h, sigma, planes);
    int n = 0;
    if (variant == 3) { /* Kruskal baseline */
        orc_edges(planes, w, h, conn, wts);
        n = orc_felz_kruskal(w, h, conn, wts, k, min_size, labels);
    } else {
        if (variant == 2) {
            float *G = (float *)malloc(V * sizeof(float));
            orc_sobel(planes, w, h, G);
            orc_strength(G, w, h, conn, wts);
            free(G);
        } else orc_edges(planes, w, h, conn, wts);
        orc_boruvka(w, h, conn, variant, wts, k, min_size, max_rounds, planes, labels, NULL, max_rounds, NULL,
                    NULL, 0, &n);
    }
    free(planes); free(wts);
    return n;
}
