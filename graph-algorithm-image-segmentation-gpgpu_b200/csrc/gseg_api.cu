// gseg_api.cu -- host side of libgseg.so: context, round scheduling, the C-ABI of include/gseg.h.
//
// There is no CPU path in this file: every compute entry point launches the CUDA kernels of
// gseg_kernels.cuh and fails with GSEG_E_CUDA when no device is usable.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gseg.h"
#include "gseg_kernels.cuh"
#include "gseg_sort.cuh"

#define GSEG_MAXMARK 1024
#define NSM 148
#define GRID_CAP (NSM * 8)
#define PERSIST_GRID (NSM * 4)

struct gseg_ctx {
    int device, max_w, max_h;
    size_t Vmax;
    cudaStream_t stream, own_stream;
    uint8_t *d_rgb, *d_dir0;
    float *d_tmp, *d_planes, *d_G, *d_wgrid, *d_export;
    u32 *d_wsel, *d_succ, *d_rank;
    u64 *d_best[2];
    u32 *d_size[2], *d_int[2];
    long long *d_csum[2];
    u32 *d_ea[2], *d_eb[2], *d_ew[2];
    u32 *d_arena;
    size_t arena_cap;
    u64 *d_statusC, *d_statusE;
    size_t ntilesC, ntilesE;
    int *d_labels[2];
    GsegCtl *d_ctl, *h_ctl;
    GsegRunParams *h_params;
    SortScratch sort;
    // run state
    gseg_params params;
    int w, h, D;
    bool valid, pending;
    u32 epoch_next;
    char err[256];
    // per-kernel profiling (host-driven schedule only) and launch accounting
    bool profiling;
    int n_marks;
    cudaEvent_t ev[GSEG_MAXMARK + 1];
    const char *mark_name[GSEG_MAXMARK];
    int mark_round[GSEG_MAXMARK];
    long long launches, graph_nodes;
    // graph cache
    cudaGraphExec_t gexec;
    int g_w, g_h, g_variant, g_D, g_rounds;
};

static int fail(gseg_ctx *c, int code, const char *what, cudaError_t e) {
    if (c) snprintf(c->err, sizeof(c->err), "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return code;
}
#define CK(call)                                                              \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return fail(ctx, GSEG_E_CUDA, #call, e__);    \
    } while (0)

extern "C" int gseg_version(void) { return GSEG_VERSION; }

extern "C" const char *gseg_strerror(int s) {
    switch (s) {
    case GSEG_OK: return "ok";
    case GSEG_E_ARG: return "bad argument";
    case GSEG_E_CUDA: return "CUDA error or no CUDA device (there is no CPU fallback)";
    case GSEG_E_SIZE: return "image larger than the context capacity";
    case GSEG_E_ARENA: return "supervertex-map arena exhausted";
    case GSEG_E_INTERNAL: return "device-side watchdog tripped";
    case GSEG_E_STATE: return "no completed segmentation in this context";
    case GSEG_E_LEVEL: return "hierarchy level out of range";
    default: return "unknown status";
    }
}
extern "C" const char *gseg_last_error(const gseg_ctx *ctx) { return ctx ? ctx->err : "null context"; }

static inline int grid_for(size_t n, int per_block, int cap = GRID_CAP) {
    size_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > (size_t)cap) g = cap;
    return (int)g;
}

template <typename T>
static cudaError_t dalloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }

extern "C" int gseg_create(gseg_ctx **out, int device, int max_w, int max_h) {
    if (!out || max_w < 1 || max_h < 1) return GSEG_E_ARG;
    *out = nullptr;
    const size_t V = (size_t)max_w * (size_t)max_h;
    if (V * 4 >= 0xFFFFFFFFull) return GSEG_E_SIZE; // 32-bit edge indices
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return GSEG_E_CUDA;
    gseg_ctx *ctx = (gseg_ctx *)calloc(1, sizeof(gseg_ctx));
    if (!ctx) return GSEG_E_ARG;
    ctx->device = device; ctx->max_w = max_w; ctx->max_h = max_h; ctx->Vmax = V;
    ctx->epoch_next = 1;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    ctx->stream = ctx->own_stream;
    const size_t Vp = V + 64; // slack for vector tails
    if (e == cudaSuccess) e = dalloc(&ctx->d_rgb, 3 * Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_dir0, Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_tmp, 3 * Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_planes, 3 * Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_G, Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_wgrid, 4 * Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_wsel, Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_succ, Vp);
    if (e == cudaSuccess) e = dalloc(&ctx->d_rank, Vp);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = dalloc(&ctx->d_best[i], Vp);
        if (e == cudaSuccess) e = dalloc(&ctx->d_size[i], Vp);
        if (e == cudaSuccess) e = dalloc(&ctx->d_int[i], Vp);
        if (e == cudaSuccess) e = dalloc(&ctx->d_ea[i], 4 * Vp);
        if (e == cudaSuccess) e = dalloc(&ctx->d_eb[i], 4 * Vp);
        if (e == cudaSuccess) e = dalloc(&ctx->d_ew[i], 4 * Vp);
        if (e == cudaSuccess) e = dalloc(&ctx->d_labels[i], Vp);
    }
    ctx->arena_cap = 6 * V + 1024;
    if (ctx->arena_cap > 0xFFFFFFF0ull) ctx->arena_cap = 0xFFFFFFF0ull;
    if (e == cudaSuccess) e = dalloc(&ctx->d_arena, ctx->arena_cap);
    ctx->ntilesC = V / TILE_C + 2;
    ctx->ntilesE = 4 * V / TILE_E + 2;
    if (ctx->ntilesE < V / NT + 2) ctx->ntilesE = V / NT + 2;
    if (e == cudaSuccess) e = dalloc(&ctx->d_statusC, ctx->ntilesC);
    if (e == cudaSuccess) e = dalloc(&ctx->d_statusE, ctx->ntilesE);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_statusC, 0, ctx->ntilesC * sizeof(u64));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_statusE, 0, ctx->ntilesE * sizeof(u64));
    if (e == cudaSuccess) e = dalloc(&ctx->d_ctl, 1);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_ctl, 0, sizeof(GsegCtl));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_ctl, sizeof(GsegCtl));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&ctx->h_params, sizeof(GsegRunParams));
    if (e != cudaSuccess) {
        fprintf(stderr, "gseg_create: %s\n", cudaGetErrorString(e));
        gseg_destroy(ctx);
        return GSEG_E_CUDA;
    }
    memset(ctx->h_ctl, 0, sizeof(GsegCtl));
    *out = ctx;
    return GSEG_OK;
}

extern "C" void gseg_destroy(gseg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    if (ctx->gexec) cudaGraphExecDestroy(ctx->gexec);
    cudaFree(ctx->d_rgb); cudaFree(ctx->d_dir0); cudaFree(ctx->d_tmp); cudaFree(ctx->d_planes);
    cudaFree(ctx->d_G); cudaFree(ctx->d_wgrid); cudaFree(ctx->d_export); cudaFree(ctx->d_wsel);
    cudaFree(ctx->d_succ); cudaFree(ctx->d_rank);
    for (int i = 0; i < 2; ++i) {
        cudaFree(ctx->d_best[i]); cudaFree(ctx->d_size[i]); cudaFree(ctx->d_int[i]); cudaFree(ctx->d_csum[i]);
        cudaFree(ctx->d_ea[i]); cudaFree(ctx->d_eb[i]); cudaFree(ctx->d_ew[i]); cudaFree(ctx->d_labels[i]);
    }
    cudaFree(ctx->d_arena); cudaFree(ctx->d_statusC); cudaFree(ctx->d_statusE); cudaFree(ctx->d_ctl);
    sort_scratch_free(&ctx->sort);
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->h_params) cudaFreeHost(ctx->h_params);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    free(ctx);
}

extern "C" int gseg_set_stream(gseg_ctx *ctx, void *s) {
    if (!ctx) return GSEG_E_ARG;
    if (ctx->pending) return GSEG_E_STATE;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    if (ctx->gexec) { cudaGraphExecDestroy(ctx->gexec); ctx->gexec = nullptr; }
    return GSEG_OK;
}

// Gaussian taps (same definition as the oracle's orc_gauss_mask; double arithmetic, one final rounding).
static int gauss_mask(float sigma, float *mask) {
    if (sigma < 0.01f) sigma = 0.01f;
    const int len = (int)ceilf(sigma * 4.0f) + 1;
    if (len > GSEG_MAXMASK) return -1;
    double m[GSEG_MAXMASK], s = 0.0;
    for (int i = 0; i < len; ++i) {
        const double t = (double)i / (double)sigma;
        m[i] = exp(-0.5 * t * t);
    }
    for (int i = 1; i < len; ++i) s += m[i];
    s = 2.0 * s + m[0];
    for (int i = 0; i < len; ++i) mask[i] = (float)(m[i] / s);
    return len;
}

// Called right before every kernel launch: counts it and, when profiling, drops an event so that
// consecutive events bracket exactly one kernel.
static inline void mark(gseg_ctx *c, cudaStream_t s, const char *name, int round) {
    ++c->launches;
    if (!c->profiling || c->n_marks >= GSEG_MAXMARK) return;
    if (!c->ev[c->n_marks]) cudaEventCreate(&c->ev[c->n_marks]);
    cudaEventRecord(c->ev[c->n_marks], s);
    c->mark_name[c->n_marks] = name;
    c->mark_round[c->n_marks] = round;
    ++c->n_marks;
}
// Closes the last kernel of a round before the host read-back gap (a nameless mark).
static inline void mark_end(gseg_ctx *c, cudaStream_t s) {
    if (!c->profiling || c->n_marks == 0 || c->n_marks >= GSEG_MAXMARK) return;
    if (!c->ev[c->n_marks]) cudaEventCreate(&c->ev[c->n_marks]);
    cudaEventRecord(c->ev[c->n_marks], s);
    c->mark_name[c->n_marks] = nullptr;
    c->mark_round[c->n_marks] = -1;
    ++c->n_marks;
}

// ---- round scheduling --------------------------------------------------------------------------
template <int VARIANT>
static void enqueue_round0(gseg_ctx *c, cudaStream_t s) {
    constexpr bool SP = VARIANT == GSEG_SUPERPIX;
    const size_t V = (size_t)c->w * c->h;
    GsegCtl *ctl = c->d_ctl;
    mark(c, s, "k_init", 0);
    k_init<<<1, 32, 0, s>>>(ctl);
    mark(c, s, "k_blur_h", 0);
    k_blur_h<<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_tmp);
    mark(c, s, "k_blur_v", 0);
    k_blur_v<<<grid_for(3 * V, NT), NT, 0, s>>>(ctl, c->d_tmp, c->d_planes);
    if (SP) {
        mark(c, s, "k_sobel", 0);
        k_sobel<<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_planes, c->d_G);
        mark(c, s, "k_weights", 0);
        k_weights<true><<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_planes, c->d_G, c->d_wgrid);
    } else {
        mark(c, s, "k_weights", 0);
        k_weights<false><<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_planes, c->d_G, c->d_wgrid);
    }
    mark(c, s, "k_r0_choose", 0);
    k_r0_choose<VARIANT><<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_wgrid, c->d_planes, c->d_dir0, c->d_wsel);
    mark(c, s, "k_r0_succ", 0);
    k_r0_succ<SP><<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_dir0, c->d_succ, c->d_size[1], c->d_int[1], c->d_best[1],
                                                  c->d_csum[1]);
    mark(c, s, "k_jump", 0);
    k_jump<<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_succ);
    mark(c, s, "k_rootscan", 0);
    k_rootscan<<<grid_for(V, TILE_C, PERSIST_GRID), NT, 0, s>>>(ctl, c->d_succ, c->d_rank, c->d_statusC);
    mark(c, s, "k_relabel", 0);
    k_relabel<true, SP><<<grid_for(V, NT), NT, 0, s>>>(ctl, c->d_succ, c->d_rank, c->d_wsel, nullptr, nullptr, nullptr,
                                                       c->d_planes, c->d_arena, c->d_size[1], c->d_int[1], c->d_csum[1]);
    mark(c, s, "k_r0_edges", 0);
    k_r0_edges<SP><<<grid_for(V, NT, PERSIST_GRID * 2), NT, 0, s>>>(ctl, c->d_wgrid, c->d_arena, c->d_ea[1], c->d_eb[1],
                                                                    c->d_ew[1], c->d_best[1], c->d_size[1], c->d_csum[1],
                                                                    c->d_statusE);
    mark(c, s, "k_advance", 0);
    k_advance<<<1, 32, 0, s>>>(ctl);
}

template <bool SP>
static void enqueue_round(gseg_ctx *c, cudaStream_t s, int r, size_t Vb, size_t Eb) {
    const int cur = r & 1, nxt = cur ^ 1;
    GsegCtl *ctl = c->d_ctl;
    mark(c, s, "k_succ", r);
    k_succ<SP><<<grid_for(Vb, NT, PERSIST_GRID), NT, 0, s>>>(ctl, c->d_best[cur], c->d_ea[cur], c->d_eb[cur], c->d_size[cur],
                                                            c->d_int[cur], c->d_succ, c->d_wsel, c->d_size[nxt],
                                                            c->d_int[nxt], c->d_best[nxt], c->d_csum[nxt]);
    mark(c, s, "k_jump", r);
    k_jump<<<grid_for(Vb, NT, PERSIST_GRID), NT, 0, s>>>(ctl, c->d_succ);
    mark(c, s, "k_rootscan", r);
    k_rootscan<<<grid_for(Vb, TILE_C, PERSIST_GRID), NT, 0, s>>>(ctl, c->d_succ, c->d_rank, c->d_statusC);
    mark(c, s, "k_relabel", r);
    k_relabel<false, SP><<<grid_for(Vb, NT, PERSIST_GRID), NT, 0, s>>>(ctl, c->d_succ, c->d_rank, c->d_wsel, c->d_size[cur],
                                                                      c->d_int[cur], c->d_csum[cur], c->d_planes, c->d_arena,
                                                                      c->d_size[nxt], c->d_int[nxt], c->d_csum[nxt]);
    mark(c, s, "k_edges", r);
    k_edges<SP><<<grid_for(Eb, TILE_E, PERSIST_GRID), NT, 0, s>>>(ctl, c->d_ea[cur], c->d_eb[cur], c->d_ew[cur], c->d_arena,
                                                                 c->d_ea[nxt], c->d_eb[nxt], c->d_ew[nxt], c->d_best[nxt],
                                                                 c->d_size[nxt], c->d_csum[nxt], c->d_statusE);
    mark(c, s, "k_advance", r);
    k_advance<<<1, 32, 0, s>>>(ctl);
}

static void enqueue_round0_v(gseg_ctx *c, cudaStream_t s) {
    switch (c->params.variant) {
    case GSEG_FELZ: enqueue_round0<GSEG_FELZ>(c, s); break;
    case GSEG_HIER: enqueue_round0<GSEG_HIER>(c, s); break;
    default: enqueue_round0<GSEG_SUPERPIX>(c, s); break;
    }
}
static void enqueue_round_v(gseg_ctx *c, cudaStream_t s, int r, size_t Vb, size_t Eb) {
    if (c->params.variant == GSEG_SUPERPIX) enqueue_round<true>(c, s, r, Vb, Eb);
    else enqueue_round<false>(c, s, r, Vb, Eb);
}

static int max_rounds_of(const gseg_params *p) {
    int r = p->max_rounds > 0 ? p->max_rounds : 48;
    return r > GSEG_MAXR ? GSEG_MAXR : r;
}

static int readback(gseg_ctx *ctx) {
    CK(cudaMemcpyAsync(ctx->h_ctl, ctx->d_ctl, sizeof(GsegCtl), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

static int finish(gseg_ctx *ctx) {
    ctx->pending = false;
    if (ctx->h_ctl->error == DERR_SCAN) return fail(ctx, GSEG_E_INTERNAL, "look-back watchdog", cudaSuccess);
    if (ctx->h_ctl->error == DERR_ARENA) return fail(ctx, GSEG_E_ARENA, "map arena", cudaSuccess);
    ctx->valid = true;
    return GSEG_OK;
}

// whole schedule captured once per (w, h, variant, connectivity, max_rounds) and replayed
static int build_graph(gseg_ctx *ctx) {
    const int R = max_rounds_of(&ctx->params);
    if (ctx->gexec && ctx->g_w == ctx->w && ctx->g_h == ctx->h && ctx->g_variant == ctx->params.variant &&
        ctx->g_D == ctx->D && ctx->g_rounds == R)
        return GSEG_OK;
    if (ctx->gexec) { cudaGraphExecDestroy(ctx->gexec); ctx->gexec = nullptr; }
    cudaGraph_t g = nullptr;
    const bool prof = ctx->profiling;
    const long long before = ctx->launches;
    ctx->profiling = false; // events are meaningless inside a captured graph
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const size_t V = (size_t)ctx->w * ctx->h, E = V * ctx->D;
    enqueue_round0_v(ctx, ctx->stream);
    for (int r = 1; r < R; ++r) {
        size_t Vb = V;
        if (ctx->params.variant != GSEG_FELZ) { Vb = V >> (r > 30 ? 30 : r); if (Vb < 1) Vb = 1; }
        enqueue_round_v(ctx, ctx->stream, r, Vb, E);
    }
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    ctx->profiling = prof;
    ctx->graph_nodes = ctx->launches - before;
    ctx->launches = before;
    if (e != cudaSuccess) return fail(ctx, GSEG_E_CUDA, "cudaStreamEndCapture", e);
    e = cudaGraphInstantiate(&ctx->gexec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { ctx->gexec = nullptr; return fail(ctx, GSEG_E_CUDA, "cudaGraphInstantiate", e); }
    ctx->g_w = ctx->w; ctx->g_h = ctx->h; ctx->g_variant = ctx->params.variant; ctx->g_D = ctx->D; ctx->g_rounds = R;
    return GSEG_OK;
}

static int ensure_csum(gseg_ctx *ctx) {
    if (ctx->d_csum[0]) return GSEG_OK;
    for (int i = 0; i < 2; ++i) CK(dalloc(&ctx->d_csum[i], 3 * (ctx->Vmax + 64)));
    return GSEG_OK;
}

extern "C" int gseg_segment_async(gseg_ctx *ctx, const uint8_t *rgb, int w, int h, int stride, int mem_kind,
                                  const gseg_params *p) {
    if (!ctx || !rgb || !p) return GSEG_E_ARG;
    if (ctx->pending) return fail(ctx, GSEG_E_STATE, "previous run not waited for", cudaSuccess);
    if (w < 1 || h < 1 || stride < 3 * w) return fail(ctx, GSEG_E_ARG, "image geometry", cudaSuccess);
    if ((size_t)w * h > ctx->Vmax) return fail(ctx, GSEG_E_SIZE, "image exceeds context capacity", cudaSuccess);
    if (p->connectivity != 4 && p->connectivity != 8) return fail(ctx, GSEG_E_ARG, "connectivity must be 4 or 8", cudaSuccess);
    if (p->variant < GSEG_FELZ || p->variant > GSEG_SUPERPIX) return fail(ctx, GSEG_E_ARG, "variant", cudaSuccess);
    if (!(p->sigma >= 0.0f) || p->max_levels < 0 || p->max_rounds < 0) return fail(ctx, GSEG_E_ARG, "parameter range", cudaSuccess);
    if (mem_kind != GSEG_MEM_HOST && mem_kind != GSEG_MEM_DEVICE) return fail(ctx, GSEG_E_ARG, "mem_kind", cudaSuccess);
    GsegRunParams *hp = ctx->h_params;
    const int len = gauss_mask(p->sigma, hp->mask);
    if (len < 0) return fail(ctx, GSEG_E_ARG, "sigma too large (more than 64 taps)", cudaSuccess);
    CK(cudaSetDevice(ctx->device));
    ctx->valid = false;
    ctx->params = *p;
    ctx->w = w; ctx->h = h; ctx->D = p->connectivity == 8 ? 4 : 2;
    if (p->variant == GSEG_SUPERPIX) { int rc = ensure_csum(ctx); if (rc) return rc; }
    const uint8_t *src = rgb;
    int dstride = stride;
    if (mem_kind == GSEG_MEM_HOST) {
        CK(cudaMemcpy2DAsync(ctx->d_rgb, (size_t)3 * w, rgb, (size_t)stride, (size_t)3 * w, (size_t)h,
                             cudaMemcpyHostToDevice, ctx->stream));
        src = ctx->d_rgb;
        dstride = 3 * w;
    }
    const int R = max_rounds_of(p);
    // look-back tags: 2 per round, 30 bits; recycle the tag space long before it wraps
    if (ctx->epoch_next + 2u * GSEG_MAXR + 8u >= (1u << 30)) {
        CK(cudaMemsetAsync(ctx->d_statusC, 0, ctx->ntilesC * sizeof(u64), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_statusE, 0, ctx->ntilesE * sizeof(u64), ctx->stream));
        ctx->epoch_next = 1;
    }
    hp->rgb = src; hp->w = w; hp->h = h; hp->stride = dstride; hp->D = ctx->D; hp->variant = p->variant;
    hp->k = p->k; hp->min_size = p->min_size; hp->max_rounds = R;
    hp->max_levels = p->max_levels > 0 ? p->max_levels : INT_MAX;
    hp->arena_cap = (u32)ctx->arena_cap;
    hp->epoch_base = ctx->epoch_next;
    hp->mask_len = len;
    ctx->epoch_next += 2u * GSEG_MAXR + 8u;
    CK(cudaMemcpyAsync(&ctx->d_ctl->p, hp, sizeof(GsegRunParams), cudaMemcpyHostToDevice, ctx->stream));

    if (p->flags & GSEG_FLAG_GRAPH) {
        int rc = build_graph(ctx);
        if (rc) return rc;
        CK(cudaGraphLaunch(ctx->gexec, ctx->stream));
        CK(cudaGetLastError());
        ctx->launches += ctx->graph_nodes;
        ctx->pending = true;
        return GSEG_OK;
    }
    // host-driven schedule: one 2 KB read-back per round decides termination and sizes the next grids
    ctx->n_marks = 0;
    enqueue_round0_v(ctx, ctx->stream);
    mark_end(ctx, ctx->stream);
    CK(cudaGetLastError());
    int rc = readback(ctx);
    if (rc) return rc;
    for (int r = 1; r < R && ctx->h_ctl->phase != PH_DONE; ++r) {
        enqueue_round_v(ctx, ctx->stream, r, ctx->h_ctl->Vcur, ctx->h_ctl->Ecur);
        mark_end(ctx, ctx->stream);
        CK(cudaGetLastError());
        rc = readback(ctx);
        if (rc) return rc;
    }
    ctx->pending = true;
    return GSEG_OK;
}

extern "C" int gseg_wait(gseg_ctx *ctx) {
    if (!ctx) return GSEG_E_ARG;
    if (!ctx->pending) return ctx->valid ? GSEG_OK : GSEG_E_STATE;
    CK(cudaSetDevice(ctx->device));
    int rc = readback(ctx);
    if (rc) { ctx->pending = false; return rc; }
    return finish(ctx);
}

extern "C" int gseg_segment(gseg_ctx *ctx, const uint8_t *rgb, int w, int h, int stride, int mem_kind,
                            const gseg_params *p) {
    int rc = gseg_segment_async(ctx, rgb, w, h, stride, mem_kind, p);
    if (rc) return rc;
    return gseg_wait(ctx);
}

extern "C" int gseg_num_levels(const gseg_ctx *ctx) {
    if (!ctx || !ctx->valid) return GSEG_E_STATE;
    return ctx->params.variant == GSEG_FELZ ? 1 : (int)ctx->h_ctl->levels;
}

extern "C" int gseg_num_components(const gseg_ctx *ctx, int level) {
    if (!ctx || !ctx->valid) return GSEG_E_STATE;
    const int nl = gseg_num_levels(ctx);
    if (level < 0) level = nl - 1;
    if (ctx->params.variant == GSEG_FELZ) return level == 0 ? (int)ctx->h_ctl->Vcur : GSEG_E_LEVEL;
    if (nl == 0 && level == -1) return (int)ctx->h_ctl->Vcur;
    if (level >= nl) return GSEG_E_LEVEL;
    return (int)ctx->h_ctl->stVafter[level];
}

static int level_to_round(const gseg_ctx *ctx, int level, int *round) {
    const int rounds = (int)ctx->h_ctl->round;
    if (ctx->params.variant == GSEG_FELZ) {
        if (level != 0 && level != -1) return GSEG_E_LEVEL;
        *round = rounds - 1;
        return GSEG_OK;
    }
    const int nl = (int)ctx->h_ctl->levels;
    if (level < 0) level = nl - 1;
    if (level >= nl && !(nl == 0 && level <= 0)) return GSEG_E_LEVEL;
    if (level < 0) level = 0; // no merging round at all (single pixel): round 0's identity map
    *round = level;
    return GSEG_OK;
}

extern "C" int gseg_labels(gseg_ctx *ctx, int level, int32_t *out, int mem_kind) {
    if (!ctx || !out) return GSEG_E_ARG;
    if (!ctx->valid) return GSEG_E_STATE;
    int round;
    int rc = level_to_round(ctx, level, &round);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t V = (size_t)ctx->w * ctx->h;
    int *dst = mem_kind == GSEG_MEM_DEVICE ? out : ctx->d_labels[0];
    k_compose<<<grid_for(V, NT), NT, 0, ctx->stream>>>(ctx->d_ctl, ctx->d_arena, round, dst);
    CK(cudaGetLastError());
    if (mem_kind != GSEG_MEM_DEVICE)
        CK(cudaMemcpyAsync(out, dst, V * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

extern "C" int gseg_labels_all(gseg_ctx *ctx, int32_t *out, int max_levels, int mem_kind) {
    if (!ctx || !out || max_levels < 1) return GSEG_E_ARG;
    if (!ctx->valid) return GSEG_E_STATE;
    CK(cudaSetDevice(ctx->device));
    int nl = gseg_num_levels(ctx);
    if (nl > max_levels) nl = max_levels;
    const size_t V = (size_t)ctx->w * ctx->h;
    if (ctx->params.variant == GSEG_FELZ) {
        int rc = gseg_labels(ctx, -1, out, mem_kind);
        return rc ? rc : 1;
    }
    const int *prev = nullptr;
    for (int l = 0; l < nl; ++l) {
        int *dst = mem_kind == GSEG_MEM_DEVICE ? out + (size_t)l * V : ctx->d_labels[l & 1];
        if (l == 0) k_compose<<<grid_for(V, NT), NT, 0, ctx->stream>>>(ctx->d_ctl, ctx->d_arena, 0, dst);
        else k_compose_step<<<grid_for(V, NT), NT, 0, ctx->stream>>>(ctx->d_ctl, ctx->d_arena, l, prev, dst);
        if (mem_kind != GSEG_MEM_DEVICE)
            CK(cudaMemcpyAsync(out + (size_t)l * V, dst, V * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        prev = dst;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return nl;
}

extern "C" int gseg_colorize(gseg_ctx *ctx, int level, uint64_t seed, uint8_t *out, int mem_kind) {
    if (!ctx || !out) return GSEG_E_ARG;
    if (!ctx->valid) return GSEG_E_STATE;
    int round;
    int rc = level_to_round(ctx, level, &round);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t V = (size_t)ctx->w * ctx->h;
    k_compose<<<grid_for(V, NT), NT, 0, ctx->stream>>>(ctx->d_ctl, ctx->d_arena, round, ctx->d_labels[0]);
    // d_tmp is free after the run; 3V bytes fit easily
    uint8_t *dst = mem_kind == GSEG_MEM_DEVICE ? out : (uint8_t *)ctx->d_tmp;
    k_colorize<<<grid_for(V, NT), NT, 0, ctx->stream>>>(ctx->d_labels[0], (u32)V, seed, dst);
    CK(cudaGetLastError());
    if (mem_kind != GSEG_MEM_DEVICE) CK(cudaMemcpyAsync(out, dst, 3 * V, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

extern "C" int gseg_weights(gseg_ctx *ctx, float *out, int mem_kind) {
    if (!ctx || !out) return GSEG_E_ARG;
    if (!ctx->valid) return GSEG_E_STATE;
    CK(cudaSetDevice(ctx->device));
    const size_t V = (size_t)ctx->w * ctx->h, n = V * ctx->D;
    float *dst = out;
    if (mem_kind != GSEG_MEM_DEVICE) {
        if (!ctx->d_export) CK(dalloc(&ctx->d_export, 4 * (ctx->Vmax + 64)));
        dst = ctx->d_export;
    }
    k_weights_export<<<grid_for(n, NT), NT, 0, ctx->stream>>>(ctx->d_wgrid, (u32)V, ctx->D, dst);
    CK(cudaGetLastError());
    if (mem_kind != GSEG_MEM_DEVICE) CK(cudaMemcpyAsync(out, dst, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

extern "C" int gseg_blurred(gseg_ctx *ctx, float *out, int mem_kind) {
    if (!ctx || !out) return GSEG_E_ARG;
    if (!ctx->valid) return GSEG_E_STATE;
    CK(cudaSetDevice(ctx->device));
    const size_t V = (size_t)ctx->w * ctx->h;
    CK(cudaMemcpyAsync(out, ctx->d_planes, 3 * V * sizeof(float),
                       mem_kind == GSEG_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

extern "C" int gseg_stats(const gseg_ctx *ctx, gseg_round_stat *out, int cap) {
    if (!ctx || !ctx->valid) return GSEG_E_STATE;
    const int n = (int)ctx->h_ctl->round;
    for (int i = 0; i < n && i < cap && out; ++i) {
        out[i].n_components = ctx->h_ctl->stV[i];
        out[i].n_edges = i == 0 ? 0 : ctx->h_ctl->stE[i];
        out[i].n_merged = ctx->h_ctl->stM[i];
        out[i].phase = (int32_t)ctx->h_ctl->stP[i];
        out[i].reserved = 0;
    }
    return n;
}

extern "C" int gseg_synth(gseg_ctx *ctx, uint8_t *out, int w, int h, uint64_t seed, int mem_kind) {
    if (!ctx || !out || w < 1 || h < 1) return GSEG_E_ARG;
    if ((size_t)w * h > ctx->Vmax) return GSEG_E_SIZE;
    if (ctx->pending) return GSEG_E_STATE;
    CK(cudaSetDevice(ctx->device));
    const size_t V = (size_t)w * h;
    uint8_t *dst = mem_kind == GSEG_MEM_DEVICE ? out : ctx->d_rgb;
    k_synth<<<grid_for(V, NT), NT, 0, ctx->stream>>>(dst, w, h, seed);
    CK(cudaGetLastError());
    if (mem_kind != GSEG_MEM_DEVICE) CK(cudaMemcpyAsync(out, dst, 3 * V, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

extern "C" int gseg_sort_pairs_u64(gseg_ctx *ctx, uint64_t *keys, uint32_t *vals, int64_t n, int begin_bit, int end_bit) {
    if (!ctx || !keys || n < 0 || begin_bit < 0 || end_bit > 64 || begin_bit >= end_bit) return GSEG_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = onesweep_sort_pairs(&ctx->sort, (u64 *)keys, (u32 *)vals, (size_t)n, begin_bit, end_bit, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, GSEG_E_CUDA, "onesweep_sort_pairs", e);
    CK(cudaStreamSynchronize(ctx->stream));
    return GSEG_OK;
}

// ---- profiling / accounting ----------------------------------------------------------------------
extern "C" int gseg_set_profiling(gseg_ctx *ctx, int on) {
    if (!ctx) return GSEG_E_ARG;
    ctx->profiling = on != 0;
    ctx->n_marks = 0;
    return GSEG_OK;
}

extern "C" long long gseg_launch_count(const gseg_ctx *ctx) { return ctx ? ctx->launches : 0; }

// Algorithmic bytes of one kernel launch (DESIGN.md "Kernels"): every input array
// read once, every output written once, gathers and atomics at element size.
static double algo_bytes(const gseg_ctx *c, const char *name, int r) {
    const GsegCtl *h = c->h_ctl;
    const double V0 = (double)c->w * c->h, D = c->D;
    const double V = r == 0 ? V0 : h->stV[r], E = r == 0 ? 0 : h->stE[r], Vn = h->stVafter[r];
    const double En = (r + 1 < (int)h->round) ? h->stE[r + 1] : h->Ecur;
    const bool sp = c->params.variant == GSEG_SUPERPIX;
    if (!strcmp(name, "k_blur_h")) return 3 * V0 + 12 * V0;
    if (!strcmp(name, "k_blur_v")) return 12 * V0 + 12 * V0;
    if (!strcmp(name, "k_sobel")) return 12 * V0 + 4 * V0;
    if (!strcmp(name, "k_weights")) return (sp ? 4 : 12) * V0 + 4 * D * V0;
    if (!strcmp(name, "k_r0_choose")) return 4 * D * V0 + (sp ? 12 * V0 : 0) + 5 * V0;
    if (!strcmp(name, "k_r0_succ")) return V0 + 4 * V0 + 16 * V0 + (sp ? 24 * V0 : 0);
    if (!strcmp(name, "k_jump")) return 4 * V + 4 * V + 4 * (V - Vn);
    if (!strcmp(name, "k_rootscan")) return 4 * V + 4 * V;
    if (!strcmp(name, "k_relabel")) return 4 * V + 4 * V + 4 * V + (r == 0 ? 4 * V : 12 * V) + 8 * V + (sp ? 48 * V : 0);
    if (!strcmp(name, "k_r0_edges")) return 4 * V0 + 4 * D * V0 + 12 * En + 16 * En;
    if (!strcmp(name, "k_succ")) return 8 * V + 8 * V + 8 * V + 8 * V + 16 * V + (sp ? 24 * V : 0);
    if (!strcmp(name, "k_edges")) return 12 * E + 8 * E + 12 * En + 16 * En;
    return 0;
}

extern "C" int gseg_profile_read(gseg_ctx *ctx, gseg_kernel_time *out, int cap) {
    if (!ctx || !ctx->valid) return GSEG_E_STATE;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    int n = 0;
    for (int i = 0; i + 1 < ctx->n_marks && n < cap; ++i) {
        // the closing event of mark i is the next mark's event; nameless marks are read-back gaps
        if (!ctx->mark_name[i]) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
        if (out) {
            snprintf(out[n].name, sizeof(out[n].name), "%s", ctx->mark_name[i]);
            out[n].round = ctx->mark_round[i];
            out[n].ms = ms;
            out[n].algo_bytes = algo_bytes(ctx, ctx->mark_name[i], ctx->mark_round[i]);
        }
        ++n;
    }
    return n;
}
