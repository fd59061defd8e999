/*
 * gseg.h -- C-ABI of the B200-native graph-segmentation engine (libgseg.so).
 *
 * Drop-in boundary for the segmentation hot path of
 * akankshabaranwal/graph-algorithm-image-segmentation-GPGPU.  The mounted reference holds no source
 * (SURVEY.md section 0), so no reference header can be cited line by line; each entry point cites the
 * reference *stage* it replaces (Report.pdf page / section) and the parameter list BASELINE.json's
 * north_star fixes: input image, sigma, k, min_size, hierarchy level -> label image + hierarchy.
 * The closest published signature is F&H `segment`'s
 *     image<rgb>* segment_image(image<rgb>* im, float sigma, float c, int min_size, int* num_ccs)
 * (Report.pdf ref [23], the report's CPU baseline, p4 "Baseline"), which gseg_segment + gseg_labels
 * replace.
 *
 * Conventions: plain pointers and sizes only; 0 = success, negative = gseg_status; nothing throws
 * across the boundary; the caller owns every input/output buffer, the context owns all device
 * scratch (allocated once in gseg_create, never inside gseg_segment); one context per GPU per host
 * thread, contexts are independent and not thread-safe.  There is no CPU fallback: every entry
 * point that computes requires a CUDA device and fails with GSEG_E_CUDA otherwise.
 */
#ifndef GSEG_H
#define GSEG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSEG_VERSION 100

/* variant -- which reference branch's semantics to run */
#define GSEG_FELZ 0     /* cuda-mst-naive: Boruvka + Felzenszwalb predicate + min-size (Report p2-3 s3.1) */
#define GSEG_HIER 1     /* fastmst_segment: one hierarchy level per Boruvka round, no predicate (p3-4 s3.2.2-3) */
#define GSEG_SUPERPIX 2 /* superpixel_gpu: per-round re-weighting by Sobel strength x mean colour (p4 s3.2.4) */

#define GSEG_MEM_HOST 0
#define GSEG_MEM_DEVICE 1

/* flags */
#define GSEG_FLAG_HOST_LOOP 1u /* host-driven schedule: one kernel per phase and a read-back per round (the
                                  reference's "ab conventional" driver, Report p3); used for per-kernel timing.
                                  Default (0): the device-resident round state drives every kernel, the host
                                  enqueues the whole run without reading anything back, and all small rounds
                                  run inside one persistent thread-block-cluster kernel. */

typedef enum gseg_status {
    GSEG_OK = 0,
    GSEG_E_ARG = -1,     /* bad argument (null pointer, size, connectivity, variant, sigma range) */
    GSEG_E_CUDA = -2,    /* CUDA runtime error or no device; gseg_last_error() has the text */
    GSEG_E_SIZE = -3,    /* image larger than the context was created for */
    GSEG_E_ARENA = -4,   /* supervertex-map arena exhausted (pathological round count) */
    GSEG_E_INTERNAL = -5, /* device-side watchdog tripped */
    GSEG_E_STATE = -6,   /* result requested before a successful gseg_segment */
    GSEG_E_LEVEL = -7,   /* hierarchy level out of range */
    GSEG_E_UNSUPPORTED = -8 /* optional dependency missing at run time (nvJPEG for gseg_segment_jpeg) */
} gseg_status;

/* Parameters of one segmentation (BASELINE.json north_star: sigma, k, min_size, hierarchy level). */
typedef struct gseg_params {
    float sigma;          /* Gaussian pre-filter (Report p2 s2.1; p3 s3.2 par.2); 0 < 4*sigma+1 <= 64 taps */
    float k;              /* Felzenszwalb scale parameter (Report p2 par.1); FELZ only */
    int32_t min_size;     /* min-size post-merge (Report p3 step 6 "post-processing"); FELZ only */
    int32_t connectivity; /* 4 or 8 (BASELINE.json configs[1], configs[2]) */
    int32_t variant;      /* GSEG_FELZ | GSEG_HIER | GSEG_SUPERPIX */
    int32_t max_levels;   /* HIER/SUPERPIX: stop after this many levels; 0 = until one component */
    int32_t max_rounds;   /* cap on Boruvka rounds (Report p5: 10-20 in practice); 0 = 48 */
    uint32_t flags;
} gseg_params;

/* One row per executed Boruvka round (SURVEY.md section 5 "metrics": V_r, E_r per round). */
typedef struct gseg_round_stat {
    int64_t n_components; /* components entering the round */
    int64_t n_edges;      /* live (inter-component) edges entering the round */
    int64_t n_merged;     /* components merged away by the round */
    int32_t phase;        /* 0 = predicate / hierarchy round, 1 = min-size round */
    int32_t in_tail;      /* 1 when the round ran inside the single-cluster tail kernel */
    float us_end;         /* device clock at the end of the round, microseconds since round 0's graph kernel started */
    float us_S, us_R, us_E; /* tail rounds: duration of the choose/scan, flatten and edge phases (else 0) */
    int32_t n_pages;      /* pages of the edge list entering the round */
    int32_t reserved;
} gseg_round_stat;

typedef struct gseg_ctx gseg_ctx;

int gseg_version(void);
const char *gseg_strerror(int status);
const char *gseg_last_error(const gseg_ctx *ctx);

/* Replaces: per-branch main() device/scratch set-up.  Allocates every
 * device buffer for images up to max_w x max_h on CUDA device `device`. */
int gseg_create(gseg_ctx **out, int device, int max_w, int max_h);
/* Same with the largest connectivity (4 or 8) the context will be used with: a 4-connected-only context
 * needs half the edge-list memory (a 32768 x 32768 image fits one B200 that way).  gseg_create = 8. */
int gseg_create_ex(gseg_ctx **out, int device, int max_w, int max_h, int max_connectivity);
void gseg_destroy(gseg_ctx *ctx);

/* Run the context's work on a caller-owned CUDA stream (cudaStream_t as void*); NULL = own stream. */
int gseg_set_stream(gseg_ctx *ctx, void *cuda_stream);

/* Scheduling knob (no effect on results): a Boruvka round whose graph has at most max_edges live edges
 * and max_components components runs inside the single-cluster tail kernel instead of grid-wide
 * kernels.  (0, 0) disables the tail.  Defaults: 262144 edges, 65536 components. */
int gseg_set_tail(gseg_ctx *ctx, uint32_t max_edges, uint32_t max_components);

/* Scheduling knob (no effect on results): resident blocks per SM the grid-wide kernels are sized for
 * (1..8, default 4).  With several contexts in flight per GPU, 2 leaves room for their kernels to overlap. */
int gseg_set_blocks_per_sm(gseg_ctx *ctx, int blocks);

/* Replaces: L3 pre-filter + L2 graph creation + L1 segmentation core of one reference executable
 * (Report p2 Fig.1; p3 s3.2.1; p2-3 s3.1 steps 1-9; p3-4 s3.2.2; p4 s3.2.4).
 * rgb: interleaved 8-bit RGB, `stride_bytes` per row (>= 3*w), in host or device memory (mem_kind).
 * Returns when the partition is complete on the device. */
int gseg_segment(gseg_ctx *ctx, const uint8_t *rgb, int w, int h, int stride_bytes, int mem_kind,
                 const gseg_params *params);

/* Asynchronous form: enqueue only (host input must be pinned and stay valid); gseg_wait completes it. */
int gseg_segment_async(gseg_ctx *ctx, const uint8_t *rgb, int w, int h, int stride_bytes, int mem_kind,
                       const gseg_params *params);
int gseg_wait(gseg_ctx *ctx);

/* Number of hierarchy levels produced (FELZ: 1).  Replaces: the stored per-round supervertex ids
 * (Report p4 s3.2.3 par.1). */
int gseg_num_levels(const gseg_ctx *ctx);
/* Components at `level` (0-based; -1 = last). */
int gseg_num_components(const gseg_ctx *ctx, int level);

/* Replaces: L0 hierarchy materialisation, one thread per pixel mapping the previous level's ids
 * through the stored supervertex ids (Report p4 s3.2.3).  out: w*h int32, row-major, dense ids in
 * [0, gseg_num_components(level)).  level -1 = last level / the FELZ partition. */
int gseg_labels(gseg_ctx *ctx, int level, int32_t *out, int mem_kind);
/* Asynchronous form of gseg_labels: enqueues the materialisation (and the copy to pinned host memory) on
 * the context's stream and returns; gseg_sync (or any later synchronous call) completes it. */
int gseg_labels_async(gseg_ctx *ctx, int level, int32_t *out, int mem_kind);
int gseg_sync(gseg_ctx *ctx);

/* All levels 0..n-1 in one pass (level l at out + l*w*h); n = min(max_levels, gseg_num_levels). */
int gseg_labels_all(gseg_ctx *ctx, int32_t *out, int max_levels, int mem_kind);

/* Replaces: random colour table + colour image (cuRAND, Report p4 s3.2.3).  out: w*h*3 bytes. */
int gseg_colorize(gseg_ctx *ctx, int level, uint64_t seed, uint8_t *out_rgb, int mem_kind);

/* Round-0 edge weights in edge-index order idx = d*w*h + (y*w+x), d in 0..D-1, D = 2 (E,S) or 4 (E,S,SE,NE);
 * +inf where the edge does not exist (SUPERPIX: the static edge strength).  For the 1e-6 check. */
int gseg_weights(gseg_ctx *ctx, float *out, int mem_kind);
/* Blurred image, 3 planes of w*h floats (R,G,B). */
int gseg_blurred(gseg_ctx *ctx, float *out, int mem_kind);

/* ---- tiled schedule (BASELINE.json configs[4], north_star: "cross-tile boundary edges exchanged over
 * NVLink via NCCL before the final Boruvka rounds"; DESIGN.md "Tiled schedule") ------------------------
 * A strip is segmented like any image; its final component graph is exported, the strips' graphs are
 * joined by the cut edges on the host side of the exchange, and the joined graph runs the same rounds.
 *
 * gseg_export_graph: components (size, Int(C)) and live inter-component edges (a, b, weight; ids = the
 * dense labels of gseg_labels(-1); list order = edge-index order) of the last FELZ/HIER run.  Call with
 * NULL arrays to get the counts, then with host arrays of at least that capacity.
 * dedup != 0: parallel edges between the same two components are reduced to their minimum (weight, list
 * position) -- the reference's sort-based duplicate elimination (Report p3 s3.2.2), on the in-house
 * onesweep radix sort; the partition the joined rounds produce is unchanged, the list is ~100x shorter. */
int gseg_export_graph(gseg_ctx *ctx, int dedup, int64_t *n_components, int64_t *n_edges, uint32_t *size, float *Int,
                      uint32_t *ea, uint32_t *eb, float *w, int64_t cap_components, int64_t cap_edges);
/* Rows y0 .. y0+nrows-1 of the blurred planes: out[3][nrows][w] (the cut-edge weights need the boundary rows). */
int gseg_blurred_rows(gseg_ctx *ctx, int y0, int nrows, float *out, int mem_kind);
/* The Boruvka rounds of `params->variant` (FELZ or HIER) on an explicit graph in host memory: components
 * with (size, Int), edges (ea, eb, w) whose list position is the tie-break.  labels_out[c] = dense final
 * component of input component c; returns the number of final components (or a negative status). */
int gseg_segment_graph(gseg_ctx *ctx, int64_t n_components, const uint32_t *size, const float *Int, int64_t n_edges,
                       const uint32_t *ea, const uint32_t *eb, const float *w, const gseg_params *params,
                       int32_t *labels_out);

/* ---- JPEG input decoded on the GPU --------------------------------------
 * The reference's batch benchmark reads a JPEG data set through cv::imread on the host (README.md:26).
 * Here the compressed bytes go to the GPU: nvJPEG (CUDA toolkit library, loaded with dlopen on first
 * use -- libgseg.so itself does not depend on it) decodes into the context's staged RGB buffer on the
 * context's stream and the usual path runs on it; the decoded image never visits the host.
 *   gseg_jpeg_info          width / height of a JPEG (host only: parses the header).
 *   gseg_segment_jpeg_async decode + enqueue the segmentation (complete it with gseg_wait, or use
 *   gseg_segment_jpeg       the blocking form); *w, *h receive the image size.
 *   gseg_input_rgb          the interleaved RGB image the last run read, when it was staged by the
 *                           context (host input or JPEG); GSEG_E_STATE for caller-owned device input.
 * GSEG_E_UNSUPPORTED when libnvjpeg cannot be loaded, GSEG_E_ARG for data nvJPEG rejects. */
int gseg_jpeg_info(const void *jpeg, size_t nbytes, int *w, int *h);
int gseg_segment_jpeg_async(gseg_ctx *ctx, const void *jpeg, size_t nbytes, const gseg_params *params, int *w, int *h);
int gseg_segment_jpeg(gseg_ctx *ctx, const void *jpeg, size_t nbytes, const gseg_params *params, int *w, int *h);
int gseg_input_rgb(gseg_ctx *ctx, uint8_t *out_rgb, int mem_kind);

/* Per-round statistics of the last run; returns number of rounds (<= cap written). */
int gseg_stats(const gseg_ctx *ctx, gseg_round_stat *out, int cap);

/* Deterministic synthetic input: w*h*3 bytes into host or device memory. */
int gseg_synth(gseg_ctx *ctx, uint8_t *out_rgb, int w, int h, uint64_t seed, int mem_kind);

/* Measurement support (SURVEY.md section 5 "tracing"; section 8d): with profiling on, the host-driven
 * schedule brackets every kernel with CUDA events on the context's stream. */
typedef struct gseg_kernel_time {
    char name[24];
    int32_t round;
    float ms;
    double algo_bytes; /* algorithmic bytes of this launch (DESIGN.md "Kernels") */
} gseg_kernel_time;
int gseg_set_profiling(gseg_ctx *ctx, int on);
int gseg_profile_read(gseg_ctx *ctx, gseg_kernel_time *out, int cap);
/* Kernels launched by this context since creation (graph replays count their kernel nodes). */
long long gseg_launch_count(const gseg_ctx *ctx);

/* Stand-alone primitives of the edge-dedup path (Report p3 s3.2.2
 * "sort"): in-house onesweep radix sort of 64-bit keys with 32-bit payload, on device memory. */
int gseg_sort_pairs_u64(gseg_ctx *ctx, uint64_t *keys, uint32_t *vals, int64_t n, int begin_bit, int end_bit);

#ifdef __cplusplus
}
#endif
#endif /* GSEG_H */
