/*
 * kernel_wrapper.c -- host side of libKernel.so, plain C99.
 *
 * Replaces the reference's host wrapper (/root/reference/KernelFolder/Kernel/Kernel.cu:873-984)
 * and the CUDA-samples helpers it leans on (common/inc/helper_cuda.h): validates the caller's
 * arrays, packs them into one single-precision problem blob, owns the device-resident chain
 * state, launches the sm_100a kernels through the thin C ABI of mh_abi.h and assembles the
 * result block the reference's callers expect.  No CUDA header is included here.
 *
 * Differences from the reference that a caller can observe (SURVEY.md quirk ledger):
 *   Q3  result[i].costs are filled in (the reference returns uninitialised memory);
 *   Q13 an out-of-range object index is impossible; Q14 an all-frozen layout returns instead of
 *       spinning; Q15 the seed can be fixed (env MH_SEED or mhOptions.seed);
 *   Q16 nothing leaks; errors return NULL + KernelLastError() instead of exit(1);
 *   gpuConfig.blockxDim is accepted and ignored (the launch shape is chosen here).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/mh_kernel.h"
#include "mh_abi.h"

#include <pthread.h>
#if defined(__linux__)
#include <sys/mman.h>
#include <unistd.h>
#endif

#define MH_TLS __thread
/* nObjs from which MH_EVAL_FULL runs in its bit-identical memo form.  Measured on B200 with 65536 chains, memo
 * against plain scan (profiles/r2_probe_logs/ab_scan_memo_crossover.log): 16 objects -17 %, 18 +23 %, 20 +1 %,
 * 22 +12 %, 24 +25 %, 26 +30 %, 28 +68 %, 50 2.3x, 200 8.7x.  With few chains the scan holds out longer (1024 chains
 * of 16 objects: scan 4.2e8 against 2.9e8), so the lower threshold applies to jobs of at least 8192 chains. */
#define MH_MEMO_MIN_OBJS 28
#define MH_MEMO_MIN_OBJS_BIG_JOB 18
#define MH_BIG_JOB_CHAINS 8192
#define MH_MAX_CHUNKS 8
#define MH_MAX_BLOCKS_PER_SM 4 /* 128-thread blocks at the chain kernel's register cap (MH_MIN_BLOCKS in mh_kernels.cu) */

static MH_TLS char g_err[512];

static void set_err(const char *fmt, const char *what, int code)
{
    if (code)
        snprintf(g_err, sizeof g_err, fmt, what, mhdev_error_string(code));
    else
        snprintf(g_err, sizeof g_err, "%s", what);
}

#define CU(call)                                                    \
    do {                                                            \
        int _e = (call);                                            \
        if (_e) {                                                   \
            set_err("%s failed: %s", #call, _e);                    \
            goto fail;                                              \
        }                                                           \
    } while (0)

MH_API const char *KernelLastError(void) { return g_err; }

/* ---------------------------------------------------------------------------------------------
 * Problem packing
 * --------------------------------------------------------------------------------------------- */

static float below_or_equal(double v) /* largest float <= v */
{
    float f = (float)v;
    if ((double)f > v)
        f = nextafterf(f, -INFINITY);
    return f;
}

static int align4(int w) { return (w + 3) & ~3; }

typedef struct mhProblem {
    float *blob; /* host copy */
    mhProblemHeader *h;
} mhProblem;

static int check_rect(const rectangle *r, int nverts, const char *what, int idx)
{
    if (r->point1Index < 0 || r->point1Index + 3 >= nverts) {
        snprintf(g_err, sizeof g_err, "%s[%d].point1Index = %d: needs 4 consecutive vertices inside vertices[%d]", what, idx,
                 r->point1Index, nverts);
        return -1;
    }
    return 0;
}

/* AABB constants of the four consecutive vertices from `start` (Kernel.cu:366-401 hoisted). */
static void rect_consts(const vertex *v, int start, float box[4], float *v0x)
{
    double m123x = fmin(v[start + 1].x, fmin(v[start + 2].x, v[start + 3].x));
    double miny = fmin(fmin(v[start].y, v[start + 1].y), fmin(v[start + 2].y, v[start + 3].y));
    double maxx = fmax(fmax(v[start].x, v[start + 1].x), fmax(v[start + 2].x, v[start + 3].x));
    double maxy = fmax(fmax(v[start].y, v[start + 1].y), fmax(v[start + 2].y, v[start + 3].y));
    box[0] = (float)m123x;
    box[1] = (float)miny;
    box[2] = (float)maxx;
    box[3] = (float)maxy;
    *v0x = (float)v[start].x; /* quirk Q6 */
}

/* Fixed-point scale of the clearance term.  ClearanceCosts (Kernel.cu:404-434) is evaluated as an INTEGER sum of
 * rectangle overlaps: coordinates in units of 2^-k, areas in units of 2^-2k, 64-bit accumulator.  Integer addition
 * is associative, so the sum does not depend on the order of its terms -- which is what lets a proposal update it
 * by the few pairs it touches and still hold, bit for bit, the value a from-scratch evaluation computes.
 * k is the largest exponent for which nothing can overflow:
 *   coordinates: every box coordinate is (rectangle constant) + (position), |.| <= E + P, and must stay below 2^30;
 *   the sum    : at most C*n overlaps, each at most (widest possible box: 2(E + P), quirk Q6 makes boxes as wide as
 *                the position) x (tallest rectangle), must stay below 2^62.
 * P bounds every position a chain can visit: the room's AABB (translations clamp to it, Kernel.cu:616-633), the
 * caller's layout(s) (swaps only permute positions).  At the BASELINE rooms k = 21..24, i.e. 5e-7 .. 6e-8 length
 * units -- the float32 evaluation it replaces had ulp(8.0) = 9.5e-7 on the same coordinates. */
static int clearance_scale(double pos_bound, double rect_bound, double rect_height, int n, int C)
{
    const double ext = pos_bound + rect_bound + 1.0;
    int k1 = 30 - (int)ceil(log2(ext));
    const double worst_area = (double)(C > 0 ? C : 1) * (double)n * (2.0 * ext) * (rect_height + 1e-9);
    int k2 = (int)floor((62.0 - log2(worst_area > 1.0 ? worst_area : 1.0)) / 2.0);
    int k = k1 < k2 ? k1 : k2;
    if (k > 30) k = 30;
    if (k < -30) k = -30;
    return k;
}

static int32_t to_fixed(double v, double scale)
{
    const double r = nearbyint(v * scale);
    if (r > 2147483647.0) return 2147483647;
    if (r < -2147483648.0) return (int32_t)(-2147483647 - 1);
    return (int32_t)r;
}

/* extra_layouts / n_extra: further layouts (n objects each) whose positions the fixed-point scale must cover
 * (KernelEvalCosts evaluates layouts no chain produced) */
static int pack_problem(const relationshipStruct *rss, const relationshipAngleStruct *rsa, const positionAndRotation *cfg,
                        const rectangle *clearances, const rectangle *offlimits, const vertex *vertices,
                        const vertex *surfaceRectangle, const Surface *srf, const positionAndRotation *extra_layouts, int n_extra,
                        mhProblem *out)
{
    out->blob = NULL;
    if (!srf || !cfg || !offlimits || !vertices || !surfaceRectangle) {
        set_err("", "null argument", 0);
        return -1;
    }
    const int n = srf->nObjs, C = srf->nClearances, R = srf->nRelationships;
    if (n < 1 || C < 0 || R < 0 || n > 65535) {
        snprintf(g_err, sizeof g_err, "bad counts: nObjs=%d nClearances=%d nRelationships=%d", n, C, R);
        return -1;
    }
    if (C > n) { /* quirk Q7: SurfaceAreaCosts reads cfg[i] for every clearance i */
        snprintf(g_err, sizeof g_err, "nClearances (%d) > nObjs (%d): the reference reads cfg[i] for clearance i", C, n);
        return -1;
    }
    if ((C > 0 && !clearances) || (R > 0 && (!rss || !rsa))) {
        set_err("", "null argument", 0);
        return -1;
    }
    const int nverts = 4 * C + 4 * n;
    for (int i = 0; i < n; i++)
        if (check_rect(&offlimits[i], nverts, "offlimits", i)) return -1;
    for (int i = 0; i < C; i++) {
        if (check_rect(&clearances[i], nverts, "clearances", i)) return -1;
        if (clearances[i].SourceIndex < 0 || clearances[i].SourceIndex >= n) {
            snprintf(g_err, sizeof g_err, "clearances[%d].SourceIndex = %d out of range", i, clearances[i].SourceIndex);
            return -1;
        }
    }
    for (int i = 0; i < R; i++) {
        if (rss[i].SourceIndex < 0 || rss[i].SourceIndex >= n || rss[i].TargetIndex < 0 || rss[i].TargetIndex >= n ||
            rsa[i].SourceIndex < 0 || rsa[i].SourceIndex >= n || rsa[i].TargetIndex < 0 || rsa[i].TargetIndex >= n) {
            snprintf(g_err, sizeof g_err, "relationship %d: object index out of range", i);
            return -1;
        }
    }

    const int hw = align4((int)(sizeof(mhProblemHeader) / 4));
    int w = hw;
    mhProblemHeader H;
    memset(&H, 0, sizeof H);
    H.off_obj_box = w;    w += 4 * n;
    H.off_clr_box = w;    w += 4 * C;
    H.off_rel_rng = w;    w += 4 * R;
    H.off_rel_aux = w;    w += 4 * R;
    H.off_rel_idx = w;    w += 4 * R;
    H.off_obj_v0x = w;    w = align4(w + n);
    H.off_obj_area = w;   w = align4(w + n);
    H.off_obj_frozen = w; w = align4(w + n);
    H.off_clr_v0x = w;    w = align4(w + C);
    H.off_clr_src = w;    w = align4(w + C);
    H.off_clr_adj_off = w; w = align4(w + n + 1);
    H.off_clr_adj = w;    w = align4(w + C);
    H.off_rel_adj_off = w; w = align4(w + n + 1);
    H.off_rel_adj = w;    w = align4(w + 4 * R);
    H.off_obj_boxq = w;   w += 4 * n;
    H.off_clr_boxq = w;   w += 4 * C;
    H.off_obj_v0xq = w;   w = align4(w + n);
    H.off_clr_v0xq = w;   w = align4(w + C);
    H.smem_words = w;
    H.off_cfg0 = w;       w = align4(w + 3 * n);
    H.off_pass = w;       w = align4(w + 3 * n);
    H.total_words = w;

    float *b = (float *)calloc((size_t)w, sizeof(float));
    if (!b) {
        set_err("", "out of host memory", 0);
        return -1;
    }
    int32_t *bi = (int32_t *)b;

    H.n = n; H.C = C; H.R = R;
    H.w_focal = srf->WeightFocalPoint; H.w_pair = srf->WeightPairWise; H.w_visual = srf->WeightVisualBalance;
    H.w_sym = srf->WeightSymmetry; H.w_off = srf->WeightOffLimits; H.w_clear = srf->WeightClearance;
    H.w_surf = srf->WeightSurfaceArea;
    H.focal_x = (float)srf->focalX; H.focal_y = (float)srf->focalY;
    H.ux = (float)cos(srf->focalRot); H.uy = (float)sin(srf->focalRot);          /* Kernel.cu:290-291 */
    H.fdotu = (float)(srf->focalX * H.ux + srf->focalY * H.uy);                  /* Kernel.cu:292 */
    H.two_focal_rot = (float)(2 * srf->focalRot);                               /* Kernel.cu:297 */
    H.cx2 = (float)(srf->centroidX / 2); H.cy2 = (float)(srf->centroidY / 2);   /* Kernel.cu:206, Q11 */
    {   /* room AABB: minValue/maxValue(surfaceRectangle, 0, 0, 0), Kernel.cu:585-591 */
        double x0 = surfaceRectangle[0].x, x1 = x0, y0 = surfaceRectangle[0].y, y1 = y0;
        for (int k = 1; k < 4; k++) {
            x0 = fmin(x0, surfaceRectangle[k].x); x1 = fmax(x1, surfaceRectangle[k].x);
            y0 = fmin(y0, surfaceRectangle[k].y); y1 = fmax(y1, surfaceRectangle[k].y);
        }
        H.room_minx = (float)x0; H.room_miny = (float)y0; H.room_maxx = (float)x1; H.room_maxy = (float)y1;
        float width = (float)(x1 - x0), height = (float)(y1 - y0);
        H.std_x = width / 16; H.std_y = height / 16;                            /* Q19 */
    }
    H.sigma_t = (float)MH_S_SIGMA_T;
    H.pi_cmp = below_or_equal(MH_PI);
    H.two_pi = (float)(2 * MH_PI);
    H.two_pi_cmp = below_or_equal(2 * MH_PI);
    H.half_pi = (float)(MH_PI / 2.0);

    float denom = 0.f;
    int any_free = 0;
    for (int i = 0; i < n; i++) {
        rect_consts(vertices, offlimits[i].point1Index, b + H.off_obj_box + 4 * i, b + H.off_obj_v0x + i);
        float area = (float)(cfg[i].length * cfg[i].width);                     /* Kernel.cu:199 */
        b[H.off_obj_area + i] = area;
        denom += area;
        bi[H.off_obj_frozen + i] = cfg[i].frozen ? 1 : 0;
        any_free |= !cfg[i].frozen;
        b[H.off_cfg0 + i] = (float)cfg[i].x;
        b[H.off_cfg0 + n + i] = (float)cfg[i].y;
        b[H.off_cfg0 + 2 * n + i] = (float)cfg[i].rotY;
        b[H.off_pass + i] = (float)cfg[i].z;
        b[H.off_pass + n + i] = (float)cfg[i].rotX;
        b[H.off_pass + 2 * n + i] = (float)cfg[i].rotZ;
    }
    H.denom = denom;
    H.inv_denom = (float)(1.0 / (double)denom);
    H.any_free = any_free;
    {   /* fixed-point clearance term: scale from the bounds of positions and rectangles, then the integer constants */
        double pos_bound = fmax(fmax(fabs((double)H.room_minx), fabs((double)H.room_maxx)), fmax(fabs((double)H.room_miny), fabs((double)H.room_maxy)));
        for (int i = 0; i < n; i++) pos_bound = fmax(pos_bound, fmax(fabs(cfg[i].x), fabs(cfg[i].y)));
        for (long long i = 0; extra_layouts && i < (long long)n_extra * n; i++)
            pos_bound = fmax(pos_bound, fmax(fabs(extra_layouts[i].x), fabs(extra_layouts[i].y)));
        double rect_bound = 0, rect_height = 0;
        for (int r = 0; r < n + C; r++) {
            const int start = r < n ? offlimits[r].point1Index : clearances[r - n].point1Index;
            double y0 = vertices[start].y, y1 = y0;
            for (int q = 0; q < 4; q++) {
                rect_bound = fmax(rect_bound, fmax(fabs(vertices[start + q].x), fabs(vertices[start + q].y)));
                y0 = fmin(y0, vertices[start + q].y); y1 = fmax(y1, vertices[start + q].y);
            }
            rect_height = fmax(rect_height, y1 - y0);
        }
        if (!(pos_bound < 1e30) || !(rect_bound < 1e30)) { free(b); set_err("", "non-finite coordinates", 0); return -1; }
        H.clr_k = clearance_scale(pos_bound, rect_bound, rect_height, n, C);
        H.clr_scale = (float)ldexp(1.0, H.clr_k);
        H.clr_unit = (float)ldexp(1.0, -2 * H.clr_k);
        H.clr_pos_limit = (float)(pos_bound + 1.0);
    }
    for (int i = 0; i < n; i++) {
        const float *fb = b + H.off_obj_box + 4 * i;
        for (int q = 0; q < 4; q++) bi[H.off_obj_boxq + 4 * i + q] = to_fixed((double)fb[q], (double)H.clr_scale);
        bi[H.off_obj_v0xq + i] = to_fixed((double)b[H.off_obj_v0x + i], (double)H.clr_scale);
    }
    for (int i = 0; i < C; i++) {
        rect_consts(vertices, clearances[i].point1Index, b + H.off_clr_box + 4 * i, b + H.off_clr_v0x + i);
        for (int q = 0; q < 4; q++) bi[H.off_clr_boxq + 4 * i + q] = to_fixed((double)b[H.off_clr_box + 4 * i + q], (double)H.clr_scale);
        bi[H.off_clr_v0xq + i] = to_fixed((double)b[H.off_clr_v0x + i], (double)H.clr_scale);
        bi[H.off_clr_src + i] = clearances[i].SourceIndex;
        bi[H.off_clr_adj_off + clearances[i].SourceIndex + 1]++;
    }
    for (int i = 0; i < n; i++) /* CSR: clearances by source object (delta evaluation) */
        bi[H.off_clr_adj_off + i + 1] += bi[H.off_clr_adj_off + i];
    {
        int *fill = (int *)calloc((size_t)n, sizeof(int));
        if (!fill) { free(b); set_err("", "out of host memory", 0); return -1; }
        for (int i = 0; i < C; i++) {
            const int s = clearances[i].SourceIndex;
            bi[H.off_clr_adj + bi[H.off_clr_adj_off + s] + fill[s]++] = i;
        }
        free(fill);
    }
    for (int i = 0; i < R; i++) {
        /* Kernel.cu:216, 243: the r-th distance relation uses rss[r]'s pair, the r-th angle
         * relation rsa[r]'s pair (callers make them equal, but nothing requires it) */
        bi[H.off_rel_idx + 4 * i] = rss[i].SourceIndex;
        bi[H.off_rel_idx + 4 * i + 1] = rss[i].TargetIndex;
        bi[H.off_rel_idx + 4 * i + 2] = rsa[i].SourceIndex;
        bi[H.off_rel_idx + 4 * i + 3] = rsa[i].TargetIndex;
        const double start = rss[i].TargetRange.targetRangeStart, end = rss[i].TargetRange.targetRangeEnd;
        const double amin = rsa[i].angleMin, amax = rsa[i].angleMax;
        const int wraps = amin > amax;                                          /* Kernel.cu:245 */
        const double norm = wraps ? (2 * MH_PI - (amax + (2 * MH_PI - amin))) / 2.0 /* Kernel.cu:246 */
                                  : (2 * MH_PI - (amax - amin)) / 2.0;              /* Kernel.cu:252 */
        float *rng = b + H.off_rel_rng + 4 * i, *aux = b + H.off_rel_aux + 4 * i;
        rng[0] = (float)(1.0 / start); rng[1] = (float)end; rng[2] = (float)amin; rng[3] = (float)amax;
        aux[0] = (float)start; aux[1] = (float)(1.0 / norm); aux[2] = wraps ? 1.f : 0.f; aux[3] = 0.f;
    }
    {   /* CSR: relationships by object, each relationship at most once per object (delta evaluation) */
        int *cnt = (int *)calloc((size_t)n + 1, sizeof(int));
        if (!cnt) { free(b); set_err("", "out of host memory", 0); return -1; }
        for (int pass = 0; pass < 2; pass++) {
            for (int i = 0; i < R; i++) {
                const int32_t *id = bi + H.off_rel_idx + 4 * i;
                for (int q = 0; q < 4; q++) {
                    int seen = 0;
                    for (int p = 0; p < q; p++) seen |= id[p] == id[q];
                    if (seen) continue;
                    if (pass == 0) bi[H.off_rel_adj_off + id[q] + 1]++;
                    else bi[H.off_rel_adj + bi[H.off_rel_adj_off + id[q]] + cnt[id[q]]++] = i;
                }
            }
            if (pass == 0)
                for (int i = 0; i < n; i++) bi[H.off_rel_adj_off + i + 1] += bi[H.off_rel_adj_off + i];
        }
        free(cnt);
    }
    memcpy(b, &H, sizeof H);
    out->blob = b;
    out->h = (mhProblemHeader *)b;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Context
 * --------------------------------------------------------------------------------------------- */

typedef struct evPair { void *e0, *e1; } evPair;

struct mhContext {
    /* multi-device parent (mhOptions.n_devices > 1): owns one ordinary context per device and no device memory of
     * its own; shard s holds the chains [shard_first[s], shard_first[s] + shards[s]->n_chains) of this context */
    int n_shards;
    struct mhContext **shards;
    int *shard_first;
    int device;        /* device the context lives on */
    int n, C, R, n_chains, lanes, score_lanes, delta_warps;
    int eval_internal; /* what the chain kernel runs: 0 full scan, 1 delta, 2 exact symmetry memo */
    mhOptions opt;
    int problem_words, smem_words;
    mhProblemHeader hdr; /* host copy of the blob's header: travels in every launch descriptor (constant bank) */
    void *d_problem;
    float *d_x, *d_y, *d_rot, *d_cur, *d_best, *d_beta, *d_beta_snap;
    uint16_t *d_perm;
    void *d_points, *d_costs, *d_scratch, *d_exch_stats;
    void *h_scratch;   /* a 64-byte pinned slot of the library's arena: small read-backs that must not block the host */
    float *ladder;     /* tempering: the current ladder, rung 0 first (host copy; chains hold a permutation of it) */
    /* one-shot calls with a large result block: the chains run as n_chunks launches, and chunk j's scoring and D2H
     * (on copy_stream) overlap the kernels of the chunks after it */
    void *copy_stream;
    void *chunk_ev[MH_MAX_CHUNKS];
    int n_chunks, chunk_first[MH_MAX_CHUNKS + 1];
    void *stream;
    int own_stream;
    uint64_t it_done;  /* iterations already run (relative to opt.iteration_offset) */
    int fresh, costs_dirty;
    evPair *ev; int n_ev, cap_ev;
    double kernel_ms; long long launches;
};

/* mhOptions.device -> CUDA ordinal or -1 = the caller's current device (include/mh_kernel.h) */
static int resolve_device(const mhOptions *o)
{
    if (o->device > 0) return o->device;
    if (o->device == 0 && (o->flags & MH_OPT_EXPLICIT_DEVICE)) return 0;
    return -1;
}

static int enter_device(int want, int *prev)
{
    int e = mhdev_get_device(prev);
    if (e) return e;
    if (want >= 0 && want != *prev) return mhdev_set_device(want);
    return 0;
}
static void leave_device(int want, int prev)
{
    if (want >= 0 && want != prev) mhdev_set_device(prev);
}

/* Lanes per chain.  A warp advances 32/G chains per proposal step at a cost of roughly
 * O + rows*(12n + 8C + 100) warp instructions, rows = ceil(n/G) and O ~ 450 for the work every
 * lane repeats (Philox, propose, accept, reductions), so the cost per chain is
 * G*(O + rows*(12n+8C+100))/32: narrow groups win on small rooms, and lose nothing on big ones
 * except shared memory -- which caps the resident warps, and throughput grows about linearly
 * with resident warps up to ~18 per SM (measured on B200).  Pick the width
 * with the best modelled throughput; MH_LANES or mhOptions.lanes_per_chain override. */
static int choose_lanes(int n, int C, int R, int smem_words, int n_chains, int requested, int eval_mode)
{
    int max_block = 0, max_sm = 0, sms = 0;
    if (mhdev_device_limits(&max_block, &max_sm, &sms, NULL, NULL, NULL, NULL, 0)) {
        set_err("%s failed: %s", "device query", mhdev_device_limits(&max_block, &max_sm, &sms, NULL, NULL, NULL, NULL, 0));
        return -1;
    }
    const char *env = getenv("MH_LANES");
    if (requested <= 0 && env) requested = atoi(env);
    static const int cand[6] = { 32, 16, 8, 4, 2, 1 };
    int best = -1;
    double best_score = -1.0;
    for (int k = 0; k < 6; k++) {
        const int G = cand[k];
        const int bytes = mhdev_chain_smem_bytes(smem_words, n, C, R, G, eval_mode, 4);
        if (bytes < 0 || bytes > max_block) continue;
        if (requested == G) return G;
        /* one lane per chain: the 32 lanes of a warp load 32 different float4s per column (4 shared-memory
         * wavefronts instead of a broadcast), which binds from ~20 objects (measured at n=24: 1.13e9 against 1.52e9) */
        if (G == 1 && n >= 20 && best > 0) continue;
        int blocks_per_sm = max_sm / (bytes + 1024);
        if (blocks_per_sm > MH_MAX_BLOCKS_PER_SM) blocks_per_sm = MH_MAX_BLOCKS_PER_SM; /* register-limited */
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        const double cap_warps = 4.0 * blocks_per_sm;
        const double cpw = 32.0 / G;
        const double warps_total = ceil((double)n_chains / cpw);
        double per_sm = warps_total / (double)sms;
        if (per_sm > cap_warps) per_sm = cap_warps;
        /* latency is covered from ~12 resident warps per SM (measured: n=8 runs best with 1 lane per chain at 14
         * warps/SM, n=16 with 16384 chains best with 4 lanes at 14 warps/SM, not with 8 at 28) */
        const double cover = per_sm >= 12.0 ? 1.0 : per_sm / 12.0;
        const double rows = ceil((double)n / G);
        /* warp instructions of the whole job: with chains to spare this is proportional to G (the old form of
         * the model); with fewer chains than a warp holds it is not -- one chain is one warp for every G, and
         * the widest group, i.e. the shortest per-lane program, wins (measured: 13.7 -> 3.0 ms for 1 chain x 1000 iterations at n=16) */
        /* (the work every lane repeats: ~150 instructions per iteration, plus ~300 of drawing the proposal recipe, which
         * the lanes of a group share -- one iteration per lane, mh_kernels.cu) */
        const double cost = warps_total * (150.0 + 300.0 / G + rows * (12.0 * n + 8.0 * C + 100.0));
        const double score = cover / cost;
        if (score > best_score * 1.03) { /* near-ties go to the wider group (less shared memory) */
            best_score = score;
            best = G;
        }
    }
    if (best < 0) snprintf(g_err, sizeof g_err, "problem does not fit in shared memory (n=%d, C=%d)", n, C);
    return best;
}

/* Launch shape of the delta kernel (mh_delta_kernel): lanes per chain and warps per block.
 * Its per-proposal work is ~(fixed scalar part) + (3n + 2C)/G pair evaluations per lane, and the fixed
 * part is repeated by every lane of a group, so narrow groups win as long as a lane keeps 6-8 rows
 * (measured on B200: n=8 -> 1, n=16 -> 2, n=50 -> 8, n=200 -> 32 lanes).  Few chains: widen the groups
 * until the machine is covered.  Blocks of 8 warps when shared memory allows and the grid still covers
 * the SMs (the problem blob is staged once per block: 8 >= 4 at n = 50, +20 % at n = 200). */
static int choose_delta_shape(int n, int C, int R, int smem_words, int job_chains, int n_chains, int requested, int eval_mode, int *lanes_out, int *warps_out)
{
    int max_block = 0, max_sm = 0, sms = 0;
    int e = mhdev_device_limits(&max_block, &max_sm, &sms, NULL, NULL, NULL, NULL, 0);
    if (e) { set_err("%s failed: %s", "device query", e); return -1; }
    const char *env = getenv("MH_LANES");
    if (requested <= 0 && env) requested = atoi(env);
    int G = 1;
    while (G < 32 && G * 8 < n) G *= 2;
    while (G < 32 && (double)job_chains * G / 32.0 < 8.0 * sms) G *= 2;    /* under-filled machine: wider groups (by the JOB's size: every shard must agree) */
    if (requested > 0) G = requested;
    for (;; G *= 2) {                                                       /* must fit with 4 warps per block */
        if (G > 32) { snprintf(g_err, sizeof g_err, "problem does not fit in shared memory (n=%d, C=%d, delta evaluation)", n, C); return -1; }
        const int bytes = mhdev_chain_smem_bytes(smem_words, n, C, R, G, eval_mode, 4);
        if (bytes >= 0 && bytes <= max_block) break;
        if (requested > 0) { snprintf(g_err, sizeof g_err, "lanes_per_chain=%d does not fit in shared memory (delta evaluation)", G); return -1; }
    }
    /* warps per block: whichever of 8, 6 and 4 keeps more warps resident -- shared memory, and registers: the kernel is
     * built twice, for blocks of up to 8 warps at 128 registers (16 warps per SM) and for blocks of 4 warps at 96
     * registers (20 warps per SM) -- the largest on a tie (the problem blob is staged once per block), and never a
     * grid smaller than the SM count when a smaller block would cover it.  Measured on B200: 50 objects, 8 lanes per
     * chain: 5 x 4 warps 1.12e9 proposals/s against 1.06e9 for 2 x 8; 200 objects, 32 lanes: 2 x 8 wins by 19 %. */
    int warps = 4, best_res = -1;
    const char *wenv = getenv("MH_DELTA_WARPS");
    const double total_warps = ceil((double)n_chains * G / 32.0);
    for (int w = 8; w >= 4; w -= 2) {
        const int bytes = mhdev_chain_smem_bytes(smem_words, n, C, R, G, eval_mode, w);
        if (bytes < 0 || bytes > max_block) continue;
        if (wenv && atoi(wenv) == w) { warps = w; break; }
        const int reg_warps = w == 4 ? 20 : 16;
        int blocks = max_sm / (bytes + 1024);
        if (blocks * w > reg_warps) blocks = reg_warps / w;
        int res = blocks * w;
        if (w > 4 && total_warps / w < (double)sms) res = 0;
        if (res > best_res) { best_res = res; warps = w; }
    }
    *lanes_out = G;
    *warps_out = warps;
    return 0;
}

static void ctx_free(mhContext *c)
{
    if (!c) return;
    if (c->n_shards || c->shards) {                            /* multi-device parent: the shards are destroyed by the caller */
        free(c->shards);
        free(c->shard_first);
        free(c->ladder);
        free(c);
        return;
    }
    void *st = c->stream;
    mhdev_free(c->d_problem, st); mhdev_free(c->d_x, st); mhdev_free(c->d_y, st); mhdev_free(c->d_rot, st); mhdev_free(c->d_cur, st);
    mhdev_free(c->d_best, st); mhdev_free(c->d_beta, st); mhdev_free(c->d_perm, st); mhdev_free(c->d_points, st);
    mhdev_free(c->d_costs, st); mhdev_free(c->d_scratch, st); mhdev_free(c->d_beta_snap, st); mhdev_free(c->d_exch_stats, st);
    for (int i = 0; i < c->n_ev; i++) { mhdev_event_destroy(c->ev[i].e0); mhdev_event_destroy(c->ev[i].e1); }
    free(c->ev);
    free(c->ladder);
    for (int j = 0; j < c->n_chunks; j++) mhdev_event_destroy(c->chunk_ev[j]);
    if (c->copy_stream) mhdev_stream_destroy(c->copy_stream);
    mhdev_scratch_release(c->h_scratch);
    if (c->own_stream) mhdev_stream_destroy(c->stream);
    free(c);
}

/* The result block must come from plain malloc() (the reference's callers free() it, Kernel.cu:928),
 * so it cannot be pinned memory.  A fresh 80 MB malloc is untouched address space: copying into it
 * page-faults 20 000 times.  Populate it first (Linux >= 5.14; harmless if unsupported) -- while the kernel
 * runs, and on several threads when the block is large: one thread populates about 3 GB/s, the 1.26 GB block of
 * BASELINE config 4 would take longer than the kernel does on 8 GPUs. */
typedef struct prefaultJob { void *p; size_t bytes; } prefaultJob;

static void *prefault_worker(void *arg)
{
#if defined(__linux__) && defined(MADV_POPULATE_WRITE)
    const prefaultJob *j = (const prefaultJob *)arg;
    const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
    const uintptr_t a = ((uintptr_t)j->p + page - 1) & ~(page - 1), e = ((uintptr_t)j->p + j->bytes) & ~(page - 1);
    if (e > a) (void)madvise((void *)a, (size_t)(e - a), MADV_POPULATE_WRITE);
#else
    (void)arg;
#endif
    return NULL;
}

static void prefault(void *p, size_t bytes)
{
    if (bytes < (1u << 20)) return;
    int threads = (int)(bytes >> 26);                           /* one thread per 64 MB ... */
    if (threads > 8) threads = 8;                               /* ... at most 8 */
    if (threads < 1) threads = 1;
    prefaultJob jobs[8];
    pthread_t th[8];
    const size_t slice = ((bytes / (size_t)threads) + 4095) & ~(size_t)4095;
    int started = 0;
    for (int i = 0; i < threads; i++) {
        const size_t off = slice * (size_t)i;
        if (off >= bytes) { threads = i; break; }
        jobs[i].p = (char *)p + off;
        jobs[i].bytes = off + slice <= bytes ? slice : bytes - off;
    }
    for (int i = 1; i < threads; i++, started++)
        if (pthread_create(&th[i], NULL, prefault_worker, &jobs[i])) { prefault_worker(&jobs[i]); th[i] = 0; }
    prefault_worker(&jobs[0]);
    for (int i = 1; i < threads; i++)
        if (th[i]) pthread_join(th[i], NULL);
    (void)started;
}

static void default_options(mhOptions *o)
{
    memset(o, 0, sizeof *o);
    o->struct_size = (uint32_t)sizeof *o;
    o->device = -1;
}

static void take_options(mhOptions *dst, const mhOptions *src)
{
    default_options(dst);
    if (src) {
        size_t sz = src->struct_size ? src->struct_size : sizeof *src;
        if (sz > sizeof *dst) sz = sizeof *dst;
        memcpy(dst, src, sz);
        dst->struct_size = (uint32_t)sizeof *dst;
    }
    if (dst->chain_stride == 0) dst->chain_stride = 1;
    if (dst->beta_start <= 0) dst->beta_start = MH_BETA;
    if (dst->beta_end <= 0) dst->beta_end = dst->beta_start;
}

/* Env MH_DEVICES = "all" | "0,1,5": the device list for callers whose options carry none (the reference's own
 * entry point has no options at all).  Returns the number of devices written (0 = not set), -1 on a bad list. */
static int devices_from_env(int32_t devices[MH_MAX_DEVICES])
{
    const char *env = getenv("MH_DEVICES");
    if (!env || !*env) return 0;
    int count = 0;
    if (mhdev_device_count(&count) || count < 1) return 0;
    int k = 0;
    if (!strcmp(env, "all")) {
        for (; k < count && k < MH_MAX_DEVICES; k++) devices[k] = k;
        return k;
    }
    const char *q = env;
    while (*q) {
        char *end = NULL;
        const long v = strtol(q, &end, 10);
        if (end == q || v < 0 || v >= count || k == MH_MAX_DEVICES) {
            snprintf(g_err, sizeof g_err, "MH_DEVICES=\"%s\": expected \"all\" or up to %d comma-separated ordinals below %d", env,
                     MH_MAX_DEVICES, count);
            return -1;
        }
        devices[k++] = (int32_t)v;
        q = *end == ',' ? end + 1 : end;
        if (*end && *end != ',') { snprintf(g_err, sizeof g_err, "MH_DEVICES=\"%s\": bad separator", env); return -1; }
    }
    return k;
}

/* rung r of every ladder starts at ladder[r]; the ladder itself starts geometric:
 * beta_start * (beta_end/beta_start)^(r/(T-1)) */
static int init_betas(mhContext *c)
{
    const int T = c->opt.tempering_rungs;
    float *hb = (float *)malloc(4 * (size_t)c->n_chains);
    if (!hb) return 2; /* cudaErrorMemoryAllocation */
    if (!c->ladder) {
        c->ladder = (float *)malloc(4 * (size_t)T);
        if (!c->ladder) { free(hb); return 2; }
        const float lr = log2f((float)(c->opt.beta_end / c->opt.beta_start));
        for (int r = 0; r < T; r++) {
            const float t = T > 1 ? (float)r / (float)(T - 1) : 0.f;
            c->ladder[r] = (float)c->opt.beta_start * exp2f(t * lr);
        }
    }
    for (int i = 0; i < c->n_chains; i++) {
        const int r = (int)((c->opt.chain_offset + (uint64_t)i * c->opt.chain_stride) % (uint64_t)T);
        hb[i] = c->ladder[r];
    }
    int e = mhdev_h2d(c->d_beta, hb, 4 * (size_t)c->n_chains, c->stream);
    if (!e) e = mhdev_stream_sync(c->stream);
    free(hb);
    return e;
}

/* One context on one device.  `o` has been through take_options; P stays the caller's. */
static mhContext *create_single(const mhProblem *P, int nChains, const mhOptions *o)
{
    mhContext *c = NULL;
    int prev = -1, entered = 0;
    const int want = resolve_device(o);
    c = (mhContext *)calloc(1, sizeof *c);
    if (!c) { set_err("", "out of host memory", 0); goto fail; }
    c->opt = *o;
    if (c->opt.tempering_rungs > 1) {
        if (c->opt.chain_stride == 1 &&
            (nChains % c->opt.tempering_rungs || c->opt.chain_offset % (uint64_t)c->opt.tempering_rungs)) {
            set_err("", "tempering: chain_offset and nChains must be multiples of tempering_rungs", 0);
            goto fail;
        }
        if (c->opt.chain_stride > 1 && (c->opt.chain_offset >= c->opt.chain_stride ||
                                         ((uint64_t)nChains * c->opt.chain_stride) % (uint64_t)c->opt.tempering_rungs)) {
            set_err("", "tempering: strided shards need chain_offset < chain_stride and whole ladders over all ranks", 0);
            goto fail;
        }
        if (c->opt.exchange_interval <= 0) c->opt.exchange_interval = 100;
    }
    CU(enter_device(want, &prev));
    entered = 1;
    c->device = want < 0 ? prev : want;
    c->n = P->h->n; c->C = P->h->C; c->R = P->h->R; c->n_chains = nChains;
    c->problem_words = P->h->total_words; c->smem_words = P->h->smem_words;
    c->hdr = *P->h;
    if (c->opt.eval_mode < MH_EVAL_FULL || c->opt.eval_mode > MH_EVAL_FULL_SCAN) { set_err("", "unknown eval_mode", 0); goto fail; }
    c->eval_internal = c->opt.eval_mode == MH_EVAL_FULL_SCAN ? 0 : c->opt.eval_mode;
    c->lanes = -1;
    /* the lane width (and with it the float reduction order) follows the size of the WHOLE job, so that every
     * shard of it -- another GPU, another rank, another call -- makes the same choice (mhOptions.total_chains) */
    uint64_t job = c->opt.total_chains ? c->opt.total_chains : (uint64_t)nChains;
    if (job < (uint64_t)nChains) job = (uint64_t)nChains;
    const int job_chains = job > 0x7fffffffull ? 0x7fffffff : (int)job;
    const int memo_min = job_chains >= MH_BIG_JOB_CHAINS ? MH_MEMO_MIN_OBJS_BIG_JOB : MH_MEMO_MIN_OBJS;
    if (c->opt.eval_mode == MH_EVAL_FULL && c->n >= memo_min && !getenv("MH_FULL_SCAN")) {
        if (choose_delta_shape(c->n, c->C, c->R, c->smem_words, job_chains, nChains, c->opt.lanes_per_chain, MH_EVAL_MEMO, &c->lanes, &c->delta_warps) == 0)
            c->eval_internal = MH_EVAL_MEMO;
        else
            c->lanes = -1;                                   /* does not fit: the plain scan needs less shared memory */
        g_err[0] = 0;
    } else if (c->eval_internal == MH_EVAL_DELTA || c->eval_internal == MH_EVAL_MEMO) {
        if (choose_delta_shape(c->n, c->C, c->R, c->smem_words, job_chains, nChains, c->opt.lanes_per_chain, c->eval_internal, &c->lanes, &c->delta_warps)) goto fail;
    }
    if (c->lanes < 0) {
        c->lanes = choose_lanes(c->n, c->C, c->R, c->smem_words, job_chains, c->opt.lanes_per_chain, c->eval_internal);
    }
    if (c->lanes < 0) goto fail;
    /* a pinned lane width also pins the scoring kernel, so that sharded runs report identical bits */
    c->score_lanes = choose_lanes(c->n, c->C, c->R, c->smem_words, job_chains, c->opt.lanes_per_chain, MH_EVAL_FULL);
    if (c->score_lanes < 0) goto fail;
    c->fresh = 1;
    CU(mhdev_stream_create(&c->stream));
    c->own_stream = 1;
    const size_t cn = (size_t)nChains * (size_t)c->n;
    CU(mhdev_malloc(&c->d_problem, 4 * (size_t)c->problem_words, c->stream));
    CU(mhdev_malloc((void **)&c->d_x, 4 * cn, c->stream));
    CU(mhdev_malloc((void **)&c->d_y, 4 * cn, c->stream));
    CU(mhdev_malloc((void **)&c->d_rot, 4 * cn, c->stream));
    CU(mhdev_malloc((void **)&c->d_perm, 2 * cn, c->stream));
    CU(mhdev_malloc((void **)&c->d_cur, 4 * (size_t)nChains, c->stream));
    CU(mhdev_malloc((void **)&c->d_best, 4 * (size_t)nChains, c->stream));
    CU(mhdev_malloc((void **)&c->d_beta, 4 * (size_t)nChains, c->stream));
    CU(mhdev_malloc((void **)&c->d_beta_snap, 4 * (size_t)nChains, c->stream));
    if (c->opt.tempering_rungs > 1) {
        CU(mhdev_malloc(&c->d_exch_stats, 16 * (size_t)c->opt.tempering_rungs, c->stream));
        CU(mhdev_memset(c->d_exch_stats, 0, 16 * (size_t)c->opt.tempering_rungs, c->stream));
    }
    CU(mhdev_malloc(&c->d_points, sizeof(point) * cn, c->stream));
    CU(mhdev_malloc(&c->d_costs, sizeof(resultCosts) * (size_t)nChains, c->stream));
    CU(mhdev_malloc(&c->d_scratch, 64, c->stream));
    CU(mhdev_scratch_acquire(&c->h_scratch));
    CU(mhdev_h2d(c->d_problem, P->blob, 4 * (size_t)c->problem_words, c->stream));
    if (c->opt.tempering_rungs > 1) CU(init_betas(c));
    CU(mhdev_stream_sync(c->stream)); /* the caller frees the blob */
    leave_device(want, prev);
    return c;
fail:
    if (c) ctx_free(c);
    if (entered) leave_device(want, prev);
    return NULL;
}

static void destroy_single(mhContext *ctx)
{
    int prev = -1;
    const int dev = ctx->device;
    int e = enter_device(dev, &prev);
    if (!e) {
        if (ctx->copy_stream) mhdev_stream_sync(ctx->copy_stream);
        mhdev_stream_sync(ctx->stream);
    }
    ctx_free(ctx);
    if (prev >= 0) leave_device(dev, prev);
}

/* mhOptions.n_devices > 1: contiguous ranges of the chains (whole ladders when tempering), one context per
 * device.  Devices that would get no chain are left out. */
static mhContext *create_multi(const mhProblem *P, int nChains, const mhOptions *o)
{
    mhContext *c = NULL;
    const int D = o->n_devices;
    if (o->chain_stride != 1) { set_err("", "n_devices > 1 needs chain_stride = 1 (strided shards are the one-process-per-GPU layout)", 0); return NULL; }
    const int unit = o->tempering_rungs > 1 ? o->tempering_rungs : 1;   /* shards hold whole ladders */
    if (nChains % unit || o->chain_offset % (uint64_t)unit) {
        set_err("", "tempering: chain_offset and nChains must be multiples of tempering_rungs", 0);
        return NULL;
    }
    for (int i = 0; i < D; i++)
        if (o->devices[i] < 0) { snprintf(g_err, sizeof g_err, "devices[%d] = %d", i, o->devices[i]); return NULL; }
    c = (mhContext *)calloc(1, sizeof *c);
    if (c) {
        c->shards = (mhContext **)calloc((size_t)D, sizeof *c->shards);
        c->shard_first = (int *)calloc((size_t)D + 1, sizeof *c->shard_first);
    }
    if (!c || !c->shards || !c->shard_first) { set_err("", "out of host memory", 0); goto fail; }
    c->opt = *o;
    c->device = -1;
    c->n = P->h->n; c->C = P->h->C; c->R = P->h->R; c->n_chains = nChains;
    const int units = nChains / unit, base = units / D, rem = units % D;
    int first = 0;
    for (int i = 0; i < D; i++) {
        const int count = (base + (i < rem ? 1 : 0)) * unit;
        if (count == 0) continue;
        mhOptions so = *o;
        so.n_devices = 0;
        so.device = o->devices[i];
        so.flags |= MH_OPT_EXPLICIT_DEVICE;
        so.chain_offset = o->chain_offset + (uint64_t)first;
        so.total_chains = o->total_chains ? o->total_chains : (uint64_t)nChains;
        mhContext *sc = create_single(P, count, &so);
        if (!sc) goto fail;
        c->shards[c->n_shards] = sc;
        c->shard_first[c->n_shards] = first;
        c->n_shards++;
        first += count;
    }
    c->shard_first[c->n_shards] = first;
    c->lanes = c->shards[0]->lanes; c->score_lanes = c->shards[0]->score_lanes; c->eval_internal = c->shards[0]->eval_internal;
    return c;
fail:
    if (c) {
        char keep[sizeof g_err];
        memcpy(keep, g_err, sizeof keep);
        for (int i = 0; i < c->n_shards; i++) destroy_single(c->shards[i]);
        memcpy(g_err, keep, sizeof keep);
        c->n_shards = 0;
        if (!c->shards) c->shards = (mhContext **)calloc(1, sizeof *c->shards);   /* marks the parent for ctx_free */
        ctx_free(c);
    }
    return NULL;
}

MH_API mhContext *KernelCreate(const relationshipStruct *rss, const relationshipAngleStruct *rsa,
                               const positionAndRotation *cfg, const rectangle *clearances, const rectangle *offlimits,
                               const vertex *vertices, const vertex *surfaceRectangle, const Surface *srf, int nChains,
                               const mhOptions *opt)
{
    g_err[0] = 0;
    mhProblem P;
    mhOptions o;
    if (nChains < 1) {
        set_err("", "nChains must be >= 1", 0);
        return NULL;
    }
    take_options(&o, opt);
    if (o.n_devices < 0 || o.n_devices > MH_MAX_DEVICES) {
        snprintf(g_err, sizeof g_err, "n_devices = %d: at most %d devices", o.n_devices, MH_MAX_DEVICES);
        return NULL;
    }
    if (o.n_devices == 0 && resolve_device(&o) < 0 && o.chain_stride == 1) {
        const int k = devices_from_env(o.devices);
        if (k < 0) return NULL;
        if (k > 1) o.n_devices = k;
        else if (k == 1) { o.device = o.devices[0]; o.flags |= MH_OPT_EXPLICIT_DEVICE; }
    }
    if (o.n_devices == 1) { o.device = o.devices[0]; o.flags |= MH_OPT_EXPLICIT_DEVICE; o.n_devices = 0; }
    if (pack_problem(rss, rsa, cfg, clearances, offlimits, vertices, surfaceRectangle, srf, NULL, 0, &P)) return NULL;
    mhContext *c = o.n_devices > 1 ? create_multi(&P, nChains, &o) : create_single(&P, nChains, &o);
    free(P.blob);
    return c;
}

static int push_events(mhContext *c, void **e0, void **e1)
{
    if (c->n_ev == c->cap_ev) {
        int cap = c->cap_ev ? 2 * c->cap_ev : 16;
        evPair *p = (evPair *)realloc(c->ev, sizeof(evPair) * (size_t)cap);
        if (!p) return -1;
        c->ev = p;
        c->cap_ev = cap;
    }
    int e = mhdev_event_create(e0);
    if (e) return e;
    e = mhdev_event_create(e1);
    if (e) {
        mhdev_event_destroy(*e0);
        *e0 = NULL;
        return e;
    }
    c->ev[c->n_ev].e0 = *e0;
    c->ev[c->n_ev].e1 = *e1;
    c->n_ev++;
    return 0;
}

/* Take back the pair push_events handed out last: the launch it was to bracket did not happen, and an event
 * that was never recorded must not reach cudaEventElapsedTime (it would turn one failure into two). */
static void pop_events(mhContext *c)
{
    if (c->n_ev == 0) return;
    c->n_ev--;
    mhdev_event_destroy(c->ev[c->n_ev].e0);
    mhdev_event_destroy(c->ev[c->n_ev].e1);
}

static int drain_events(mhContext *c)
{
    for (int i = 0; i < c->n_ev; i++) {
        float ms = 0.f;
        int e = mhdev_event_elapsed_ms(c->ev[i].e0, c->ev[i].e1, &ms);
        if (e) return e;
        c->kernel_ms += ms;
        mhdev_event_destroy(c->ev[i].e0);
        mhdev_event_destroy(c->ev[i].e1);
    }
    c->n_ev = 0;
    return 0;
}

/* One launch of the chain kernel over the chains [first, first + count) of the context (the whole context, or one
 * chunk of a one-shot call).  A chain's result does not depend on which launch, block or lane runs it. */
static int launch_chains_range(mhContext *c, int first, int count, int iterations, void *d_trace)
{
    mhLaunch L;
    const size_t fo = (size_t)first * (size_t)c->n;
    memset(&L, 0, sizeof L);
    L.d_problem = c->d_problem; L.problem_words = c->problem_words; L.smem_words = c->smem_words;
    L.n = c->n; L.C = c->C; L.R = c->R; L.n_chains = count; L.lanes = c->lanes; L.fresh = c->fresh;
    L.seed = c->opt.seed; L.chain_offset = c->opt.chain_offset + (uint64_t)first * c->opt.chain_stride; L.chain_stride = c->opt.chain_stride;
    L.it_begin = c->opt.iteration_offset + c->it_done; L.it_count = iterations;
    L.schedule = c->opt.tempering_rungs > 1 ? MH_SCHED_PER_CHAIN : c->opt.schedule;
    L.schedule_length = c->opt.schedule_length;
    L.result_mode = c->opt.result_mode;
    L.eval_mode = c->eval_internal;
    L.warps_per_block = c->delta_warps;
    L.beta_start = (float)c->opt.beta_start; L.beta_end = (float)c->opt.beta_end;
    L.beta_log2_ratio = log2f((float)(c->opt.beta_end / c->opt.beta_start));
    L.d_x = c->d_x + fo; L.d_y = c->d_y + fo; L.d_rot = c->d_rot + fo; L.d_perm = c->d_perm + fo; L.d_cur_total = c->d_cur + first;
    L.d_best_total = c->d_best + first; L.d_beta = c->d_beta + first;
    L.d_points = (char *)c->d_points + sizeof(point) * fo; L.d_costs = (char *)c->d_costs + sizeof(resultCosts) * (size_t)first;
    L.d_trace = d_trace;
    L.stream = c->stream;
    L.hdr = c->hdr;
    void *e0 = NULL, *e1 = NULL;
    int e = push_events(c, &e0, &e1);
    if (e) return e;
    e = mhdev_event_record(e0, c->stream);
    if (!e) {
        const char *fault = getenv("MH_FAULT");                /* test hook: MH_FAULT=launch fails the next chain launch */
        e = (fault && !strcmp(fault, "launch")) ? 1 /* cudaErrorInvalidValue */ : mhdev_launch_chains(&L);
    }
    if (!e) e = mhdev_event_record(e1, c->stream);
    if (e) {
        pop_events(c);
        return e;
    }
    c->launches++;
    return 0;
}

static int launch_segment(mhContext *c, int iterations, void *d_trace)
{
    int e = launch_chains_range(c, 0, c->n_chains, iterations, d_trace);
    if (e) return e;
    c->fresh = 0;
    c->it_done += (uint64_t)iterations;
    c->costs_dirty = 1;
    if (c->n_ev >= 1024) return drain_events(c);
    return 0;
}

static int run_iterations(mhContext *c, int iterations, mhTraceEntry *trace)
{
    void *d_trace = NULL;
    int prev = -1, rc = -1;
    g_err[0] = 0;
    if (!c || iterations < 0) { set_err("", "bad arguments", 0); return -1; }
    CU(enter_device(c->device, &prev));
    if (c->opt.schedule_length <= 0 && c->opt.schedule != MH_SCHEDULE_CONSTANT) c->opt.schedule_length = iterations;
    if (trace) CU(mhdev_malloc(&d_trace, sizeof(mhTraceEntry) * (size_t)iterations * (size_t)c->n_chains, c->stream));
    if (c->opt.tempering_rungs > 1 && c->opt.chain_stride > 1) {
        /* ladders span several contexts: the caller exchanges (KernelTemperingExchange) */
        const uint64_t ex = (uint64_t)c->opt.exchange_interval;
        const uint64_t git = c->opt.iteration_offset + c->it_done;
        if ((git % ex) + (uint64_t)iterations > ex) {
            set_err("", "tempering with chain_stride > 1: a run may not cross an exchange boundary", 0);
            goto fail;
        }
        if (iterations > 0 || c->fresh) CU(launch_segment(c, iterations, d_trace));
    } else if (c->opt.tempering_rungs > 1) {
        /* segments end on exchange boundaries; the whole ladder lives in this context */
        const uint64_t ex = (uint64_t)c->opt.exchange_interval;
        int left = iterations;
        size_t traced = 0;
        while (left > 0) {
            const uint64_t git = c->opt.iteration_offset + c->it_done;
            uint64_t to_boundary = ex - (git % ex);
            int seg = left < (int)to_boundary ? left : (int)to_boundary;
            CU(launch_segment(c, seg, d_trace ? (char *)d_trace + sizeof(mhTraceEntry) * traced * (size_t)c->n_chains : NULL));
            traced += (size_t)seg;
            left -= seg;
            const uint64_t gnow = c->opt.iteration_offset + c->it_done;
            if (gnow % ex == 0) {
                /* both members of a pair must decide on the same betas: read a snapshot */
                CU(mhdev_d2d(c->d_beta_snap, c->d_beta, 4 * (size_t)c->n_chains, c->stream));
                CU(mhdev_launch_exchange(c->n_chains, c->opt.chain_offset, 1, c->opt.tempering_rungs, gnow / ex, gnow - 1,
                                         c->opt.seed, c->d_cur, c->d_beta_snap, c->opt.chain_offset, 1, (uint64_t)c->n_chains,
                                         c->d_beta, c->d_exch_stats, c->stream));
                c->launches++;
            }
        }
    } else if (iterations > 0 || c->fresh) {
        CU(launch_segment(c, iterations, d_trace));
    }
    if (trace) {
        CU(mhdev_d2h(trace, d_trace, sizeof(mhTraceEntry) * (size_t)iterations * (size_t)c->n_chains, c->stream));
        CU(mhdev_stream_sync(c->stream));
    }
    rc = 0;
fail:
    if (d_trace) { mhdev_stream_sync(c->stream); mhdev_free(d_trace, c->stream); }
    if (prev >= 0) leave_device(c->device, prev);
    return rc;
}

/* ---- multi-device plumbing ----------------------------------------------------------------------- */

#define IS_MULTI(ctx) ((ctx)->n_shards > 0)

static int multi_unsupported(const char *what)
{
    snprintf(g_err, sizeof g_err, "%s is not available on a multi-device context (n_devices > 1): it names ONE device's memory or stream", what);
    return -1;
}

typedef struct shardJob {
    mhContext *c;
    point *points;
    resultCosts *costs;
    int rc;
    char err[sizeof g_err];
} shardJob;

static int results_single(mhContext *ctx, point *points, resultCosts *costs);

static void *results_worker(void *arg)
{
    shardJob *j = (shardJob *)arg;
    j->rc = results_single(j->c, j->points, j->costs);
    if (j->rc) memcpy(j->err, g_err, sizeof j->err);
    return NULL;
}

MH_API int KernelRun(mhContext *ctx, int iterations)
{
    if (ctx && IS_MULTI(ctx)) {                                  /* launches are asynchronous: the devices run side by side */
        for (int i = 0; i < ctx->n_shards; i++)
            if (run_iterations(ctx->shards[i], iterations, NULL)) return -1;
        return 0;
    }
    return run_iterations(ctx, iterations, NULL);
}

MH_API int KernelRunTraced(mhContext *ctx, int iterations, mhTraceEntry *trace)
{
    if (!trace) { set_err("", "trace buffer is NULL", 0); return -1; }
    if (ctx && IS_MULTI(ctx)) {                                  /* a test facility: one device after the other */
        if (iterations < 0) { set_err("", "bad arguments", 0); return -1; }
        for (int i = 0; i < ctx->n_shards; i++) {
            mhContext *sc = ctx->shards[i];
            mhTraceEntry *tmp = (mhTraceEntry *)malloc(sizeof(mhTraceEntry) * (size_t)(iterations ? iterations : 1) * (size_t)sc->n_chains);
            if (!tmp) { set_err("", "out of host memory", 0); return -1; }
            const int rc = run_iterations(sc, iterations, tmp);
            for (int it = 0; !rc && it < iterations; it++)
                memcpy(trace + (size_t)it * (size_t)ctx->n_chains + (size_t)ctx->shard_first[i], tmp + (size_t)it * (size_t)sc->n_chains,
                       sizeof(mhTraceEntry) * (size_t)sc->n_chains);
            free(tmp);
            if (rc) return -1;
        }
        return 0;
    }
    return run_iterations(ctx, iterations, trace);
}

static int ensure_scored(mhContext *c)
{
    if (c->fresh) { /* nothing has run: emit the initial layout */
        int e = launch_segment(c, 0, NULL);
        if (e) return e;
    }
    if (!c->costs_dirty) return 0;
    int e = mhdev_launch_score(c->d_problem, c->smem_words, c->n, c->C, c->R, c->n_chains, c->score_lanes, c->d_points, c->d_costs,
                               c->stream);
    if (e) return e;
    c->launches++;
    c->costs_dirty = 0;
    return 0;
}

MH_API int KernelSynchronize(mhContext *ctx)
{
    int prev = -1, rc = -1;
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (IS_MULTI(ctx)) {
        for (int i = 0; i < ctx->n_shards; i++)
            if (KernelSynchronize(ctx->shards[i])) return -1;
        return 0;
    }
    CU(enter_device(ctx->device, &prev));
    CU(mhdev_stream_sync(ctx->stream));
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

static int results_single(mhContext *ctx, point *points, resultCosts *costs)
{
    int prev = -1, rc = -1;
    CU(enter_device(ctx->device, &prev));
    CU(ensure_scored(ctx));
    if (points) CU(mhdev_d2h(points, ctx->d_points, sizeof(point) * (size_t)ctx->n_chains * (size_t)ctx->n, ctx->stream));
    if (costs) CU(mhdev_d2h(costs, ctx->d_costs, sizeof(resultCosts) * (size_t)ctx->n_chains, ctx->stream));
    CU(mhdev_stream_sync(ctx->stream));
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

/* Every device copies its slice straight into the caller's one block (no collective).
 * A device-to-pageable-host copy blocks the calling thread, so each device gets a thread of its own and
 * the copies share the host's memory bandwidth instead of queueing behind each other. */
static int fetch_all_shards(mhContext *ctx, point *points, resultCosts *costs, void *(*worker)(void *))
{
    const int S = ctx->n_shards;
    shardJob *jobs = (shardJob *)calloc((size_t)S, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)S, sizeof *th);
    if (!jobs || !th) { free(jobs); free(th); set_err("", "out of host memory", 0); return -1; }
    int rc = 0;
    for (int i = 0; i < S; i++) {
        jobs[i].c = ctx->shards[i];
        jobs[i].points = points ? points + (size_t)ctx->shard_first[i] * (size_t)ctx->n : NULL;
        jobs[i].costs = costs ? costs + ctx->shard_first[i] : NULL;
        jobs[i].rc = -2;
    }
    for (int i = 1; i < S; i++)
        if (pthread_create(&th[i], NULL, worker, &jobs[i])) jobs[i].rc = -3;   /* no thread: done inline below */
    worker(&jobs[0]);
    for (int i = 1; i < S; i++) {
        if (jobs[i].rc == -3) worker(&jobs[i]);
        else pthread_join(th[i], NULL);
    }
    for (int i = 0; i < S; i++)
        if (jobs[i].rc) { memcpy(g_err, jobs[i].err, sizeof g_err); rc = -1; break; }
    free(jobs); free(th);
    return rc;
}

MH_API int KernelResults(mhContext *ctx, point *points, resultCosts *costs)
{
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (!IS_MULTI(ctx)) return results_single(ctx, points, costs);
    return fetch_all_shards(ctx, points, costs, results_worker);
}

MH_API int KernelDeviceResults(mhContext *ctx, void **d_points, void **d_costs)
{
    int prev = -1, rc = -1;
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (IS_MULTI(ctx)) return multi_unsupported("KernelDeviceResults");
    CU(enter_device(ctx->device, &prev));
    CU(ensure_scored(ctx));
    if (d_points) *d_points = ctx->d_points;
    if (d_costs) *d_costs = ctx->d_costs;
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelSetStream(mhContext *ctx, void *stream)
{
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (IS_MULTI(ctx)) return multi_unsupported("KernelSetStream");
    if (KernelSynchronize(ctx)) return -1;
    if (ctx->own_stream) {
        int prev = -1;
        if (!enter_device(ctx->device, &prev)) {
            mhdev_stream_destroy(ctx->stream);
            leave_device(ctx->device, prev);
        }
        ctx->own_stream = 0;
    }
    ctx->stream = stream;
    if (!stream) { /* NULL = a stream of our own; the legacy default stream is the handle 0x1 */
        int prev = -1, e = enter_device(ctx->device, &prev);
        if (!e) e = mhdev_stream_create(&ctx->stream);
        if (prev >= 0) leave_device(ctx->device, prev);
        if (e) { set_err("%s failed: %s", "stream create", e); return -1; }
        ctx->own_stream = 1;
    }
    return 0;
}

/* rank key (csrc/mh_kernels.cu: rank_key) -> chain index and totalCosts; 0 = no chain */
static int decode_rank_key(uint64_t key, float *total)
{
    if (!key) return -1;
    uint32_t u = (uint32_t)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    if (total) memcpy(total, &u, 4);
    return (int)(0xFFFFFFFFu - (uint32_t)key);
}

/* Enqueue the arg-max of one context and the read-back of its key into the context's pinned scratch word. */
static int best_enqueue(mhContext *ctx)
{
    int prev = -1, rc = -1;
    CU(enter_device(ctx->device, &prev));
    CU(ensure_scored(ctx));
    CU(mhdev_launch_argmax(ctx->d_costs, ctx->n_chains, ctx->d_scratch, ctx->stream));
    ctx->launches++;
    CU(mhdev_d2h(ctx->h_scratch, ctx->d_scratch, 8, ctx->stream));
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelBest(mhContext *ctx, int *bestChain, float *bestTotal)
{
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    mhContext *one[1] = { ctx };
    mhContext **list = IS_MULTI(ctx) ? ctx->shards : one;
    const int S = IS_MULTI(ctx) ? ctx->n_shards : 1;
    for (int i = 0; i < S; i++)                                 /* every device reduces at the same time ... */
        if (best_enqueue(list[i])) return -1;
    int best = -1;
    float best_total = 0.f;
    for (int i = 0; i < S; i++) {                               /* ... and the host compares at most 8 keys */
        if (KernelSynchronize(list[i])) return -1;
        float t = 0.f;
        const int idx = decode_rank_key(*(const uint64_t *)list[i]->h_scratch, &t);
        if (idx < 0) continue;
        if (best < 0 || t > best_total) {                       /* shards are in chain order: a tie stays with the lower id */
            best = (IS_MULTI(ctx) ? ctx->shard_first[i] : 0) + idx;
            best_total = t;
        }
    }
    if (bestChain) *bestChain = best;
    if (bestTotal) *bestTotal = best_total;
    return 0;
}

typedef struct rankItem { float total; int32_t chain; } rankItem;

static int rank_cmp(const void *a, const void *b)
{
    const rankItem *x = (const rankItem *)a, *y = (const rankItem *)b;
    if (x->total != y->total) return x->total > y->total ? -1 : 1; /* NaN sorts wherever; totals are finite in practice */
    return x->chain < y->chain ? -1 : (x->chain > y->chain);
}

#define MH_TOPK_DEVICE_MAX 512 /* kTopkMaxK of mh_kernels.cu */

/* The k best chains of ONE context into items[0..k), best first.  k <= 512: sorted on the device (tiles of 1024
 * rank keys, bitonic in shared memory, survivors merged stage by stage), only k keys cross PCIe; larger k:
 * every total is read back and sorted here. */
static int topk_single(mhContext *ctx, int k, rankItem *items)
{
    int prev = -1, rc = -1;
    void *d_work = NULL, *d_out = NULL;
    uint64_t *hk = NULL;
    resultCosts *hc = NULL;
    rankItem *all = NULL;
    if (k > ctx->n_chains) k = ctx->n_chains;
    CU(enter_device(ctx->device, &prev));
    CU(ensure_scored(ctx));
    if (k <= MH_TOPK_DEVICE_MAX) {
        hk = (uint64_t *)malloc(8 * (size_t)k);
        if (!hk) { set_err("", "out of host memory", 0); goto fail; }
        CU(mhdev_malloc(&d_work, 16 * (size_t)mhdev_topk_work_items(ctx->n_chains, k), ctx->stream));
        CU(mhdev_malloc(&d_out, 8 * (size_t)k, ctx->stream));
        CU(mhdev_launch_topk(ctx->d_costs, ctx->n_chains, k, d_work, d_out, ctx->stream));
        ctx->launches++;
        CU(mhdev_d2h(hk, d_out, 8 * (size_t)k, ctx->stream));
        CU(mhdev_stream_sync(ctx->stream));
        for (int i = 0; i < k; i++) items[i].chain = decode_rank_key(hk[i], &items[i].total);
    } else {
        hc = (resultCosts *)malloc(sizeof(resultCosts) * (size_t)ctx->n_chains);
        all = (rankItem *)malloc(sizeof(rankItem) * (size_t)ctx->n_chains);
        if (!hc || !all) { set_err("", "out of host memory", 0); goto fail; }
        CU(mhdev_d2h(hc, ctx->d_costs, sizeof(resultCosts) * (size_t)ctx->n_chains, ctx->stream));
        CU(mhdev_stream_sync(ctx->stream));
        for (int i = 0; i < ctx->n_chains; i++) { all[i].total = hc[i].totalCosts; all[i].chain = i; }
        qsort(all, (size_t)ctx->n_chains, sizeof(rankItem), rank_cmp);
        memcpy(items, all, sizeof(rankItem) * (size_t)k);
    }
    rc = k;
fail:
    if (d_work || d_out) { mhdev_stream_sync(ctx->stream); mhdev_free(d_work, ctx->stream); mhdev_free(d_out, ctx->stream); }
    free(hk); free(hc); free(all);
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelTopK(mhContext *ctx, int k, int *chains, float *totals)
{
    g_err[0] = 0;
    if (!ctx || k < 1) { set_err("", "bad arguments", 0); return -1; }
    if (k > ctx->n_chains) k = ctx->n_chains;
    const int S = IS_MULTI(ctx) ? ctx->n_shards : 1;
    rankItem *items = (rankItem *)malloc(sizeof(rankItem) * (size_t)k * (size_t)S);
    if (!items) { set_err("", "out of host memory", 0); return -1; }
    int have = 0;
    if (IS_MULTI(ctx)) {                                         /* the k best of every device, merged here */
        for (int i = 0; i < S; i++) {
            const int m = topk_single(ctx->shards[i], k, items + have);
            if (m < 0) { free(items); return -1; }
            for (int j = 0; j < m; j++) items[have + j].chain += ctx->shard_first[i];
            have += m;
        }
        qsort(items, (size_t)have, sizeof(rankItem), rank_cmp);
    } else {
        have = topk_single(ctx, k, items);
        if (have < 0) { free(items); return -1; }
    }
    if (k > have) k = have;
    for (int i = 0; i < k; i++) {
        if (chains) chains[i] = items[i].chain;
        if (totals) totals[i] = items[i].total;
    }
    free(items);
    return k;
}

/* KernelTopKDistinct on a context spread over several devices.  Every round each device runs the fused round kernel on
 * its own chains (distance to the previous pick, masked arg-max), the host compares the <= 8 candidate keys, reads the
 * winner's layout (n x 24 bytes) from its device and hands it to every device as the next round's reference.  One small
 * host round trip per pick instead of none -- the price of the layouts living on different devices. */
typedef struct distinctShard { float *d_mind; void *d_keys, *d_ref; } distinctShard;

static int topk_distinct_multi(mhContext *ctx, int k, float minDistance, float rotWeight, int *chains, float *totals)
{
    const int S = ctx->n_shards, n = ctx->n;
    int rc = -1, found = 0, prev = -1;
    distinctShard *ds = (distinctShard *)calloc((size_t)S, sizeof *ds);
    point *ref = (point *)malloc(sizeof(point) * (size_t)n);
    if (!ds || !ref) { set_err("", "out of host memory", 0); goto done; }
    for (int i = 0; i < S; i++) {
        mhContext *c = ctx->shards[i];
        prev = -1;
        CU(enter_device(c->device, &prev));
        CU(ensure_scored(c));
        CU(mhdev_malloc((void **)&ds[i].d_mind, sizeof(float) * (size_t)c->n_chains, c->stream));
        CU(mhdev_malloc(&ds[i].d_keys, 8 * (size_t)k, c->stream));
        CU(mhdev_malloc(&ds[i].d_ref, sizeof(point) * (size_t)n, c->stream));
        CU(mhdev_memset(ds[i].d_keys, 0, 8 * (size_t)k, c->stream));
        leave_device(c->device, prev);
        prev = -1;
    }
    for (found = 0; found < k; found++) {
        for (int i = 0; i < S; i++) {                           /* every device works on the round at the same time */
            mhContext *c = ctx->shards[i];
            prev = -1;
            CU(enter_device(c->device, &prev));
            CU(mhdev_launch_distinct_round(c->d_costs, c->d_points, n, c->n_chains, found, minDistance, rotWeight, (float)(2 * MH_PI),
                                           ds[i].d_mind, ds[i].d_keys, found ? ds[i].d_ref : NULL, c->stream));
            c->launches++;
            CU(mhdev_d2h(c->h_scratch, (const char *)ds[i].d_keys + 8 * (size_t)found, 8, c->stream));
            leave_device(c->device, prev);
            prev = -1;
        }
        int best_shard = -1, best_idx = -1;
        float best_total = 0.f;
        for (int i = 0; i < S; i++) {
            if (KernelSynchronize(ctx->shards[i])) goto done;
            float t = 0.f;
            const int idx = decode_rank_key(*(const uint64_t *)ctx->shards[i]->h_scratch, &t);
            if (idx >= 0 && (best_shard < 0 || t > best_total)) { best_shard = i; best_idx = idx; best_total = t; }
        }
        if (best_shard < 0) break;                              /* every remaining chain is a near-duplicate */
        if (chains) chains[found] = ctx->shard_first[best_shard] + best_idx;
        if (totals) totals[found] = best_total;
        if (found + 1 < k) {                                    /* the pick's layout becomes every device's next reference */
            mhContext *w = ctx->shards[best_shard];
            prev = -1;
            CU(enter_device(w->device, &prev));
            CU(mhdev_d2h(ref, (const char *)w->d_points + sizeof(point) * (size_t)best_idx * (size_t)n, sizeof(point) * (size_t)n, w->stream));
            CU(mhdev_stream_sync(w->stream));
            leave_device(w->device, prev);
            prev = -1;
            for (int i = 0; i < S; i++) {
                mhContext *c = ctx->shards[i];
                CU(enter_device(c->device, &prev));
                CU(mhdev_h2d(ds[i].d_ref, ref, sizeof(point) * (size_t)n, c->stream));
                CU(mhdev_stream_sync(c->stream));               /* `ref` is pageable and reused */
                leave_device(c->device, prev);
                prev = -1;
            }
        }
    }
    rc = found;
    goto done;
fail:
    if (prev >= 0) mhdev_set_device(prev);
done:
    for (int i = 0; ds && i < S; i++) {
        mhContext *c = ctx->shards[i];
        int p2 = -1;
        if (!enter_device(c->device, &p2)) {
            mhdev_stream_sync(c->stream);
            mhdev_free(ds[i].d_mind, c->stream); mhdev_free(ds[i].d_keys, c->stream); mhdev_free(ds[i].d_ref, c->stream);
            leave_device(c->device, p2);
        }
    }
    free(ds); free(ref);
    return rc;
}

MH_API int KernelTopKDistinct(mhContext *ctx, int k, float minDistance, float rotWeight, int *chains, float *totals)
{
    int prev = -1, rc = -1, found = 0;
    float *d_mind = NULL;
    void *d_keys = NULL;
    uint64_t *hk = NULL;
    g_err[0] = 0;
    if (!ctx || k < 1 || !(minDistance >= 0.f) || !(rotWeight >= 0.f)) { set_err("", "bad arguments", 0); return -1; }
    if (k > ctx->n_chains) k = ctx->n_chains;
    if (IS_MULTI(ctx)) return topk_distinct_multi(ctx, k, minDistance, rotWeight, chains, totals);
    hk = (uint64_t *)malloc(8 * (size_t)k);
    if (!hk) { set_err("", "out of host memory", 0); return -1; }
    CU(enter_device(ctx->device, &prev));
    CU(ensure_scored(ctx));
    /* k rounds enqueued back to back: round r reads round r-1's pick on the device, updates every chain's distance
     * to the picks so far and arg-maxes what is still far enough (one fused kernel per round); ONE read-back of
     * the k picks at the end instead of a blocking 8-byte copy per pick */
    CU(mhdev_malloc((void **)&d_mind, sizeof(float) * (size_t)ctx->n_chains, ctx->stream));
    CU(mhdev_malloc(&d_keys, 8 * (size_t)k, ctx->stream));
    CU(mhdev_memset(d_keys, 0, 8 * (size_t)k, ctx->stream));
    for (int r = 0; r < k; r++) {
        CU(mhdev_launch_distinct_round(ctx->d_costs, ctx->d_points, ctx->n, ctx->n_chains, r, minDistance, rotWeight, (float)(2 * MH_PI),
                                       d_mind, d_keys, NULL, ctx->stream));
        ctx->launches++;
    }
    CU(mhdev_d2h(hk, d_keys, 8 * (size_t)k, ctx->stream));
    CU(mhdev_stream_sync(ctx->stream));
    for (found = 0; found < k; found++) {
        float t = 0.f;
        const int idx = decode_rank_key(hk[found], &t);
        if (idx < 0) break;                                     /* every remaining chain is a near-duplicate */
        if (chains) chains[found] = idx;
        if (totals) totals[found] = t;
    }
    rc = found;
fail:
    if (d_mind || d_keys) { mhdev_stream_sync(ctx->stream); mhdev_free(d_mind, ctx->stream); mhdev_free(d_keys, ctx->stream); }
    free(hk);
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelBestKey(mhContext *ctx, void *d_key)
{
    int prev = -1, rc = -1;
    g_err[0] = 0;
    if (!ctx || !d_key) { set_err("", "bad arguments", 0); return -1; }
    if (IS_MULTI(ctx)) return multi_unsupported("KernelBestKey");
    /* the key keeps 32 bits of the global chain id */
    if (ctx->opt.chain_offset + (uint64_t)(ctx->n_chains - 1) * ctx->opt.chain_stride > 0xFFFFFFFFull) {
        set_err("", "KernelBestKey: this context holds global chain ids >= 2^32, the packed key keeps 32 bits", 0);
        return -1;
    }
    CU(enter_device(ctx->device, &prev));
    CU(ensure_scored(ctx));
    CU(mhdev_launch_argmax(ctx->d_costs, ctx->n_chains, ctx->d_scratch, ctx->stream));
    CU(mhdev_launch_bestkey(ctx->d_scratch, ctx->opt.chain_offset, ctx->opt.chain_stride, d_key, ctx->stream));
    ctx->launches += 2;
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API void KernelDecodeBestKey(long long key, unsigned long long *globalChain, float *total)
{
    const uint64_t k = (uint64_t)key ^ 0x8000000000000000ull;
    uint32_t u = (uint32_t)(k >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    if (total) memcpy(total, &u, 4);
    if (globalChain) *globalChain = (unsigned long long)(0xFFFFFFFFu - (uint32_t)k);
}

MH_API int KernelReset(mhContext *ctx)
{
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (IS_MULTI(ctx)) {
        for (int i = 0; i < ctx->n_shards; i++)
            if (KernelReset(ctx->shards[i])) return -1;
        return 0;
    }
    ctx->fresh = 1;
    ctx->it_done = 0;
    ctx->costs_dirty = 1;
    if (ctx->opt.tempering_rungs > 1) {
        int prev = -1, rc = -1;
        CU(enter_device(ctx->device, &prev));
        CU(init_betas(ctx));
        CU(mhdev_memset(ctx->d_exch_stats, 0, 16 * (size_t)ctx->opt.tempering_rungs, ctx->stream));
        rc = 0;
    fail:
        if (prev >= 0) leave_device(ctx->device, prev);
        return rc;
    }
    return 0;
}

MH_API int KernelTemperingState(mhContext *ctx, void **d_totals, void **d_betas)
{
    g_err[0] = 0;
    if (ctx && IS_MULTI(ctx)) return multi_unsupported("KernelTemperingState");
    if (!ctx || ctx->opt.tempering_rungs <= 1) { set_err("", "context has no tempering ladder", 0); return -1; }
    if (d_totals) *d_totals = ctx->d_cur;
    if (d_betas) *d_betas = ctx->d_beta;
    return 0;
}

MH_API int KernelTemperingExchange(mhContext *ctx, const void *d_all_totals, const void *d_all_betas)
{
    int prev = -1, rc = -1;
    g_err[0] = 0;
    if (ctx && IS_MULTI(ctx)) return multi_unsupported("KernelTemperingExchange");
    if (!ctx || ctx->opt.tempering_rungs <= 1 || !d_all_totals || !d_all_betas) { set_err("", "bad arguments", 0); return -1; }
    const uint64_t ex = (uint64_t)ctx->opt.exchange_interval;
    const uint64_t gnow = ctx->opt.iteration_offset + ctx->it_done;
    if (gnow == 0 || gnow % ex) { set_err("", "not on an exchange boundary", 0); return -1; }
    CU(enter_device(ctx->device, &prev));
    CU(mhdev_launch_exchange(ctx->n_chains, ctx->opt.chain_offset, ctx->opt.chain_stride, ctx->opt.tempering_rungs, gnow / ex, gnow - 1,
                             ctx->opt.seed, (const float *)d_all_totals, (const float *)d_all_betas, 0, ctx->opt.chain_stride,
                             (uint64_t)ctx->n_chains, ctx->d_beta, ctx->d_exch_stats, ctx->stream));
    ctx->launches++;
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelTemperingStats(mhContext *ctx, long long *attempts, long long *accepted)
{
    int prev = -1, rc = -1;
    unsigned long long *h = NULL;
    g_err[0] = 0;
    if (!ctx || ctx->opt.tempering_rungs <= 1) { set_err("", "context has no tempering ladder", 0); return -1; }
    const int pairs = ctx->opt.tempering_rungs - 1;
    if (IS_MULTI(ctx)) {                                         /* whole ladders per device: the counts add up */
        long long *a = (long long *)calloc(2 * (size_t)pairs + 2, sizeof *a);
        if (!a) { set_err("", "out of host memory", 0); return -1; }
        for (int r = 0; r < pairs; r++) { if (attempts) attempts[r] = 0; if (accepted) accepted[r] = 0; }
        for (int i = 0; i < ctx->n_shards; i++) {
            if (KernelTemperingStats(ctx->shards[i], a, a + pairs) != pairs) { free(a); return -1; }
            for (int r = 0; r < pairs; r++) { if (attempts) attempts[r] += a[r]; if (accepted) accepted[r] += a[pairs + r]; }
        }
        free(a);
        return pairs;
    }
    h = (unsigned long long *)malloc(16 * (size_t)ctx->opt.tempering_rungs);
    if (!h) { set_err("", "out of host memory", 0); return -1; }
    CU(enter_device(ctx->device, &prev));
    CU(mhdev_d2h(h, ctx->d_exch_stats, 16 * (size_t)ctx->opt.tempering_rungs, ctx->stream));
    CU(mhdev_stream_sync(ctx->stream));
    for (int r = 0; r < pairs; r++) {
        if (attempts) attempts[r] = (long long)h[2 * r];
        if (accepted) accepted[r] = (long long)h[2 * r + 1];
    }
    rc = pairs;
fail:
    free(h);
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelTemperingLadder(mhContext *ctx, double *betas)
{
    g_err[0] = 0;
    if (!ctx || ctx->opt.tempering_rungs <= 1) { set_err("", "context has no tempering ladder", 0); return -1; }
    const mhContext *src = IS_MULTI(ctx) ? ctx->shards[0] : ctx;
    for (int r = 0; betas && r < ctx->opt.tempering_rungs; r++) betas[r] = (double)src->ladder[r];
    return ctx->opt.tempering_rungs;
}

MH_API int KernelTemperingProposeLadder(int rungs, const double *current, const long long *attempts, const long long *accepted,
                                        double damping, double *proposed)
{
    g_err[0] = 0;
    if (rungs < 2 || !current || !attempts || !accepted || !proposed || !(damping > 0.0) || damping > 1.0) { set_err("", "bad arguments", 0); return -1; }
    for (int r = 0; r < rungs; r++)
        if (!(current[r] > 0.0)) { set_err("", "betas must be positive", 0); return -1; }
    double *cum = (double *)malloc(sizeof(double) * (size_t)rungs);
    if (!cum) { set_err("", "out of host memory", 0); return -1; }
    cum[0] = 0.0;
    for (int r = 0; r + 1 < rungs; r++) {                       /* the "distance" of every gap */
        double rate = attempts[r] >= 8 ? (double)accepted[r] / (double)attempts[r] : 0.5;
        if (rate < 0.01) rate = 0.01;
        if (rate > 0.99) rate = 0.99;
        cum[r + 1] = cum[r] - log(rate);
    }
    proposed[0] = current[0];
    proposed[rungs - 1] = current[rungs - 1];
    int gap = 0;
    for (int r = 1; r + 1 < rungs; r++) {                       /* rung r at distance r/(T-1) of the total */
        const double want = cum[rungs - 1] * (double)r / (double)(rungs - 1);
        while (gap + 2 < rungs && cum[gap + 1] < want) gap++;
        const double span = cum[gap + 1] - cum[gap];
        const double f = span > 0 ? (want - cum[gap]) / span : 0.0;
        const double lb = log(current[gap]) + f * (log(current[gap + 1]) - log(current[gap]));
        proposed[r] = exp(log(current[r]) + damping * (lb - log(current[r])));
    }
    free(cum);
    return 0;
}

MH_API int KernelTemperingSetLadder(mhContext *ctx, const double *betas)
{
    int prev = -1, rc = -1;
    void *d_lad = NULL;
    float *h = NULL;
    g_err[0] = 0;
    if (!ctx || ctx->opt.tempering_rungs <= 1 || !betas) { set_err("", "context has no tempering ladder", 0); return -1; }
    if (IS_MULTI(ctx)) {
        for (int i = 0; i < ctx->n_shards; i++)
            if (KernelTemperingSetLadder(ctx->shards[i], betas)) return -1;
        return 0;
    }
    const int T = ctx->opt.tempering_rungs;
    for (int r = 0; r < T; r++)
        if (!(betas[r] > 0.0)) { set_err("", "betas must be positive", 0); return -1; }
    h = (float *)malloc(8 * (size_t)T);
    if (!h) { set_err("", "out of host memory", 0); return -1; }
    for (int r = 0; r < T; r++) { h[r] = ctx->ladder[r]; h[T + r] = (float)betas[r]; }
    CU(enter_device(ctx->device, &prev));
    CU(mhdev_malloc(&d_lad, 8 * (size_t)T, ctx->stream));
    CU(mhdev_h2d(d_lad, h, 8 * (size_t)T, ctx->stream));
    CU(mhdev_launch_retarget(ctx->n_chains, T, (const float *)d_lad, (const float *)d_lad + T, ctx->d_beta, ctx->stream));
    ctx->launches++;
    CU(mhdev_memset(ctx->d_exch_stats, 0, 16 * (size_t)T, ctx->stream));
    CU(mhdev_stream_sync(ctx->stream));                         /* h and d_lad are released below */
    for (int r = 0; r < T; r++) ctx->ladder[r] = h[T + r];
    rc = 0;
fail:
    if (d_lad) mhdev_free(d_lad, ctx->stream);
    free(h);
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelStats(mhContext *ctx, double *kernel_ms, long long *launches)
{
    int prev = -1, rc = -1;
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (IS_MULTI(ctx)) {                                         /* concurrent devices: the slowest one's time, everybody's launches */
        double ms_max = 0;
        long long total = 0;
        for (int i = 0; i < ctx->n_shards; i++) {
            double ms = 0;
            long long l = 0;
            if (KernelStats(ctx->shards[i], &ms, &l)) return -1;
            if (ms > ms_max) ms_max = ms;
            total += l;
        }
        if (kernel_ms) *kernel_ms = ms_max;
        if (launches) *launches = total;
        return 0;
    }
    CU(enter_device(ctx->device, &prev));
    CU(drain_events(ctx));
    if (kernel_ms) *kernel_ms = ctx->kernel_ms;
    if (launches) *launches = ctx->launches;
    rc = 0;
fail:
    if (prev >= 0) leave_device(ctx->device, prev);
    return rc;
}

MH_API int KernelShape(mhContext *ctx, int *lanesPerChain, int *evalForm, int *nDevices, int *deviceOrdinals, int *chainsPerDevice)
{
    g_err[0] = 0;
    if (!ctx) { set_err("", "null context", 0); return -1; }
    if (lanesPerChain) *lanesPerChain = ctx->lanes;
    if (evalForm) *evalForm = ctx->eval_internal == 0 ? MH_EVAL_FULL_SCAN : ctx->eval_internal;
    const int S = IS_MULTI(ctx) ? ctx->n_shards : 1;
    if (nDevices) *nDevices = S;
    for (int i = 0; i < MH_MAX_DEVICES; i++) {
        if (deviceOrdinals) deviceOrdinals[i] = i < S ? (IS_MULTI(ctx) ? ctx->shards[i]->device : ctx->device) : -1;
        if (chainsPerDevice) chainsPerDevice[i] = i < S ? (IS_MULTI(ctx) ? ctx->shards[i]->n_chains : ctx->n_chains) : 0;
    }
    return 0;
}

MH_API int KernelDeviceCount(void)
{
    int count = 0;
    g_err[0] = 0;
    const int e = mhdev_device_count(&count);
    if (e) { set_err("%s failed: %s", "device count", e); return -1; }
    return count;
}

MH_API void KernelDestroy(mhContext *ctx)
{
    if (!ctx) return;
    if (IS_MULTI(ctx)) {
        for (int i = 0; i < ctx->n_shards; i++) destroy_single(ctx->shards[i]);
        ctx_free(ctx);
        return;
    }
    destroy_single(ctx);
}

/* ---- one-shot calls: chunked run, scoring and D2H of a chunk overlap the kernels of the chunks after it -------- */

/* In how many launches a one-shot call runs this context's chains.  The D2H of the result block into the caller's
 * pageable memory (~10-17 GB/s) is the one part of KernelWrapper that cannot be made faster -- but it can be hidden:
 * with K chunks only the last chunk's copy is exposed.  One chunk per 48 MB of results, at most 8, never fewer than
 * 8192 chains per chunk (every chunk must still fill the machine); env MH_CHUNKS overrides. */
static int oneshot_chunks(const mhContext *c)
{
    const char *env = getenv("MH_CHUNKS");
    if (c->opt.tempering_rungs > 1) return 1;                   /* ladders exchange between launches: one launch sequence */
    int k;
    if (env && *env) k = atoi(env);
    else {
        const double bytes = (double)c->n_chains * (double)c->n * (double)sizeof(point);
        k = (int)(bytes / (48.0 * 1048576.0));
        while (k > 1 && c->n_chains / k < 8192) k--;
    }
    if (k > MH_MAX_CHUNKS) k = MH_MAX_CHUNKS;
    if (k > c->n_chains) k = c->n_chains;
    return k < 1 ? 1 : k;
}

/* phase 0: every chunk's chain kernel on the context's stream, an event after each (asynchronous) */
static int oneshot_enqueue(mhContext *c, int iterations)
{
    int prev = -1, rc = -1;
    const int K = oneshot_chunks(c);
    c->n_chunks = 0;
    if (K <= 1) return run_iterations(c, iterations, NULL);
    CU(enter_device(c->device, &prev));
    if (c->opt.schedule_length <= 0 && c->opt.schedule != MH_SCHEDULE_CONSTANT) c->opt.schedule_length = iterations;
    if (!c->copy_stream) CU(mhdev_stream_create(&c->copy_stream));
    for (int j = 0; j <= K; j++) {
        long long f = (long long)c->n_chains * j / K;
        if (j > 0 && j < K) f = (f + 127) / 256 * 256;          /* whole blocks */
        if (f > c->n_chains) f = c->n_chains;
        c->chunk_first[j] = (int)f;
    }
    for (int j = 0; j < K; j++) {
        const int first = c->chunk_first[j], count = c->chunk_first[j + 1] - first;
        if (count > 0) CU(launch_chains_range(c, first, count, iterations, NULL));
        CU(mhdev_event_create(&c->chunk_ev[j]));
        c->n_chunks = j + 1;
        CU(mhdev_event_record(c->chunk_ev[j], c->stream));
    }
    c->fresh = 0;
    c->it_done += (uint64_t)iterations;
    c->costs_dirty = 1;
    rc = 0;
fail:
    if (prev >= 0) leave_device(c->device, prev);
    return rc;
}

/* phase 1: chunk by chunk -- wait for its kernel, score it, copy it out (the copy blocks this thread, not the GPU) */
static int oneshot_fetch(mhContext *c, point *points, resultCosts *costs)
{
    int prev = -1, rc = -1;
    if (c->n_chunks == 0) return results_single(c, points, costs);
    CU(enter_device(c->device, &prev));
    for (int j = 0; j < c->n_chunks; j++) {
        const int first = c->chunk_first[j], count = c->chunk_first[j + 1] - first;
        const size_t fo = (size_t)first * (size_t)c->n;
        CU(mhdev_stream_wait_event(c->copy_stream, c->chunk_ev[j]));
        if (count <= 0) continue;
        CU(mhdev_launch_score(c->d_problem, c->smem_words, c->n, c->C, c->R, count, c->score_lanes, (char *)c->d_points + sizeof(point) * fo,
                              (char *)c->d_costs + sizeof(resultCosts) * (size_t)first, c->copy_stream));
        c->launches++;
        if (points) CU(mhdev_d2h(points + fo, (char *)c->d_points + sizeof(point) * fo, sizeof(point) * (size_t)count * (size_t)c->n, c->copy_stream));
        if (costs) CU(mhdev_d2h(costs + first, (char *)c->d_costs + sizeof(resultCosts) * (size_t)first, sizeof(resultCosts) * (size_t)count, c->copy_stream));
    }
    CU(mhdev_stream_sync(c->copy_stream));
    CU(mhdev_stream_sync(c->stream));
    c->costs_dirty = 0;
    rc = 0;
fail:
    for (int j = 0; j < c->n_chunks; j++) mhdev_event_destroy(c->chunk_ev[j]);
    c->n_chunks = 0;
    if (prev >= 0) leave_device(c->device, prev);
    return rc;
}

static void *oneshot_worker(void *arg)
{
    shardJob *j = (shardJob *)arg;
    j->rc = oneshot_fetch(j->c, j->points, j->costs);
    if (j->rc) memcpy(j->err, g_err, sizeof j->err);
    return NULL;
}

static int oneshot_run(mhContext *ctx, int iterations)
{
    if (!IS_MULTI(ctx)) return oneshot_enqueue(ctx, iterations);
    for (int i = 0; i < ctx->n_shards; i++)
        if (oneshot_enqueue(ctx->shards[i], iterations)) return -1;
    return 0;
}

static int oneshot_results(mhContext *ctx, point *points, resultCosts *costs)
{
    g_err[0] = 0;
    if (!IS_MULTI(ctx)) return oneshot_fetch(ctx, points, costs);
    return fetch_all_shards(ctx, points, costs, oneshot_worker);
}

/* ---------------------------------------------------------------------------------------------
 * One-shot entry points
 * --------------------------------------------------------------------------------------------- */

MH_API result *KernelWrapperEx(const relationshipStruct *rss, const relationshipAngleStruct *rsa,
                               const positionAndRotation *cfg, const rectangle *clearances, const rectangle *offlimits,
                               const vertex *vertices, const vertex *surfaceRectangle, const Surface *srf,
                               const gpuConfig *gpuCfg, const mhOptions *opt)
{
    g_err[0] = 0;
    if (!gpuCfg || !srf) { set_err("", "null argument", 0); return NULL; }
    const int chains = gpuCfg->gridxDim, iterations = gpuCfg->iterations;
    if (chains < 1 || iterations < 0) {
        snprintf(g_err, sizeof g_err, "bad gpuConfig: gridxDim=%d iterations=%d", chains, iterations);
        return NULL;
    }
    const int n = srf->nObjs;
    const int timing = getenv("MH_TIMING") != NULL;
    struct timespec ts[6];
    clock_gettime(CLOCK_MONOTONIC, &ts[0]);
    mhContext *c = KernelCreate(rss, rsa, cfg, clearances, offlimits, vertices, surfaceRectangle, srf, chains, opt);
    if (!c) return NULL;
    clock_gettime(CLOCK_MONOTONIC, &ts[1]);
    /* Kernel.cu:928, 970: ONE block of points and one array of results, both plain malloc */
    point *pts = (point *)malloc(sizeof(point) * (size_t)chains * (size_t)n);
    result *res = (result *)malloc(sizeof(result) * (size_t)chains);
    resultCosts *costs = (resultCosts *)malloc(sizeof(resultCosts) * (size_t)chains);
    if (!pts || !res || !costs) { set_err("", "out of host memory", 0); goto fail; }
    clock_gettime(CLOCK_MONOTONIC, &ts[2]);
    if (oneshot_run(c, iterations)) goto fail;
    prefault(pts, sizeof(point) * (size_t)chains * (size_t)n); /* overlaps with the kernel, which is asynchronous */
    /* MH_PIN_RESULT=1: page-lock the (malloc'd, caller-owned) result block in place for the duration of the copy,
     * also while the kernel runs; the copy then runs at pinned-memory speed and, on a multi-device context, truly
     * concurrently.  Off by default: measured in profiles/ (registration cost against copy time). */
    int pinned = 0;
    if (getenv("MH_PIN_RESULT") && sizeof(point) * (size_t)chains * (size_t)n >= (1u << 20))
        pinned = mhdev_host_register(pts, sizeof(point) * (size_t)chains * (size_t)n) == 0;
    if (timing) KernelSynchronize(c);
    clock_gettime(CLOCK_MONOTONIC, &ts[3]);
    const int res_rc = oneshot_results(c, pts, costs);
    if (pinned) mhdev_host_unregister(pts);
    if (res_rc) goto fail;
    clock_gettime(CLOCK_MONOTONIC, &ts[4]);
    for (int i = 0; i < chains; i++) {
        res[i].points = pts + (size_t)i * (size_t)n; /* Kernel.cu:981 */
        res[i].costs = costs[i];
    }
    free(costs);
    struct timespec td;
    clock_gettime(CLOCK_MONOTONIC, &td);
    KernelDestroy(c);
    clock_gettime(CLOCK_MONOTONIC, &ts[5]);
    if (timing) fprintf(stderr, "[mh] destroy alone %.2f ms\n", 1e3 * (double)(ts[5].tv_sec - td.tv_sec) + 1e-6 * (double)(ts[5].tv_nsec - td.tv_nsec));
    if (timing) {
        double d[5];
        for (int i = 0; i < 5; i++) d[i] = 1e3 * (double)(ts[i + 1].tv_sec - ts[i].tv_sec) + 1e-6 * (double)(ts[i + 1].tv_nsec - ts[i].tv_nsec);
        fprintf(stderr, "[mh] create %.2f ms, host alloc %.2f ms, run %.2f ms, results %.2f ms, assemble+destroy %.2f ms\n", d[0], d[1], d[2],
                d[3], d[4]);
    }
    return res;
fail:
    free(pts); free(res); free(costs);
    {   /* keep the message across the clean-up */
        char keep[sizeof g_err];
        memcpy(keep, g_err, sizeof keep);
        KernelDestroy(c);
        memcpy(g_err, keep, sizeof keep);
    }
    return NULL;
}

MH_API result *KernelWrapper(relationshipStruct *rss, relationshipAngleStruct *rsa, positionAndRotation *cfg,
                             rectangle *clearances, rectangle *offlimits, vertex *vertices, vertex *surfaceRectangle,
                             Surface *srf, gpuConfig *gpuCfg)
{
    static unsigned long long calls = 0;
    mhOptions o;
    default_options(&o);
    const char *env = getenv("MH_SEED");
    if (env && *env)
        o.seed = strtoull(env, NULL, 0);
    else /* Kernel.cu:943 seeds with time(NULL); the call counter keeps two calls in one second apart */
        o.seed = (uint64_t)time(NULL) ^ ((uint64_t)__atomic_fetch_add(&calls, 1ULL, __ATOMIC_RELAXED) << 40);
    return KernelWrapperEx(rss, rsa, cfg, clearances, offlimits, vertices, surfaceRectangle, srf, gpuCfg, &o);
}

MH_API void KernelFree(result *res)
{
    if (!res) return;
    free(res[0].points);
    free(res);
}

MH_API int KernelEvalCosts(const relationshipStruct *rss, const relationshipAngleStruct *rsa, const positionAndRotation *layouts,
                           int nLayouts, const rectangle *clearances, const rectangle *offlimits, const vertex *vertices,
                           const vertex *surfaceRectangle, const Surface *srf, resultCosts *out)
{
    g_err[0] = 0;
    mhProblem P;
    void *d_problem = NULL, *d_points = NULL, *d_costs = NULL, *stream = NULL;
    point *pts = NULL;
    int rc = -1;
    if (nLayouts < 1 || !out || !layouts) { set_err("", "bad arguments", 0); return -1; }
    if (pack_problem(rss, rsa, layouts, clearances, offlimits, vertices, surfaceRectangle, srf, layouts, nLayouts, &P)) return -1;
    const int n = P.h->n;
    const size_t cn = (size_t)nLayouts * (size_t)n;
    pts = (point *)calloc(cn, sizeof(point));
    if (!pts) { set_err("", "out of host memory", 0); goto fail; }
    for (size_t i = 0; i < cn; i++) {
        pts[i].x = (float)layouts[i].x; pts[i].y = (float)layouts[i].y; pts[i].rotY = (float)layouts[i].rotY;
    }
    const int lanes = choose_lanes(n, P.h->C, P.h->R, P.h->smem_words, nLayouts, 0, MH_EVAL_FULL);
    if (lanes < 0) goto fail;
    CU(mhdev_stream_create(&stream));
    CU(mhdev_malloc(&d_problem, 4 * (size_t)P.h->total_words, stream));
    CU(mhdev_malloc(&d_points, sizeof(point) * cn, stream));
    CU(mhdev_malloc(&d_costs, sizeof(resultCosts) * (size_t)nLayouts, stream));
    CU(mhdev_h2d(d_problem, P.blob, 4 * (size_t)P.h->total_words, stream));
    CU(mhdev_h2d(d_points, pts, sizeof(point) * cn, stream));
    CU(mhdev_launch_score(d_problem, P.h->smem_words, n, P.h->C, P.h->R, nLayouts, lanes, d_points, d_costs, stream));
    CU(mhdev_d2h(out, d_costs, sizeof(resultCosts) * (size_t)nLayouts, stream));
    CU(mhdev_stream_sync(stream));
    rc = 0;
fail:
    if (stream) mhdev_stream_sync(stream);
    mhdev_free(d_problem, stream); mhdev_free(d_points, stream); mhdev_free(d_costs, stream);
    if (stream) mhdev_stream_destroy(stream);
    free(pts);
    free(P.blob);
    return rc;
}

MH_API int KernelTrim(void)
{
    g_err[0] = 0;
    int e = mhdev_trim();
    if (e) { set_err("%s failed: %s", "pool trim", e); return -1; }
    return 0;
}

MH_API int KernelDeviceInfo(int *smCount, int *smClockKHz, int *ccMajor, int *ccMinor, char *name, int nameLen)
{
    g_err[0] = 0;
    int e = mhdev_device_limits(NULL, NULL, smCount, smClockKHz, ccMajor, ccMinor, name, nameLen);
    if (e) { set_err("%s failed: %s", "device query", e); return -1; }
    return 0;
}
