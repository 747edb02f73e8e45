/*
 * mh_abi.h -- the thin C ABI between the plain-C host code (kernel_wrapper.c) and the CUDA
 * translation unit (mh_kernels.cu).  Launch descriptors and raw device-memory helpers only:
 * no CUDA type crosses this line, so the host side compiles with a C compiler.
 *
 * Internal to libKernel.so (hidden visibility); the public ABI is include/mh_kernel.h.
 */
#ifndef MH_ABI_H
#define MH_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Packed, single-precision problem description.  The host builds one blob: this header,
 * then the arrays at the word offsets it names.  The first `smem_words` 4-byte words are
 * staged into shared memory by every block; the tail (initial layout, pass-through fields)
 * is read from global memory at chain start / write-out only.
 *
 * Every constant is derived from the caller's doubles in double precision and narrowed once,
 * at the point where the reference narrows (see kernel_wrapper.c: pack_problem). */
typedef struct mhProblemHeader {
    int32_t n, C, R, any_free;
    float w_focal, w_pair, w_visual, w_sym;
    float w_off, w_clear, w_surf, denom;   /* denom = sum of length*width (Kernel.cu:199-202) */
    float focal_x, focal_y, ux, uy;        /* (float)cos/sin(focalRot), Kernel.cu:290-291      */
    float fdotu, two_focal_rot, cx2, cy2;  /* focal . u ; 2*focalRot ; centroid/2 (Q11)        */
    float room_minx, room_miny, room_maxx, room_maxy; /* AABB of surfaceRectangle             */
    float std_x, std_y, sigma_t, pi_cmp;   /* W/16, H/16 (Q19); 15/90*PI; largest float <= 3.1416 */
    float two_pi, half_pi, two_pi_cmp, inv_denom; /* (float)6.2832, (float)1.5708, largest float <= 6.2832, (float)(1 / denom) */
    int32_t off_obj_box;    /* float4[n] {min(v1x,v2x,v3x), min y, max x, max y} of the off-limit rect */
    int32_t off_obj_v0x;    /* float[n]  x of its first vertex, never translated (quirk Q6)   */
    int32_t off_obj_area;   /* float[n]  (float)(length*width)                                */
    int32_t off_obj_frozen; /* int32[n]                                                        */
    int32_t off_clr_box;    /* float4[C] same constants for the clearance rects               */
    int32_t off_clr_v0x;    /* float[C]                                                        */
    int32_t off_clr_src;    /* int32[C]  SourceIndex                                           */
    int32_t off_rel_idx;    /* int4[R]   {rss Source, rss Target, rsa Source, rsa Target}      */
    int32_t off_rel_rng;    /* float4[R] {1/start, end, angleMin, angleMax}                    */
    int32_t off_rel_aux;    /* float4[R] {start, 1/norm, wraps (angleMin > angleMax), 0}       */
    int32_t off_clr_adj_off; /* int32[n+1] CSR offsets: clearances whose SourceIndex is object i   */
    int32_t off_clr_adj;    /* int32[C]   CSR list of clearance indices                            */
    int32_t smem_words;     /* words [0, smem_words) go to shared memory                       */
    int32_t off_cfg0;       /* float[3n] x, y, rotY of the caller's layout (global only)       */
    int32_t off_pass;       /* float[3n] z, rotX, rotZ pass-through, narrowed (global only)    */
    int32_t total_words;
    int32_t off_rel_adj_off; /* int32[n+1] CSR offsets: relationships that name object i (any of the 4 indices) */
    int32_t off_rel_adj;    /* int32[<=4R] CSR list of relationship indices, each at most once per object   */
    int32_t pad3, pad4;
    /* ClearanceCosts in fixed point (csrc/mh_costs.cuh: "the clearance term is an integer sum"): the AABB
     * constants above rounded to multiples of 2^-clr_k, positions are rounded the same way on the device */
    int32_t off_obj_boxq;   /* int4[n]  round(off_obj_box * 2^clr_k)                                   */
    int32_t off_obj_v0xq;   /* int32[n]                                                                */
    int32_t off_clr_boxq;   /* int4[C]                                                                 */
    int32_t off_clr_v0xq;   /* int32[C]                                                                */
    int32_t clr_k;          /* coordinates are held in units of 2^-clr_k                               */
    float clr_scale;        /* 2^clr_k                                                                 */
    float clr_unit;         /* 2^(-2 clr_k): one unit of the integer area sum                          */
    float clr_pos_limit;    /* positions are clamped to +-this before the conversion (no int32 overflow) */
} mhProblemHeader;

enum { MH_SCHED_CONSTANT = 0, MH_SCHED_GEOMETRIC = 1, MH_SCHED_LINEAR = 2, MH_SCHED_PER_CHAIN = 3 };

/* One launch of the chain kernel: advance `n_chains` chains by `it_count` iterations. */
typedef struct mhLaunch {
    const void *d_problem;
    int32_t problem_words;  /* total words of the blob                                         */
    int32_t smem_words;
    int32_t n, C, R;
    int32_t n_chains;       /* chains of this context (local)                                  */
    int32_t lanes;          /* lanes per chain: 1,2,4,8,16,32                                  */
    int32_t fresh;          /* 1: start from the problem's initial layout                      */
    uint64_t seed;
    uint64_t chain_offset;  /* global id of local chain i = chain_offset + i*chain_stride      */
    uint64_t chain_stride;
    uint64_t it_begin;
    int32_t it_count;
    int32_t schedule;       /* MH_SCHED_*                                                      */
    int32_t schedule_length;
    int32_t result_mode;    /* 0 final layout, 1 best layout                                   */
    int32_t eval_mode;      /* 0 full re-evaluation per proposal, 1 delta evaluation            */
    int32_t warps_per_block; /* delta kernel only: 4 or 8 (the other forms are compiled for 4) */
    float beta_start;
    float beta_end;
    float beta_log2_ratio;  /* log2f(beta_end/beta_start)                                      */
    float pad;
    /* chain state, [chain][object] */
    float *d_x, *d_y, *d_rot;
    uint16_t *d_perm;       /* which original object's z/rotX/rotZ sits in slot i (swap moves)  */
    float *d_cur_total;     /* [chain] totalCosts of the current layout                        */
    float *d_best_total;    /* [chain]                                                         */
    float *d_beta;          /* [chain] used when schedule == MH_SCHED_PER_CHAIN                */
    void *d_points;         /* point[chain][object]                                            */
    void *d_costs;          /* resultCosts[chain]                                              */
    void *d_trace;          /* mhTraceEntry[it][chain] or NULL                                 */
    void *stream;
    /* A copy of the problem header.  The launch descriptor is the kernel's parameter, i.e. it lives in the constant
     * bank: every scalar of the header the loops use (reflection axis, scales, weights, room bounds) becomes a
     * constant operand of the arithmetic instruction that needs it instead of a shared-memory load + a register. */
    mhProblemHeader hdr;
} mhLaunch;

int mhdev_launch_chains(const mhLaunch *l);
/* resultCosts of the layouts in d_points (all eight terms). */
int mhdev_launch_score(const void *d_problem, int smem_words, int n, int C, int R, int n_layouts, int lanes,
                       const void *d_points, void *d_costs, void *stream);
/* Replica exchange between neighbouring rungs (extension).  all_total/all_beta hold the
 * values of every chain of the ladders this context takes part in; chain g sits at
 * ((g-gather_base) % gather_stride) * gather_local + (g-gather_base) / gather_stride, i.e. rank-major
 * after an all-gather over gather_stride ranks (gather_stride = 1: plain global order). */
int mhdev_launch_exchange(int n_chains, uint64_t chain_offset, uint64_t chain_stride, int rungs, uint64_t epoch,
                          uint64_t it_last, uint64_t seed, const float *d_all_total, const float *d_all_beta,
                          uint64_t gather_base, uint64_t gather_stride, uint64_t gather_local, float *d_beta,
                          void *d_stats /* uint64[2*rungs] {attempts, accepted} per pair (lower rung), or NULL */, void *stream);
/* Ladder re-targeting: every chain whose beta is (closest to) old_ladder[r] gets new_ladder[r]; both device float[rungs]. */
int mhdev_launch_retarget(int n_chains, int rungs, const float *d_old_ladder, const float *d_new_ladder, float *d_beta, void *stream);
/* Rank keys.  A chain's rank key is the unsigned 64-bit word  orderable(totalCosts) << 32 | (0xFFFFFFFF - chain):
 * an unsigned MAX over keys is the arg-max of totalCosts with ties going to the lower chain index; 0 = "no
 * chain".  (mh_rank_key / mhdev_decode_rank_key) */
/* arg-max of totalCosts over the context's chains, many blocks, one atomicMax per block: *d_key (device
 * uint64, zeroed by this call) = the best chain's rank key. */
int mhdev_launch_argmax(const void *d_costs, int n_chains, void *d_key, void *stream);
/* Top-k by rank key, on the device: d_out[0..k) = the k largest rank keys in descending order (k <= 512).
 * d_work: 2 * mhdev_topk_work_items(n_chains, k) uint64 of scratch. */
int mhdev_topk_work_items(int n_chains, int k);
int mhdev_launch_topk(const void *d_costs, int n_chains, int k, void *d_work, void *d_out, void *stream);
/* Distinct suggestions, round r of k (all rounds are enqueued back to back, no host round trip): the chain
 * picked in round r-1 is read from d_keys[r-1] ON THE DEVICE; every chain's d_mind = min(d_mind, distance to
 * that pick's layout) (round 0: d_mind = +inf); then d_keys[r] = rank key of the best chain with
 * d_mind > min_dist.  d_keys[0..k) must be zeroed before round 0.  d_ref_layout != NULL: the previous pick's layout
 * (n point records on THIS device, handed over by the host: multi-device contexts) instead of d_keys[r-1]. */
int mhdev_launch_distinct_round(const void *d_costs, const void *d_points, int n, int n_chains, int round, float min_dist,
                                float rot_weight, float two_pi, float *d_mind, void *d_keys, const void *d_ref_layout, void *stream);
/* d_key (device int64) = order-preserving (totalCosts, GLOBAL chain id) key of the arg-max rank key. */
int mhdev_launch_bestkey(const void *d_rank_key, uint64_t chain_offset, uint64_t chain_stride, void *d_key, void *stream);
/* Largest dynamic shared memory per block and SM count / clock of the current device. */
int mhdev_device_limits(int *max_smem_per_block, int *max_smem_per_sm, int *sm_count, int *clock_khz, int *cc_major,
                        int *cc_minor, char *name, int name_len);
/* Dynamic shared memory of one block of the chain kernel; `warps` matters for the delta form only. */
int mhdev_chain_smem_bytes(int smem_words, int n, int C, int R, int lanes, int eval_mode, int warps);

/* raw runtime helpers */
int mhdev_device_count(int *count);
int mhdev_get_device(int *dev);
int mhdev_set_device(int dev);
int mhdev_malloc(void **p, size_t bytes, void *stream); /* stream-ordered, from the library's pool */
void mhdev_free(void *p, void *stream);
int mhdev_trim(void);                                    /* return cached blocks to the driver  */
int mhdev_h2d(void *dst, const void *src, size_t bytes, void *stream);
int mhdev_d2h(void *dst, const void *src, size_t bytes, void *stream);
int mhdev_d2d(void *dst, const void *src, size_t bytes, void *stream);
int mhdev_memset(void *dst, int value, size_t bytes, void *stream);
int mhdev_stream_create(void **stream);
void mhdev_stream_destroy(void *stream);
int mhdev_stream_sync(void *stream);
int mhdev_event_create(void **ev);
void mhdev_event_destroy(void *ev);
int mhdev_event_record(void *ev, void *stream);
int mhdev_stream_wait_event(void *stream, void *ev);   /* work enqueued on `stream` after this call waits for `ev` */
int mhdev_event_elapsed_ms(void *e0, void *e1, float *ms); /* synchronises on e1 */
int mhdev_scratch_acquire(void **p);                        /* a 64-byte pinned read-back slot from the process-wide arena */
void mhdev_scratch_release(void *p);
int mhdev_host_alloc(void **p, size_t bytes);               /* pinned staging memory */
void mhdev_host_free(void *p);
int mhdev_host_register(void *p, size_t bytes);             /* page-lock caller memory in place (portable) */
void mhdev_host_unregister(void *p);
const char *mhdev_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* MH_ABI_H */
