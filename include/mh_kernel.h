/*
 * mh_kernel.h -- C ABI of libKernel.so, the B200-native drop-in for the reference DLL
 * `Kernel` (Kernel.vcxproj:22-29).
 *
 * Every entry point is extern "C", takes plain pointers and sizes, and never exposes a
 * CUDA or torch type.  `KernelWrapper` is the one symbol the reference exports
 * (/root/reference/KernelFolder/Kernel/Kernel.cu:873) and keeps its exact signature;
 * everything else is additive.
 *
 * Error convention: the reference prints, resets the device and calls exit(1) on any CUDA
 * error (common/inc/helper_cuda.h:985-999).  This library never exits and never resets a
 * context it does not own: pointer-returning calls return NULL, int-returning calls return
 * non-zero, and KernelLastError() returns the message for the calling thread.
 *
 * There is no CPU fallback: without a CUDA device (or without the sm_100a kernels) every
 * compute entry point fails.
 */
#ifndef MH_KERNEL_H
#define MH_KERNEL_H

#include "mh_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MH_API __declspec(dllexport)
#else
#define MH_API __attribute__((visibility("default")))
#endif

/* ---- options for the extended entry points (all additive; zero = reference behaviour) ---- */

enum {
    MH_SCHEDULE_CONSTANT = 0, /* beta = beta_start for every iteration (reference: BETA = 2.0, Kernel.cu:33) */
    MH_SCHEDULE_GEOMETRIC = 1, /* beta_i = beta_start * (beta_end/beta_start)^(i/(I-1))                      */
    MH_SCHEDULE_LINEAR = 2     /* beta_i = beta_start + (beta_end-beta_start) * i/(I-1)                        */
};

enum {
    MH_RESULT_FINAL = 0, /* each chain returns its final current layout (Kernel.cu:834-842)             */
    MH_RESULT_BEST = 1   /* each chain returns the layout with the highest totalCosts it visited; the
                            reference has this only as dead code (Kernel.cu:779-782, 808-816)            */
};

enum {
    MH_EVAL_FULL = 0, /* every proposal re-evaluates every live cost term from scratch (Kernel.cu:804); for
                         nObjs >= 28 (>= 18 for jobs of 8192 chains and more) the library uses the bit-identical
                         MH_EVAL_MEMO form, which is faster                                               */
    MH_EVAL_FULL_SCAN = 3, /* MH_EVAL_FULL with the plain n^2 scan forced (verification)                    */
    MH_EVAL_MEMO = 2, /* full evaluation through exact memos (symmetry row minima, relationship penalties,
                         surface values; the clearance term is an exact integer sum updated pair by
                         pair): totals bit-identical to MH_EVAL_FULL_SCAN
                         for the same lane width                                                         */
    MH_EVAL_DELTA = 1 /* incremental evaluation: only what the moved objects touch is recomputed, running
                         sums rebuilt from scratch every 128 iterations; statistically equivalent to
                         MH_EVAL_FULL, not bit-identical (csrc/mh_delta.cuh)                            */
};

/* mhOptions.flags */
enum {
    MH_OPT_EXPLICIT_DEVICE = 1u /* `device` names a CUDA ordinal even when it is 0.  Without this bit device <= 0
                                   means the caller's CURRENT device, so that a zero-initialised struct behaves
                                   like the reference, which never calls cudaSetDevice (Kernel.cu:873-984); a
                                   positive `device` is an ordinal with or without the bit                   */
};

#define MH_MAX_DEVICES 8

typedef struct mhOptions {
    uint32_t struct_size;       /* = sizeof(mhOptions); lets the struct grow compatibly: a caller compiled
                                   against an older, shorter struct keeps working (missing tail = zeros)  */
    uint32_t flags;             /* MH_OPT_*                                                          */
    uint64_t seed;              /* Philox key.  KernelWrapper (no options) uses time(NULL) like the
                                   reference (Kernel.cu:943) unless env MH_SEED is set               */
    uint64_t chain_offset;      /* global id of this call's first chain: chain g = chain_offset + i
                                   draws Philox stream g, so a sharded run equals the unsharded one
                                   bit for bit -- provided every shard uses the same lane width (it
                                   fixes the order of the float reductions): pass the job's size in
                                   total_chains (below) or pin lanes_per_chain                       */
    uint64_t iteration_offset;  /* first iteration index (resume: continue the same Philox stream)   */
    double beta_start;          /* 0 -> 2.0                                                          */
    double beta_end;            /* 0 -> beta_start                                                   */
    int32_t schedule;           /* MH_SCHEDULE_*                                                     */
    int32_t schedule_length;    /* iterations over which the schedule runs; 0 -> this call's count   */
    int32_t result_mode;        /* MH_RESULT_*                                                       */
    int32_t eval_mode;          /* MH_EVAL_*                                                         */
    int32_t lanes_per_chain;    /* 0 = choose from nObjs and the job's chain count; else 1,2,4,8,16,32 */
    int32_t device;             /* <= 0: the caller's current device (0 with MH_OPT_EXPLICIT_DEVICE: GPU 0);
                                   > 0: that CUDA ordinal.  Ignored when n_devices > 1                 */
    /* parallel tempering (extension; 0 rungs = off).  Chains are grouped in ladders of
     * `tempering_rungs` consecutive global chain ids; rung r starts at
     * beta_start * (beta_end/beta_start)^(r/(rungs-1)); every `exchange_interval`
     * iterations neighbouring rungs exchange their betas with the usual PT acceptance. */
    int32_t tempering_rungs;
    int32_t exchange_interval;
    /* Strided sharding (0 -> 1): local chain i has global id chain_offset + i*chain_stride.  With
     * tempering and chain_stride = number of ranks, rank r holds every ladder's chains g with
     * g % ranks == r, so neighbouring rungs live on different GPUs and exchange through
     * KernelTemperingExchange (below). */
    uint64_t chain_stride;
    /* ---- appended in round 2 (offset 88) ---- */
    /* Chains of the WHOLE job of which this context is one shard (0 -> this context's own count).  The
     * default lane width and evaluation form are chosen from it, never from the shard's own count, so
     * that every shard of a job -- and the same job run unsharded -- use the same float reduction order
     * and return the same bits without anybody pinning lanes_per_chain. */
    uint64_t total_chains;
    /* Multi-GPU inside ONE process, behind this C ABI (the reference's caller is a single C# process,
     * Kernel.cu:873).  n_devices > 1: the chains [chain_offset, chain_offset + nChains) are split into
     * contiguous ranges over devices[0 .. n_devices-1] (whole ladders when tempering), one context and one
     * stream per device, all running concurrently; results are copied from every device straight into
     * the caller's one result block; KernelBest / KernelTopK merge the per-device answers.  Per-chain
     * results are bit-identical to the one-device run.  0 or 1: one device (`device`).  An ordinal may repeat
 * (those shards then share that GPU; the tests use it to run this path on a one-GPU box).
     * Env MH_DEVICES ("all" or a comma list) supplies the list when the options carry none -- the way to
     * spread the reference's own entry point, KernelWrapper, over several GPUs. */
    int32_t n_devices;
    int32_t devices[MH_MAX_DEVICES];
    int32_t reserved0;
} mhOptions;
MH_STATIC_ASSERT(sizeof(mhOptions) == 136, "mhOptions");

/* One record per chain per iteration, for trajectory tests (KernelRunTraced). */
typedef struct mhTraceEntry {
    int32_t move;     /* 0 translate, 1 rotate, 2 swap (Kernel.cu:595, 634, 655)       */
    int32_t obj1;     /* moved object, or first swap partner; -1 if no move was made    */
    int32_t obj2;     /* second swap partner, else -1                                   */
    int32_t accepted; /* Accept() outcome (Kernel.cu:819)                               */
    float star_total; /* totalCosts of the proposal                                     */
    float cur_total;  /* totalCosts of the current layout AFTER the accept/reject       */
    float u;          /* the acceptance uniform                                         */
    float beta;       /* beta used by this iteration                                    */
} mhTraceEntry;
MH_STATIC_ASSERT(sizeof(mhTraceEntry) == 32, "mhTraceEntry");

typedef struct mhContext mhContext; /* opaque: device-resident problem + chain state */

/* ---- the reference's entry point -------------------------------------------------------- */

/* Replaces Kernel.cu:873-984.  Array lengths are implicit: rss[R], rsa[R] with
 * R = srf->nRelationships; cfg[n]; clearances[C]; offlimits[n]; vertices[4C+4n];
 * surfaceRectangle[4].  Runs gpuCfg->gridxDim chains of gpuCfg->iterations MH steps from
 * `cfg` and returns a malloc'd array result[gridxDim]; result[i].points points into ONE
 * malloc'd block of gridxDim*n points whose base is result[0].points (Kernel.cu:928, 970,
 * 981).  The caller owns both (free() or KernelFree).  Unlike the reference (quirk Q3) the
 * `costs` of every result are filled in.  NULL on error. */
MH_API result *KernelWrapper(relationshipStruct *rss, relationshipAngleStruct *rsa,
                             positionAndRotation *cfg, rectangle *clearances,
                             rectangle *offlimits, vertex *vertices, vertex *surfaceRectangle,
                             Surface *srf, gpuConfig *gpuCfg);

/* ---- additive entry points ---------------------------------------------------------------- */

/* KernelWrapper with explicit options (seed, schedule, sharding, best tracking). */
MH_API result *KernelWrapperEx(const relationshipStruct *rss, const relationshipAngleStruct *rsa,
                               const positionAndRotation *cfg, const rectangle *clearances,
                               const rectangle *offlimits, const vertex *vertices,
                               const vertex *surfaceRectangle, const Surface *srf,
                               const gpuConfig *gpuCfg, const mhOptions *opt);

/* Frees what KernelWrapper/KernelWrapperEx returned (the reference has no such call and
 * its only caller leaks; free(res[0].points); free(res) is equivalent). */
MH_API void KernelFree(result *res);

/* Message of the last failure on the calling thread ("" if none). */
MH_API const char *KernelLastError(void);

/* Pure cost evaluation on the GPU: the reference's Costs() (Kernel.cu:516-550) applied to
 * nLayouts layouts of n objects each (layouts[l*n + i]; only x, y, rotY vary in practice,
 * length/width/frozen are taken from layout 0).  out[nLayouts].  Returns 0 on success. */
MH_API int KernelEvalCosts(const relationshipStruct *rss, const relationshipAngleStruct *rsa,
                           const positionAndRotation *layouts, int nLayouts,
                           const rectangle *clearances, const rectangle *offlimits,
                           const vertex *vertices, const vertex *surfaceRectangle,
                           const Surface *srf, resultCosts *out);

/* Persistent, device-resident run: replaces the reference's per-call 12x cudaMalloc, H2D
 * staging and curand init (Kernel.cu:879-943) for callers that step the same room many
 * times.  All nChains chains start from `cfg`. */
MH_API mhContext *KernelCreate(const relationshipStruct *rss, const relationshipAngleStruct *rsa,
                               const positionAndRotation *cfg, const rectangle *clearances,
                               const rectangle *offlimits, const vertex *vertices,
                               const vertex *surfaceRectangle, const Surface *srf,
                               int nChains, const mhOptions *opt);
/* Advance every chain by `iterations` MH steps (asynchronous on the context's stream). */
MH_API int KernelRun(mhContext *ctx, int iterations);
/* Same, and record one mhTraceEntry per chain per iteration: trace[it*nChains + chain]
 * (host buffer, iterations*nChains entries).  Synchronous. */
MH_API int KernelRunTraced(mhContext *ctx, int iterations, mhTraceEntry *trace);
/* Wait for the context's stream. */
MH_API int KernelSynchronize(mhContext *ctx);
/* Copy results to host buffers: points[nChains*n], costs[nChains] (either may be NULL). */
MH_API int KernelResults(mhContext *ctx, point *points, resultCosts *costs);
/* Device addresses of the result buffers (valid until KernelDestroy), for zero-copy
 * consumers such as a collective over the per-chain costs.  STREAM ORDERING: the buffers are written by
 * work enqueued on the context's stream (a non-blocking stream of the library's own unless
 * KernelSetStream named another).  A consumer on a different stream must first KernelSynchronize(ctx),
 * or make the context run on its own stream with KernelSetStream.  Not available on a multi-device
 * context (n_devices > 1): there is no single device address. */
MH_API int KernelDeviceResults(mhContext *ctx, void **d_points, void **d_costs);
/* Run on a caller-provided CUDA stream (a cudaStream_t passed as void*).  NULL = a stream of the
 * library's own; to name the legacy default stream pass cudaStreamLegacy, i.e. (void*)0x1. */
MH_API int KernelSetStream(mhContext *ctx, void *stream);
/* Index (local to this context) and totalCosts of the chain with the highest totalCosts
 * (the sampler maximises totalCosts, quirk Q10); reduced on the device. */
MH_API int KernelBest(mhContext *ctx, int *bestChain, float *bestTotal);
/* The k chains with the highest totalCosts, best first (ties: lower chain index first): indices local
 * to this context and their totals (either output may be NULL).  Returns how many were written
 * (min(k, nChains)) or -1.  The ranking the reference leaves to its caller (it cannot rank: its costs
 * are garbage, quirk Q3). */
MH_API int KernelTopK(mhContext *ctx, int k, int *chains, float *totals);
/* The same, keeping only suggestions that DIFFER: greedy by descending totalCosts, a chain is taken if its
 * layout is farther than minDistance from every chain taken before it.  The distance of two layouts is the
 * largest displacement of any object: max_i max(|dx_i|, |dy_i|, rotWeight * |drotY_i| wrapped into
 * [0, 3.1416]); rotWeight = 0 ignores rotations.  Distances and the masked arg-max run on the device
 * (one fused kernel per pick, all k enqueued at once, one read-back; on a multi-device context one small host
 * round trip per pick hands the pick's layout to the other devices); many chains of a converged run are
 * near-duplicates, and the caller wants a handful of different rooms to show.  Returns how many were written
 * (<= k; fewer when everything left is a near-duplicate) or -1. */
MH_API int KernelTopKDistinct(mhContext *ctx, int k, float minDistance, float rotWeight, int *chains, float *totals);
/* Multi-GPU arg-best: writes to the DEVICE address d_key one signed 64-bit key that orders like
 * (totalCosts of this context's best chain, lower global chain id first).  A MAX all-reduce of
 * the keys over all ranks (NCCL has no arg-max) yields the global best; KernelDecodeBestKey
 * recovers the global chain id and its totalCosts.  The key keeps the low 32 bits of the global id:
 * the call fails if this context holds a chain with a global id >= 2^32.
 * STREAM ORDERING: asynchronous on the context's stream -- d_key is valid for work enqueued LATER ON
 * THAT STREAM; a collective on another stream (torch's, NCCL's) must be ordered after it with
 * KernelSynchronize(ctx) or by running the context on that stream (KernelSetStream).  The same holds for
 * the arrays of KernelTemperingState. */
MH_API int KernelBestKey(mhContext *ctx, void *d_key);
MH_API void KernelDecodeBestKey(long long key, unsigned long long *globalChain, float *total);
/* Restart every chain from the caller's layout (iteration counter back to 0).  A tempering context restarts on its
 * CURRENT ladder (the tuned one, if KernelTemperingSetLadder was called). */
MH_API int KernelReset(mhContext *ctx);
/* Cross-GPU replica exchange (extension).  A context created with tempering_rungs > 1 and
 * chain_stride = S > 1 runs its chains at their current betas and never exchanges by itself.
 * Every exchange_interval iterations the caller (1) takes the device arrays of per-chain
 * totalCosts and betas from KernelTemperingState, (2) all-gathers them over the S ranks into
 * rank-major arrays all[rank*nChains + i] (NCCL all_gather does exactly that), and (3) hands them
 * to KernelTemperingExchange, which decides every neighbour swap of this rank's chains from the
 * shared Philox stream -- both members of a pair reach the same decision -- and updates the local
 * betas.  Betas move, layouts never cross the link: 8 bytes per chain per exchange. */
MH_API int KernelTemperingState(mhContext *ctx, void **d_totals, void **d_betas);
MH_API int KernelTemperingExchange(mhContext *ctx, const void *d_all_totals, const void *d_all_betas);
/* Replica-exchange bookkeeping since creation / the last KernelReset, for tuning a ladder: for every pair of
 * neighbouring rungs (r, r+1), r = 0 .. rungs-2, how many exchanges THIS context's chains attempted as the
 * pair's lower member and how many were accepted (sum over the ranks for a ladder spread over GPUs).  A
 * pair that hardly ever swaps is a gap in the ladder; one that always swaps is a rung too many.  Returns
 * rungs-1 or -1. */
MH_API int KernelTemperingStats(mhContext *ctx, long long *attempts, long long *accepted);
/* Ladder tuning (the policy layer on top of the statistics).  The reference has one fixed
 * BETA (Kernel.cu:33); a tempering ladder only works when neighbouring rungs exchange often enough, and the
 * geometric ladder mhOptions describes is a starting guess.
 *   KernelTemperingLadder        the context's current ladder, betas[rungs], rung 0 first.  Returns rungs or -1.
 *   KernelTemperingProposeLadder a pure function, no context: from the current ladder and the exchange counts of
 *                                every pair of neighbouring rungs (summed over all ranks when the ladder is spread
 *                                over GPUs) it places the interior rungs so that every pair is expected to exchange
 *                                equally often: the gap (in log beta) of pair r is taken to carry a "distance"
 *                                -log(acceptance_r) (a pair that always swaps is close, one that never swaps is
 *                                far; rates are clamped to [0.01, 0.99], pairs with fewer than 8 attempts count as
 *                                0.5), and the new rungs sit at equal steps of the cumulative distance, linearly in
 *                                log beta inside a gap.  End points stay.  `damping` in (0, 1] moves only that
 *                                fraction of the way (1 = all the way).  Returns 0 or -1.
 *   KernelTemperingSetLadder     re-targets every chain from the current ladder to betas[rungs]: a chain that
 *                                sits on rung r (whichever chain that currently is: exchanges permute them) moves
 *                                to betas[r]; clears the exchange statistics.  Asynchronous on the context's stream.
 * Changing the ladder mid-run breaks detailed balance for that step: adapt during burn-in, then leave it. */
MH_API int KernelTemperingLadder(mhContext *ctx, double *betas);
MH_API int KernelTemperingProposeLadder(int rungs, const double *current, const long long *attempts, const long long *accepted,
                                        double damping, double *proposed);
MH_API int KernelTemperingSetLadder(mhContext *ctx, const double *betas);

/* Milliseconds the device spent in the MH kernels since creation (CUDA events) and how many
 * kernels were launched.  Multi-device context: the devices run concurrently, kernel_ms is the largest
 * per-device sum and launches the total over the devices. */
MH_API int KernelStats(mhContext *ctx, double *kernel_ms, long long *launches);
MH_API void KernelDestroy(mhContext *ctx);

/* Device memory is drawn from a pool the library keeps between calls (the reference re-allocates
 * and frees 12 buffers per call, Kernel.cu:879-967).  KernelTrim returns the cached blocks of
 * the current device to the driver. */
MH_API int KernelTrim(void);

/* How the context was laid out: lanes per chain of the chain kernel (the lane width fixes the float
 * reduction order: contexts compare bit for bit only at equal width), the evaluation form actually run
 * (MH_EVAL_FULL_SCAN, MH_EVAL_MEMO or MH_EVAL_DELTA), the number of devices, and each device's ordinal and
 * chain count (arrays of MH_MAX_DEVICES entries; any pointer may be NULL).  Returns 0 or -1. */
MH_API int KernelShape(mhContext *ctx, int *lanesPerChain, int *evalForm, int *nDevices, int *deviceOrdinals, int *chainsPerDevice);

/* Number of CUDA devices visible to the process (-1 on error). */
MH_API int KernelDeviceCount(void);

/* Library / device facts for harnesses: returns 0 and fills what is non-NULL. */
MH_API int KernelDeviceInfo(int *smCount, int *smClockKHz, int *ccMajor, int *ccMinor,
                            char *name, int nameLen);

#ifdef __cplusplus
}
#endif
#endif /* MH_KERNEL_H */
