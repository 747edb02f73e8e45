// mh_kernels.cu -- sm_100a kernels of the Metropolis-Hastings layout optimiser and the
// device half of the thin C ABI (mh_abi.h).
//
// mh_chain_kernel<G>   the hot path: one group of G lanes runs one chain (replaces the
//                      reference's one-block-per-chain Kernel, Kernel.cu:754-871): in-register
//                      Philox proposals, full cost re-evaluation from shared memory, the
//                      exp(beta dE) accept test, optional best-layout tracking, coalesced
//                      write-out.  No block-level barrier inside the iteration loop.
// mh_delta_kernel<G>   the same chain with incremental evaluation (MH_EVAL_DELTA, mh_delta.cuh).
// mh_score_kernel<G>   all eight cost terms of given layouts (fills resultCosts, which the
//                      reference forgets -- quirk Q3; also the KernelEvalCosts parity hook).
// mh_exchange_kernel   replica exchange between neighbouring temperature rungs (extension).
// mh_argmax_kernel, mh_topk_stage_kernel, mh_distinct_round_kernel, mh_bestkey_kernel
//                      ranking of a context's chains on the device (best, top-k, distinct top-k, the packed
//                      key of the multi-GPU arg-best).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "mh_abi.h"
#include "mh_costs.cuh"
#include "mh_delta.cuh"
#include "philox.cuh"

namespace mh {

constexpr int WARPS_PER_BLOCK = 4;
constexpr int THREADS = WARPS_PER_BLOCK * 32;
#ifndef MH_MIN_BLOCKS
#define MH_MIN_BLOCKS 4 // resident 128-thread blocks per SM the chain kernel is compiled for: <= 128 registers (measured on B200 with the integer
                        // clearance sum: 4 blocks / 123 registers, no spill, beat 5 blocks / 96 registers (20-byte spill) by 7 % at 8 objects, 12 % at 16, 16 % at 24)
#endif

struct PointRec {
    float x, y, z, rotX, rotY, rotZ;
};
struct TraceRec {
    int32_t move, obj1, obj2, accepted;
    float star_total, cur_total, u, beta;
};

__device__ __forceinline__ void stage_problem(float *smem, const float *g, int words)
{
    // words is a multiple of 4 and both sides are 16-byte aligned
    const float4 *src = reinterpret_cast<const float4 *>(g);
    float4 *dst = reinterpret_cast<float4 *>(smem);
    for (int i = threadIdx.x; i < words / 4; i += blockDim.x)
        dst[i] = __ldg(src + i);
    __syncthreads();
}

// Write one chain's layout as point records: every lane of the warp takes part and the 6n
// floats of the chain leave as consecutive 8-byte stores (256 B per warp instruction).
// Kernel.cu:706-713: u < min(1, exp(beta (star - cur))), the exponential in double like the reference.
__device__ __forceinline__ bool accept_move(float u, float beta, float star, float cur)
{
    return u < fminf(1.0f, (float)exp((double)beta * ((double)star - (double)cur)));
}

// u < min(1, exp(x)), x = beta (star - cur) in double (accept_move), decided from a float estimate of the
// exponential whenever u is clear of the threshold by more than the estimate's error; the rare
// in-between case (probability ~2e-4) takes the double-precision exponential.  Same decisions, always.
__device__ __forceinline__ bool accept_move_fast(float u, float beta, float star, float cur)
{
    const float ef = __expf(beta * (star - cur));                // relative error < 4e-5 for |x| <= 100
    if (u < fminf(1.0f, ef * 0.9999f)) return true;
    if (u >= ef * 1.0001f) return false;
    return accept_move(u, beta, star, cur);
}

template <class WS>
__device__ __forceinline__ void write_points_warp(const WS &S, int cc, int n, const float *pass, const uint16_t *perm,
                                                  PointRec *out, int lane)
{
    float2 *o2 = reinterpret_cast<float2 *>(out);
    for (int q = lane; q < 3 * n; q += 32) {
        const int i = q / 3, part = q - 3 * i;
        const int src = perm[i];
        float2 v;
        const float4 p = S.P4[WS::at(i, cc)];
        if (part == 0) v = make_float2(p.x, p.y);
        else if (part == 1) v = make_float2(pass[src], pass[n + src]);           // z, rotX travel with swaps
        else v = make_float2(p.z, pass[2 * n + src]);                            // rotY, rotZ
        o2[q] = v;
    }
}

// Every proposal re-evaluates every live cost term from scratch (Kernel.cu:804).  (The forms that do
// less work for the same or a statistically equivalent result, MH_EVAL_MEMO and MH_EVAL_DELTA, are
// mh_delta_kernel below.)
template <int G>
__global__ void __launch_bounds__(THREADS, MH_MIN_BLOCKS) mh_chain_kernel(const __grid_constant__ mhLaunch L)
{
    using WS = WarpState<G, true>;
    constexpr int CPW = WS::CPW;
    extern __shared__ __align__(16) float smem[];
    const float *gprob = static_cast<const float *>(L.d_problem);
    stage_problem(smem, gprob, L.smem_words);
    const SmemProblem P = bind_problem(smem, &L.hdr);
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C;

    using LM = LaneMap<G, false>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = LM::chain(lane), g = LM::lane_in_group(lane);
    WS S;
    S.bind(smem + L.smem_words + warp * WS::words(n, C), n, C);

    const int chain_raw = (blockIdx.x * WARPS_PER_BLOCK + warp) * CPW + c;
    const bool live = chain_raw < L.n_chains;
    const int chain = live ? chain_raw : L.n_chains - 1;   // idle groups shadow the last chain, write nothing
    const uint64_t gchain = L.chain_offset + (uint64_t)chain * L.chain_stride;
    const float *cfg0 = gprob + h->off_cfg0;
    const float *pass = gprob + h->off_pass;
    PointRec *points = static_cast<PointRec *>(L.d_points);

    // ---- load chain state -----------------------------------------------------------------------
    for (int i = g; i < n; i += G) {
        float x, y, r;
        if (L.fresh) {
            x = cfg0[i]; y = cfg0[n + i]; r = cfg0[2 * n + i];
            if (live) L.d_perm[(size_t)chain * n + i] = (uint16_t)i;
        } else {
            const size_t o = (size_t)chain * n + i;
            x = L.d_x[o]; y = L.d_y[o]; r = L.d_rot[o];
        }
        S.P4[WS::at(i, c)] = make_float4(x, y, r, focal_cos(h, x, y, r));
    }
    __syncwarp();

    float cur, best;
    if (L.fresh) {
        RawTerms t;
        eval_terms<G, false>(P, S, c, g, t);
        cur = combine(h, t).total;                            // Kernel.cu:778
        best = cur;
        if (L.result_mode == 1) {
            __syncwarp();
            for (int cc = 0; cc < CPW; cc++) {
                const int ch = (blockIdx.x * WARPS_PER_BLOCK + warp) * CPW + cc;
                if (ch < L.n_chains) write_points_warp(S, cc, n, pass, L.d_perm + (size_t)ch * n, points + (size_t)ch * n, lane);
            }
        }
    } else {
        cur = L.d_cur_total[chain];
        best = L.d_best_total[chain];
    }

    const float room_x0 = h->room_minx, room_y0 = h->room_miny, room_x1 = h->room_maxx, room_y1 = h->room_maxy;
    const int any_free = h->any_free;
    TraceRec *trace = static_cast<TraceRec *>(L.d_trace);
    float beta = L.beta_start;
    if (L.schedule == MH_SCHED_PER_CHAIN) beta = L.d_beta[chain];

    int pa_mine = 0, b_mine = -1;                               // this lane's share of the current batch of proposal recipes
    float n0_mine = 0.f, n1_mine = 0.f, u_mine = 0.f;

    // ---- the chain (Kernel.cu:785-827, Semantics S of SURVEY.md section 8a) ---------------------
    for (int k = 0; k < L.it_count; k++) {
        const uint64_t it = L.it_begin + (uint64_t)k;
        if (L.schedule == MH_SCHED_GEOMETRIC || L.schedule == MH_SCHED_LINEAR) {
            const int len = L.schedule_length;
            const uint64_t ic = it < (uint64_t)(len - 1) ? it : (uint64_t)(len - 1);
            const float tt = len > 1 ? (float)((double)ic / (double)(len - 1)) : 0.f;
            beta = L.schedule == MH_SCHED_GEOMETRIC ? L.beta_start * exp2f(tt * L.beta_log2_ratio)
                                                    : L.beta_start + (L.beta_end - L.beta_start) * tt;
        }

        // -- propose (Kernel.cu:576-704).  The stream is counter-based, so the G lanes of a group draw the proposal
        //    recipes of G DIFFERENT iterations at once (lane g: iteration k0 + g -- Philox blocks 0 and 1, the re-draws
        //    while a picked object is frozen, Box-Muller) and every iteration fetches its recipe with five shuffles:
        //    the same draws, off the per-iteration path on G - 1 of G iterations (see mh_delta_kernel).  One lane per
        //    chain: nothing to share, every iteration draws for itself. --
        int p, a = -1, b = -1;
        float n0, n1, u;
        if (G == 1 || (k & (G - 1)) == 0) {                    // (uniform over the warp)
            const uint64_t itb = it + (uint64_t)g;
            const Philox4 wb = draw_block(L.seed, gchain, itb, 0);
            u_mine = uniform01(draw_block(L.seed, gchain, itb, 1).x);
            const int pb = random_int(uniform01(wb.x), 2);
            int ab = -1, bb = -1;
            if (any_free && (pb != 2 || n >= 2)) {
                uint32_t redraw = 2;
                ab = random_int(uniform01(wb.y), n - 1);
                if (pb == 2) bb = random_int(uniform01(wb.z), n - 1);
                while (P.obj_frozen[ab] || (bb >= 0 && P.obj_frozen[bb])) {       // Kernel.cu:601, 637, 662, 666
                    const Philox4 rw = draw_block(L.seed, gchain, itb, redraw++);
                    if (P.obj_frozen[ab]) ab = random_int(uniform01(rw.x), n - 1);
                    if (bb >= 0 && P.obj_frozen[bb]) bb = random_int(uniform01(rw.y), n - 1);
                }
            }
            box_muller(wb.z, wb.w, n0_mine, n1_mine);
            pa_mine = pb | ((ab + 1) << 2);
            b_mine = bb;
        }
        if (G == 1) {
            p = pa_mine & 3; a = (pa_mine >> 2) - 1; b = b_mine; n0 = n0_mine; n1 = n1_mine; u = u_mine;
        } else {
            const int src_lane = LM::first_lane(c) + (k & (G - 1));     // the lane of this group that drew for iteration k
            const int pa = __shfl_sync(0xffffffffu, pa_mine, src_lane);
            p = pa & 3; a = (pa >> 2) - 1;
            b = __shfl_sync(0xffffffffu, b_mine, src_lane);
            n0 = __shfl_sync(0xffffffffu, n0_mine, src_lane);
            n1 = __shfl_sync(0xffffffffu, n1_mine, src_lane);
            u = __shfl_sync(0xffffffffu, u_mine, src_lane);
        }
        float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb = na;   // proposed state of objects a, b
        float4 oa = na, ob = na;                                // state to restore on rejection
        if (a >= 0) {
            MH_CHECK(a >= 0 && a < n && b >= -1 && b < n);
            oa = S.P4[WS::at(a, c)];
            na = oa;
            if (p == 0) {                                      // translate, sigma = W/16, H/16 (Q19), snap to the room
                const float nx = oa.x + n0 * h->std_x, ny = oa.y + n1 * h->std_y;
                na.x = nx > room_x1 ? room_x1 : (nx < room_x0 ? room_x0 : nx);
                na.y = ny > room_y1 ? room_y1 : (ny < room_y0 ? room_y0 : ny);
                na.w = focal_cos(h, na.x, na.y, na.z);
            } else if (p == 1) {                               // rotate, one wrap into [0, 2 PI] (Kernel.cu:645-651)
                float ar = oa.z + n0 * h->sigma_t;
                if (ar < 0.f) ar += h->two_pi;
                else if (ar > h->two_pi_cmp) ar -= h->two_pi;
                na.z = ar;
                na.w = focal_cos(h, na.x, na.y, na.z);
            } else {                                           // swap position and rotation (Kernel.cu:675-700)
                ob = S.P4[WS::at(b, c)];
                na = ob;                                       // the memoised cosine travels with (x, y, rot)
                nb = oa;
            }
        }
        __syncwarp();
        if (g == 0 && a >= 0) {
            S.P4[WS::at(a, c)] = na;
            if (b >= 0) S.P4[WS::at(b, c)] = nb;
        }
        __syncwarp();

        // -- evaluate the proposal (Kernel.cu:804): every live term, from scratch ---------------------
        float star;
        {
            RawTerms t;
            eval_terms<G, false>(P, S, c, g, t);
            star = combine(h, t).total;
        }

        // -- accept (Kernel.cu:706-713): u < min(1, exp(beta (star - cur))), maximises (Q10) ------
        const bool acc = accept_move_fast(u, beta, star, cur);
        __syncwarp();
        if (acc) {
            cur = star;
            if (g == 0 && b >= 0 && live) {                    // z, rotX, rotZ travel with the swap: the
                uint16_t *pm = L.d_perm + (size_t)chain * n;    // permutation lives in global memory, touched
                const uint16_t pa = pm[a];                      // only by accepted swaps and by the write-out
                pm[a] = pm[b];
                pm[b] = pa;
            }
        } else if (g == 0 && a >= 0) {
            S.P4[WS::at(a, c)] = oa;
            if (b >= 0) S.P4[WS::at(b, c)] = ob;
        }
        __syncwarp();
        if (L.result_mode == 1) {
            // best layout so far = highest totalCosts; it can only improve on an accepted move
            const bool improved = acc && cur > best;
            if (improved) best = cur;
            const unsigned mask = __ballot_sync(0xffffffffu, improved && live);
            if (mask) {
                for (int cc = 0; cc < CPW; cc++)
                    if (mask & (1u << LM::first_lane(cc))) {
                        const int ch = (blockIdx.x * WARPS_PER_BLOCK + warp) * CPW + cc;
                        write_points_warp(S, cc, n, pass, L.d_perm + (size_t)ch * n, points + (size_t)ch * n, lane);
                    }
            }
        }
        if (trace && live && g == 0) {
            TraceRec r;
            r.move = p; r.obj1 = a; r.obj2 = b; r.accepted = acc ? 1 : 0;
            r.star_total = star; r.cur_total = cur; r.u = u; r.beta = beta;
            trace[(size_t)k * L.n_chains + chain] = r;
        }
    }

    // ---- persist the chain, emit the result ------------------------------------------------------
    __syncwarp();
    if (live) {
        for (int i = g; i < n; i += G) {
            const size_t o = (size_t)chain * n + i;
            const float4 p = S.P4[WS::at(i, c)];
            L.d_x[o] = p.x;
            L.d_y[o] = p.y;
            L.d_rot[o] = p.z;
        }
        if (g == 0) {
            L.d_cur_total[chain] = cur;
            L.d_best_total[chain] = best;
        }
    }
    if (L.result_mode == 0) {                                  // Kernel.cu:834-842: the final current layout
        for (int cc = 0; cc < CPW; cc++) {
            const int ch = (blockIdx.x * WARPS_PER_BLOCK + warp) * CPW + cc;
            if (ch < L.n_chains) write_points_warp(S, cc, n, pass, L.d_perm + (size_t)ch * n, points + (size_t)ch * n, lane);
        }
    }
}

// The chain of mh_chain_kernel with incremental evaluation (mh_delta.cuh).  Differences in shape:
//  * the per-iteration code path is kept inside the 32 KB instruction cache (rolled loops, a rolled
//    Philox, helpers out of line): with ~20 warps per SM each at its own place in the iteration, a longer
//    path makes the warps evict each other's lines (ncu: stall_no_instruction 6.6 per issue on the first
//    version; a block barrier per iteration cured that too, but made every warp wait for the slowest);
//  * blocks of 4 or 8 warps (blockDim.x), 128 registers;
//  * no branch on the move type: translate / rotate / swap are computed side by side and selected,
//    so the chains of a warp do not serialise;
//  * the two Philox blocks of an iteration are computed by different lanes of the group at once.
// MODE = kModeDelta: MH_EVAL_DELTA (delta_eval: running float sums, statistically equivalent to full evaluation);
// MODE = kModeExact: MH_EVAL_MEMO (exact_eval: every total bit-identical to the full evaluation's).
// Neither keeps the plain scan's per-warp array of clearance rectangles: the clearance term is an integer sum
// updated pair by pair (mh_costs.cuh), the old rectangles are rebuilt from the moved objects' old positions.
// Two builds of every memo kernel: WPB = 8 for blocks of up to 8 warps, two per SM (128 registers per thread), and
// WPB = 4 for blocks of 4 warps, five per SM (96 registers: no spill since the header's scalars come from the constant
// bank).  Which one runs is the host's choice (choose_delta_shape: whichever keeps more warps resident given the
// shared memory a chain needs): at 50 objects 5 x 4 warps beat 2 x 8 by 5 %, at 200 objects they lose 19 %.
#define MH_DELTA_THREADS 256
template <int G, int MODE, int WPB>
__global__ void __launch_bounds__(WPB * 32, WPB == 4 ? 5 : 2) mh_delta_kernel(const __grid_constant__ mhLaunch L)
{
    using WS = WarpState<G>;
    using DS = DeltaState<G>;
    using LM = LaneMap<G, kDeltaStr>;
    constexpr int CPW = WS::CPW;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool EXACT = MODE != kModeDelta;
    extern __shared__ __align__(16) float smem[];
    const float *gprob = static_cast<const float *>(L.d_problem);
    stage_problem(smem, gprob, L.smem_words);
    const SmemProblem P = bind_problem(smem, &L.hdr);
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C;
    const int warps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = LM::chain(lane), g = LM::lane_in_group(lane);
    WS S;
    DS D;
    {
        float *base = smem + L.smem_words + warp * (WS::words(n, 0) + DS::words(n, C, h->R, MODE));
        S.bind(base, n, 0);                                     // no clearance-rectangle array in this kernel
        D.bind(base + WS::words(n, 0), n, C, h->R);
    }
    const int chain0 = (blockIdx.x * warps + warp) * CPW;     // first chain of this warp
    const bool live = chain0 + c < L.n_chains;
    const int chain = live ? chain0 + c : L.n_chains - 1;     // idle groups shadow the last chain, write nothing
    const uint64_t gchain = L.chain_offset + (uint64_t)chain * L.chain_stride;
    const float *cfg0 = gprob + h->off_cfg0;
    const float *pass = gprob + h->off_pass;
    PointRec *points = static_cast<PointRec *>(L.d_points);

    for (int i = g; i < n; i += G) {
        float x, y, r;
        if (L.fresh) {
            x = cfg0[i]; y = cfg0[n + i]; r = cfg0[2 * n + i];
            if (live) L.d_perm[(size_t)chain * n + i] = (uint16_t)i;
        } else {
            const size_t o = (size_t)chain * n + i;
            x = L.d_x[o]; y = L.d_y[o]; r = L.d_rot[o];
        }
        S.P4[WS::at(i, c)] = make_float4(x, y, r, focal_cos(h, x, y, r));
    }
    __syncwarp();

    RunSums sums;
    int sel = 0;
    float cur = delta_rebuild<G, MODE>(P, S, D, c, g, sel, sums);   // memos (and running sums) of the current layout; = Kernel.cu:778
    float best = L.fresh ? cur : L.d_best_total[chain];
    if (L.fresh && L.result_mode == 1) {
        for (int cc = 0; cc < CPW; cc++)
            if (chain0 + cc < L.n_chains)
                write_points_warp(S, cc, n, pass, L.d_perm + (size_t)(chain0 + cc) * n, points + (size_t)(chain0 + cc) * n, lane);
    }

    const float room_x0 = h->room_minx, room_y0 = h->room_miny, room_x1 = h->room_maxx, room_y1 = h->room_maxy;
    const int any_free = h->any_free;
    TraceRec *trace = static_cast<TraceRec *>(L.d_trace);
    float beta = L.beta_start;
    if (L.schedule == MH_SCHED_PER_CHAIN) beta = L.d_beta[chain];
    int recipe_mine = 0;                                        // this lane's share of the current batch of proposal recipes
    float n0_mine = 0.f, n1_mine = 0.f, u_mine = 0.f;

    for (int k = 0; k < L.it_count; k++) {
        const uint64_t it = L.it_begin + (uint64_t)k;
        if (L.schedule == MH_SCHED_GEOMETRIC || L.schedule == MH_SCHED_LINEAR) {
            const int len = L.schedule_length;
            const uint64_t ic = it < (uint64_t)(len - 1) ? it : (uint64_t)(len - 1);
            const float tt = len > 1 ? (float)((double)ic / (double)(len - 1)) : 0.f;
            beta = L.schedule == MH_SCHED_GEOMETRIC ? L.beta_start * exp2f(tt * L.beta_log2_ratio)
                                                    : L.beta_start + (L.beta_end - L.beta_start) * tt;
        }
        if (!EXACT && k > 0 && (it % (uint64_t)kRefresh) == 0)   // bound the drift of the running sums
            cur = delta_rebuild<G, kModeDelta>(P, S, D, c, g, sel, sums);

        // -- random numbers.  The stream is counter-based: what iteration k draws does not depend on what happened
        //    before it.  So the G lanes of a group draw for G DIFFERENT iterations at once -- lane g computes the
        //    whole proposal recipe of iteration k0 + g (Philox blocks 0 and 1, the re-draws while a picked object is
        //    frozen, Box-Muller) -- and every iteration fetches its recipe from the lane that holds it with four
        //    shuffles.  Same draws as computing them per iteration (same counters), but Philox, logf / sqrtf /
        //    sincospif and the integer draws leave the per-iteration path on G - 1 of G iterations, and with them
        //    the longest serial dependency chain before the evaluation can start. --
        constexpr int B = G;                                    // iterations per batch (a power of two)
        const int kb = k & (B - 1);
        if (kb == 0) {                                          // (uniform over the warp)
            const uint64_t itb = it + (uint64_t)g;
            const Philox4 wb = draw_block<2>(L.seed, gchain, itb, 0);
            u_mine = uniform01(draw_block<2>(L.seed, gchain, itb, 1).x);
            const int pb = random_int(uniform01(wb.x), 2);
            int ab = -1, bb = -1;
            if (any_free && (pb != 2 || n >= 2)) {
                uint32_t redraw = 2;
                ab = random_int(uniform01(wb.y), n - 1);
                if (pb == 2) bb = random_int(uniform01(wb.z), n - 1);
                while (P.obj_frozen[ab] || (bb >= 0 && P.obj_frozen[bb])) {          // Kernel.cu:601, 637, 662, 666
                    const Philox4 rw = draw_block<2>(L.seed, gchain, itb, redraw++);
                    if (P.obj_frozen[ab]) ab = random_int(uniform01(rw.x), n - 1);
                    if (bb >= 0 && P.obj_frozen[bb]) bb = random_int(uniform01(rw.y), n - 1);
                }
            }
            box_muller(wb.z, wb.w, n0_mine, n1_mine);
            recipe_mine = pb | ((ab + 1) << 2) | ((bb + 1) << 14);      // n <= 32 G <= 1024 in this kernel: 12 bits each
        }
        const int src_lane = LM::first_lane(c) + kb * LM::xor_step;    // the lane of this group that drew for iteration k
        const int recipe = __shfl_sync(FULL, recipe_mine, src_lane);
        const float n0 = __shfl_sync(FULL, n0_mine, src_lane), n1 = __shfl_sync(FULL, n1_mine, src_lane);
        const float u = __shfl_sync(FULL, u_mine, src_lane);

        // -- propose (Kernel.cu:576-704), the three moves side by side ----------------------------------
        const int p = recipe & 3, a = ((recipe >> 2) & 0xFFF) - 1, b = ((recipe >> 14) & 0xFFF) - 1;
        const bool moved = a >= 0;
        MH_CHECK(a >= -1 && a < n && b >= -1 && b < n && (b < 0 || a >= 0));
        const int a_e = moved ? a : 0;                           // no move: "move" object 0 onto itself (all deltas are 0)
        const float4 oa = S.P4[WS::at(a_e, c)];
        const float4 ob = S.P4[WS::at(b >= 0 ? b : a_e, c)];
        float4 na = oa, nb = oa;
        {
            const float nx = oa.x + n0 * h->std_x, ny = oa.y + n1 * h->std_y;   // translate (Q19), snapped to the room
            const float tx = nx > room_x1 ? room_x1 : (nx < room_x0 ? room_x0 : nx);
            const float ty = ny > room_y1 ? room_y1 : (ny < room_y0 ? room_y0 : ny);
            float ar = oa.z + n0 * h->sigma_t;                                   // rotate, one wrap (Kernel.cu:645-651)
            if (ar < 0.f) ar += h->two_pi;
            else if (ar > h->two_pi_cmp) ar -= h->two_pi;
            if (moved && p == 0) { na.x = tx; na.y = ty; }
            if (moved && p == 1) na.z = ar;
            const float fc = focal_cos(h, na.x, na.y, na.z);
            if (moved && p < 2) na.w = fc;
            if (p == 2 && moved) { na = ob; nb = oa; }                           // swap (Kernel.cu:675-700)
        }
        const int b_eff = (b == a) ? -1 : b;                     // a swap of an object with itself moves nothing twice
        __syncwarp();
        if (g == 0 && moved) {
            S.P4[WS::at(a, c)] = na;
            if (b >= 0) S.P4[WS::at(b, c)] = nb;
        }
        __syncwarp();

        // -- evaluate: only what the moved objects touch (EXACT: the additive terms from scratch) ---------
        RunSums star_sums;
        RelStash stash;
        ExactStash<G> xstash;
        float star;
        if (EXACT) star = exact_eval<G>(P, S, D, c, g, sel, a_e, b_eff, oa, ob, na, nb, sums.clr_q, star_sums.clr_q, xstash);
        else star = delta_eval<G>(P, S, D, c, g, sel, a_e, b_eff, oa, ob, na, nb, sums, star_sums, stash);

        // -- accept (Kernel.cu:706-713) -------------------------------------------------------------------
        const bool acc = accept_move_fast(u, beta, star, cur);
        __syncwarp();
        if (acc) {
            cur = star;
            sel ^= 1;
            if (EXACT) {
                sums.clr_q = star_sums.clr_q;
            } else {
                sums = star_sums;
                delta_commit<G>(P, S, D, c, g, a_e, b_eff, stash);
            }
            if (g == 0 && b >= 0 && live) {                      // z, rotX, rotZ travel with the swap
                uint16_t *pm = L.d_perm + (size_t)chain * n;
                const uint16_t pa = pm[a];
                pm[a] = pm[b];
                pm[b] = pa;
            }
        } else if (g == 0 && moved) {
            S.P4[WS::at(a, c)] = oa;
            if (b >= 0) S.P4[WS::at(b, c)] = ob;
        }
        __syncwarp();
        if (EXACT && !acc) exact_reject<G>(P, S, D, c, g, a_e, b_eff, xstash);
        if (L.result_mode == 1) {
            const bool improved = acc && cur > best;
            if (improved) best = cur;
            const unsigned mask = __ballot_sync(FULL, improved && live);
            if (mask) {
                for (int cc = 0; cc < CPW; cc++)
                    if (mask & (1u << LM::first_lane(cc)))
                        write_points_warp(S, cc, n, pass, L.d_perm + (size_t)(chain0 + cc) * n, points + (size_t)(chain0 + cc) * n, lane);
            }
        }
        if (trace && live && g == 0) {
            TraceRec r;
            r.move = p; r.obj1 = a; r.obj2 = b; r.accepted = acc ? 1 : 0;
            r.star_total = star; r.cur_total = cur; r.u = u; r.beta = beta;
            trace[(size_t)k * L.n_chains + chain] = r;
        }
    }

    __syncwarp();
    if (live) {
        for (int i = g; i < n; i += G) {
            const size_t o = (size_t)chain * n + i;
            const float4 p = S.P4[WS::at(i, c)];
            L.d_x[o] = p.x;
            L.d_y[o] = p.y;
            L.d_rot[o] = p.z;
        }
        if (g == 0) {
            L.d_cur_total[chain] = cur;
            L.d_best_total[chain] = best;
        }
    }
    if (L.result_mode == 0) {
        for (int cc = 0; cc < CPW; cc++)
            if (chain0 + cc < L.n_chains)
                write_points_warp(S, cc, n, pass, L.d_perm + (size_t)(chain0 + cc) * n, points + (size_t)(chain0 + cc) * n, lane);
    }
}

// resultCosts of given layouts: loads x, y, rotY from point records and evaluates all eight terms.
template <int G>
__global__ void __launch_bounds__(THREADS) mh_score_kernel(const float *__restrict__ gprob, int smem_words, int n_layouts,
                                                           const PointRec *__restrict__ points, Costs8 *__restrict__ costs)
{
    using WS = WarpState<G, true>;
    constexpr int CPW = WS::CPW;
    extern __shared__ __align__(16) float smem[];
    stage_problem(smem, gprob, smem_words);
    const SmemProblem P = bind_problem(smem);
    const int n = P.h->n, C = P.h->C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = lane / G, g = lane % G;
    WS S;
    S.bind(smem + smem_words + warp * WS::words(n, C), n, C);
    const int raw = (blockIdx.x * WARPS_PER_BLOCK + warp) * CPW + c;
    const bool live = raw < n_layouts;
    const int l = live ? raw : n_layouts - 1;
    for (int i = g; i < n; i += G) {
        const PointRec pr = points[(size_t)l * n + i];
        S.P4[WS::at(i, c)] = make_float4(pr.x, pr.y, pr.rotY, focal_cos(P.h, pr.x, pr.y, pr.rotY));
    }
    __syncwarp();
    RawTerms t;
    eval_terms<G, true>(P, S, c, g, t);
    const Costs8 r = combine(P.h, t);
    if (live && g == 0) costs[l] = r;
}

// Replica exchange (extension; the reference has a single fixed BETA).  One thread per local
// chain.  Pairs (r, r+1) with r = epoch parity, r+1 < rungs; both members evaluate the same
// decision from the gathered totals/betas, so no communication beyond the gather is needed.
__global__ void mh_exchange_kernel(int n_chains, uint64_t chain_offset, uint64_t chain_stride, int rungs, uint64_t epoch,
                                   uint64_t it_last, uint64_t seed, const float *__restrict__ all_total,
                                   const float *__restrict__ all_beta, uint64_t gather_base, uint64_t gather_stride,
                                   uint64_t gather_local, float *__restrict__ beta, unsigned long long *__restrict__ stats)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains) return;
    const uint64_t gch = chain_offset + (uint64_t)i * chain_stride;
    const int r = (int)(gch % (uint64_t)rungs);
    const int par = (int)(epoch & 1);
    int lo_r;
    if (((r - par) & 1) == 0 && r >= par) lo_r = r;     // this chain is the lower member of its pair
    else lo_r = r - 1;
    if (lo_r < par || lo_r + 1 >= rungs) return;
    const uint64_t glo = gch - (uint64_t)(r - lo_r), ghi = glo + 1;
    const float u = uniform01(draw_block(seed, glo, it_last, 0xFFFFu).x);
    const uint64_t ilo = ((glo - gather_base) % gather_stride) * gather_local + (glo - gather_base) / gather_stride;
    const uint64_t ihi = ((ghi - gather_base) % gather_stride) * gather_local + (ghi - gather_base) / gather_stride;
    const double Ea = -(double)all_total[ilo], Eb = -(double)all_total[ihi];
    const double ba = (double)all_beta[ilo], bb = (double)all_beta[ihi];
    const float pacc = fminf(1.0f, (float)exp((ba - bb) * (Ea - Eb)));
    if (u < pacc) beta[i] = (float)(r == lo_r ? bb : ba);
    if (stats && r == lo_r) {                                   // the pair's lower member keeps the books: {attempts, accepted}
        atomicAdd(&stats[2 * lo_r], 1ull);
        if (u < pacc) atomicAdd(&stats[2 * lo_r + 1], 1ull);
    }
}

// Ladder tuning (KernelTemperingSetLadder): a chain sitting on rung r of the old ladder moves to rung r of the new one.
__global__ void mh_retarget_kernel(int n_chains, int rungs, const float *__restrict__ old_ladder, const float *__restrict__ new_ladder,
                                   float *__restrict__ beta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains) return;
    const float b = beta[i];
    int best = 0;
    float bd = fabsf(b - old_ladder[0]);
    for (int r = 1; r < rungs; r++) {
        const float d = fabsf(b - old_ladder[r]);
        if (d < bd) { bd = d; best = r; }
    }
    beta[i] = new_ladder[best];
}

// ---- ranking on the device -----------------------------------------------------------------------------
// A chain's rank key: orderable(totalCosts) << 32 | (0xFFFFFFFF - chain).  An unsigned MAX over keys is the
// arg-max of totalCosts (the sampler maximises it, quirk Q10) with ties going to the lower chain index;
// 0 means "no chain".  -0.0 ranks as +0.0, a NaN below everything.
__device__ __forceinline__ unsigned long long rank_key(float total, int chain)
{
    uint32_t u = __float_as_uint(total + 0.0f);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if (total != total) u = 1u;
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)chain);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, m);
        v = o > v ? o : v;
    }
    return v;
}

// Block maximum -> one atomicMax on *dst (256-thread blocks).
__device__ __forceinline__ void block_max_to(unsigned long long v, unsigned long long *dst)
{
    __shared__ unsigned long long sm[8];
    v = warp_max_u64(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0ull;
        v = warp_max_u64(v);
        if (threadIdx.x == 0 && v) atomicMax(dst, v);
    }
}

// arg-max of totalCosts over a context's chains: many blocks, one atomic per block (the round-1 kernel was a
// single 1024-thread block: 39 us at 65536 chains, 150 us at 262144, on the time-to-best-cost polling path).
__global__ void __launch_bounds__(256) mh_argmax_kernel(const Costs8 *__restrict__ costs, int n, unsigned long long *key)
{
    unsigned long long best = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long k = rank_key(costs[i].total, i);
        best = k > best ? k : best;
    }
    block_max_to(best, key);
}

// Top-k by rank key.  One stage: every block sorts a tile of kTopkTile keys (bitonic, shared memory, descending)
// and keeps its k largest; stages repeat on the survivors until one tile is left, whose first k are the answer.
constexpr int kTopkTile = 1024, kTopkThreads = 512, kTopkMaxK = 512;

__global__ void __launch_bounds__(kTopkThreads) mh_topk_stage_kernel(const Costs8 *__restrict__ costs, const unsigned long long *__restrict__ in,
                                                                     int n_in, int k, unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long t[kTopkTile];
    const int base = blockIdx.x * kTopkTile;
    for (int j = threadIdx.x; j < kTopkTile; j += kTopkThreads) {
        const int i = base + j;
        unsigned long long v = 0ull;
        if (i < n_in) v = costs ? rank_key(costs[i].total, i) : in[i];
        t[j] = v;
    }
    __syncthreads();
    for (int size = 2; size <= kTopkTile; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int j = threadIdx.x;                          // compare-exchange number j of kTopkTile / 2
            const int lo = 2 * j - (j & (stride - 1));          // index of the pair's lower element
            const int hi = lo + stride;
            const bool desc = (lo & size) == 0;                 // direction of this bitonic run
            const unsigned long long a = t[lo], b = t[hi];
            if ((a < b) == desc) { t[lo] = b; t[hi] = a; }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < k; j += kTopkThreads)
        out[(size_t)blockIdx.x * k + j] = t[j];
}

// Distinct suggestions (KernelTopKDistinct), round `round`.  The distance of two layouts is the largest
// displacement of any object, max_i max(|dx|, |dy|, rot_weight |drot|) with the rotation difference wrapped into
// [0, PI]; mind[chain] keeps the minimum over the picks so far.  One warp per chain, lanes over objects.  The
// previous round's pick is read from keys[round - 1] on the device, so the host enqueues all rounds at once.
// ref_ext != NULL (a context spread over several devices): the previous pick's layout was handed over by the host,
// it may live on another device.
__global__ void __launch_bounds__(256) mh_distinct_round_kernel(const Costs8 *__restrict__ costs, const PointRec *__restrict__ points, int n,
                                                                int n_chains, int round, float min_dist, float rot_weight, float two_pi,
                                                                float *__restrict__ mind, unsigned long long *keys,
                                                                const PointRec *__restrict__ ref_ext)
{
    const PointRec *ref_layout = ref_ext;
    if (round > 0 && !ref_ext) {
        const unsigned long long pk = keys[round - 1];
        if (pk == 0ull) return;                                 // nothing was left in the previous round (uniform over the grid)
        ref_layout = points + (size_t)(0xFFFFFFFFu - (uint32_t)pk) * n;
    }
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long best = 0ull;
    for (int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chain < n_chains; chain += warps) {
        float m = INFINITY;
        if (round > 0) {
            const PointRec *a = points + (size_t)chain * n, *b = ref_layout;
            float d = 0.f;
            for (int i = lane; i < n; i += 32) {
                const PointRec p = a[i], q = b[i];
                float dr = fabsf(p.rotY - q.rotY);
                dr = fminf(dr, fabsf(two_pi - dr));
                d = fmaxf(d, fmaxf(fmaxf(fabsf(p.x - q.x), fabsf(p.y - q.y)), rot_weight * dr));
            }
            for (int s = 16; s > 0; s >>= 1)
                d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, s));
            m = fminf(mind[chain], d);
        }
        if (lane == 0) {
            mind[chain] = m;
            if (round == 0 || m > min_dist) {
                const unsigned long long k = rank_key(costs[chain].total, chain);
                best = k > best ? k : best;
            }
        }
    }
    block_max_to(best, keys + round);
}

// Packs the context's best chain for a MAX all-reduce: NCCL has no arg-max, so the totalCosts
// (made order-preserving as an unsigned integer) goes in the high word and the complemented
// GLOBAL chain id in the low word (ties go to the lower id); the top bit is flipped so that a
// SIGNED 64-bit MAX orders the keys correctly.
__global__ void mh_bestkey_kernel(const unsigned long long *__restrict__ rank, uint64_t chain_offset, uint64_t chain_stride,
                                  long long *__restrict__ key)
{
    const unsigned long long rk = *rank;
    const uint32_t idx = 0xFFFFFFFFu - (uint32_t)rk;
    const uint64_t g = chain_offset + (uint64_t)idx * chain_stride;
    const uint64_t k = (rk & 0xFFFFFFFF00000000ull) | (uint64_t)(0xFFFFFFFFu - (uint32_t)g);
    *key = (long long)(k ^ 0x8000000000000000ull);
}

template <int G> static int launch_scan_g(const mhLaunch &L)
{
    using WS = WarpState<G, true>;
    const int chains_per_block = WARPS_PER_BLOCK * WS::CPW;
    const int blocks = (L.n_chains + chains_per_block - 1) / chains_per_block;
    const size_t smem = sizeof(float) * ((size_t)L.smem_words + (size_t)WARPS_PER_BLOCK * WS::words(L.n, L.C));
    cudaError_t e = cudaFuncSetAttribute(mh_chain_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    mh_chain_kernel<G><<<blocks, THREADS, smem, static_cast<cudaStream_t>(L.stream)>>>(L);
    return (int)cudaGetLastError();
}

template <int G, int MODE, int WPB> static int launch_delta_w(const mhLaunch &L, int warps)
{
    using WS = WarpState<G>;
    const int chains_per_block = warps * WS::CPW;
    const int blocks = (L.n_chains + chains_per_block - 1) / chains_per_block;
    const size_t smem = sizeof(float) * ((size_t)L.smem_words + (size_t)warps * (WS::words(L.n, 0) + DeltaState<G>::words(L.n, L.C, L.R, MODE)));
    cudaError_t e = cudaFuncSetAttribute(mh_delta_kernel<G, MODE, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    mh_delta_kernel<G, MODE, WPB><<<blocks, warps * 32, smem, static_cast<cudaStream_t>(L.stream)>>>(L);
    return (int)cudaGetLastError();
}

template <int G, int MODE> static int launch_delta_g(const mhLaunch &L)
{
    const int warps = L.warps_per_block >= 1 && L.warps_per_block <= MH_DELTA_THREADS / 32 ? L.warps_per_block : 4;
    return warps <= 4 ? launch_delta_w<G, MODE, 4>(L, warps) : launch_delta_w<G, MODE, 8>(L, warps);
}

template <int G> static int launch_chains_g(const mhLaunch &L)
{
    switch (L.eval_mode) {
    case 1: return launch_delta_g<G, kModeDelta>(L);
    case 2: return launch_delta_g<G, kModeExact>(L);
    default: return launch_scan_g<G>(L);
    }
}

template <int G>
static int launch_score_g(const void *d_problem, int smem_words, int n, int C, int n_layouts, const void *d_points, void *d_costs,
                          void *stream)
{
    using WS = WarpState<G, true>;
    const int per_block = WARPS_PER_BLOCK * WS::CPW;
    const int blocks = (n_layouts + per_block - 1) / per_block;
    const size_t smem = sizeof(float) * ((size_t)smem_words + (size_t)WARPS_PER_BLOCK * WS::words(n, C));
    cudaError_t e = cudaFuncSetAttribute(mh_score_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    mh_score_kernel<G><<<blocks, THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(d_problem), smem_words, n_layouts, static_cast<const PointRec *>(d_points),
        static_cast<Costs8 *>(d_costs));
    return (int)cudaGetLastError();
}

template <int G> static int chain_words(int n, int C, int R, int eval_mode)
{
    if (eval_mode == 1 || eval_mode == 2)                       // the memo kernels keep no clearance-rectangle array
        return WarpState<G>::words(n, 0) + DeltaState<G>::words(n, C, R, eval_mode == 1 ? kModeDelta : kModeExact);
    return WarpState<G, true>::words(n, C);
}

} // namespace mh

extern "C" {

int mhdev_chain_smem_bytes(int smem_words, int n, int C, int R, int lanes, int eval_mode, int warps)
{
    int w = 0;
    const bool memo = eval_mode == 1 || eval_mode == 2;
    switch (lanes) {
    case 1: w = mh::chain_words<1>(n, C, R, eval_mode); break;
    case 2: w = mh::chain_words<2>(n, C, R, eval_mode); break;
    case 4: w = mh::chain_words<4>(n, C, R, eval_mode); break;
    case 8: w = mh::chain_words<8>(n, C, R, eval_mode); break;
    case 16: w = mh::chain_words<16>(n, C, R, eval_mode); break;
    case 32: w = mh::chain_words<32>(n, C, R, eval_mode); break;
    default: return -1;
    }
    if (memo && (n + lanes - 1) / lanes > 32) return -1; /* the per-lane row flags are one 32-bit word */
    if (!memo || warps < 1 || warps > MH_DELTA_THREADS / 32) warps = mh::WARPS_PER_BLOCK;
    return 4 * (smem_words + warps * w);
}

int mhdev_launch_chains(const mhLaunch *l)
{
    if (l->n_chains <= 0) return 0;
    switch (l->lanes) {
    case 1: return mh::launch_chains_g<1>(*l);
    case 2: return mh::launch_chains_g<2>(*l);
    case 4: return mh::launch_chains_g<4>(*l);
    case 8: return mh::launch_chains_g<8>(*l);
    case 16: return mh::launch_chains_g<16>(*l);
    case 32: return mh::launch_chains_g<32>(*l);
    }
    return (int)cudaErrorInvalidValue;
}

int mhdev_launch_score(const void *d_problem, int smem_words, int n, int C, int R, int n_layouts, int lanes, const void *d_points,
                       void *d_costs, void *stream)
{
    (void)R;
    if (n_layouts <= 0) return 0;
    switch (lanes) {
    case 1: return mh::launch_score_g<1>(d_problem, smem_words, n, C, n_layouts, d_points, d_costs, stream);
    case 2: return mh::launch_score_g<2>(d_problem, smem_words, n, C, n_layouts, d_points, d_costs, stream);
    case 4: return mh::launch_score_g<4>(d_problem, smem_words, n, C, n_layouts, d_points, d_costs, stream);
    case 8: return mh::launch_score_g<8>(d_problem, smem_words, n, C, n_layouts, d_points, d_costs, stream);
    case 16: return mh::launch_score_g<16>(d_problem, smem_words, n, C, n_layouts, d_points, d_costs, stream);
    case 32: return mh::launch_score_g<32>(d_problem, smem_words, n, C, n_layouts, d_points, d_costs, stream);
    }
    return (int)cudaErrorInvalidValue;
}

int mhdev_launch_exchange(int n_chains, uint64_t chain_offset, uint64_t chain_stride, int rungs, uint64_t epoch, uint64_t it_last,
                          uint64_t seed, const float *d_all_total, const float *d_all_beta, uint64_t gather_base,
                          uint64_t gather_stride, uint64_t gather_local, float *d_beta, void *d_stats, void *stream)
{
    if (n_chains <= 0) return 0;
    mh::mh_exchange_kernel<<<(n_chains + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        n_chains, chain_offset, chain_stride, rungs, epoch, it_last, seed, d_all_total, d_all_beta, gather_base, gather_stride,
        gather_local, d_beta, static_cast<unsigned long long *>(d_stats));
    return (int)cudaGetLastError();
}

int mhdev_launch_retarget(int n_chains, int rungs, const float *d_old_ladder, const float *d_new_ladder, float *d_beta, void *stream)
{
    if (n_chains <= 0) return 0;
    mh::mh_retarget_kernel<<<(n_chains + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(n_chains, rungs, d_old_ladder, d_new_ladder, d_beta);
    return (int)cudaGetLastError();
}

static int rank_grid(long long threads_wanted)
{
    long long blocks = (threads_wanted + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int mhdev_launch_argmax(const void *d_costs, int n_chains, void *d_key, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(d_key, 0, 8, st);
    if (e != cudaSuccess) return (int)e;
    mh::mh_argmax_kernel<<<rank_grid(n_chains), 256, 0, st>>>(static_cast<const mh::Costs8 *>(d_costs), n_chains,
                                                              static_cast<unsigned long long *>(d_key));
    return (int)cudaGetLastError();
}

int mhdev_topk_work_items(int n_chains, int k)
{
    const int tiles = (n_chains + mh::kTopkTile - 1) / mh::kTopkTile;
    return tiles * (k < 1 ? 1 : k);
}

int mhdev_launch_topk(const void *d_costs, int n_chains, int k, void *d_work, void *d_out, void *stream)
{
    if (k < 1 || k > mh::kTopkMaxK) return (int)cudaErrorInvalidValue;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int items = mhdev_topk_work_items(n_chains, k);
    unsigned long long *w[2] = { static_cast<unsigned long long *>(d_work), static_cast<unsigned long long *>(d_work) + items };
    const mh::Costs8 *costs = static_cast<const mh::Costs8 *>(d_costs);
    const unsigned long long *in = nullptr;
    int n_in = n_chains, side = 0;
    for (;;) {
        const int tiles = (n_in + mh::kTopkTile - 1) / mh::kTopkTile;
        unsigned long long *out = tiles == 1 ? static_cast<unsigned long long *>(d_out) : w[side];
        mh::mh_topk_stage_kernel<<<tiles, mh::kTopkThreads, 0, st>>>(costs, in, n_in, k, out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        if (tiles == 1) return 0;
        costs = nullptr;
        in = out;
        n_in = tiles * k;
        side ^= 1;
    }
}

int mhdev_launch_distinct_round(const void *d_costs, const void *d_points, int n, int n_chains, int round, float min_dist,
                                float rot_weight, float two_pi, float *d_mind, void *d_keys, const void *d_ref_layout, void *stream)
{
    mh::mh_distinct_round_kernel<<<rank_grid((long long)n_chains * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const mh::Costs8 *>(d_costs), static_cast<const mh::PointRec *>(d_points), n, n_chains, round, min_dist, rot_weight,
        two_pi, d_mind, static_cast<unsigned long long *>(d_keys), static_cast<const mh::PointRec *>(d_ref_layout));
    return (int)cudaGetLastError();
}

int mhdev_launch_bestkey(const void *d_rank_key, uint64_t chain_offset, uint64_t chain_stride, void *d_key, void *stream)
{
    mh::mh_bestkey_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned long long *>(d_rank_key), chain_offset,
                                                                        chain_stride, static_cast<long long *>(d_key));
    return (int)cudaGetLastError();
}

int mhdev_device_limits(int *max_smem_per_block, int *max_smem_per_sm, int *sm_count, int *clock_khz, int *cc_major, int *cc_minor,
                        char *name, int name_len)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int v = 0;
    if (max_smem_per_block) { e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); if (e) return (int)e; *max_smem_per_block = v; }
    if (max_smem_per_sm) { e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev); if (e) return (int)e; *max_smem_per_sm = v; }
    if (sm_count) { e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); if (e) return (int)e; *sm_count = v; }
    if (clock_khz) { e = cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev); if (e) return (int)e; *clock_khz = v; }
    if (cc_major) { e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); if (e) return (int)e; *cc_major = v; }
    if (cc_minor) { e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); if (e) return (int)e; *cc_minor = v; }
    if (name && name_len > 0) {
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e) return (int)e;
        snprintf(name, (size_t)name_len, "%s", prop.name);
    }
    return 0;
}

int mhdev_device_count(int *count) { return (int)cudaGetDeviceCount(count); }
int mhdev_get_device(int *dev) { return (int)cudaGetDevice(dev); }
int mhdev_set_device(int dev) { return (int)cudaSetDevice(dev); }

// Device memory comes from a stream-ordered pool owned by the library, one per device, that
// keeps freed blocks for the next call (release threshold = unlimited): the reference pays 12
// cudaMalloc + cudaFree per call (Kernel.cu:879-967), and on this driver a cudaFree of the
// ~130 MB a 65536-chain call needs was measured at up to 600 ms.  KernelTrim() gives the cache back.
static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64];
static bool g_pool_ready[64];

static cudaError_t pool_for_current_device(cudaMemPool_t *pool)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pool_ready[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        e = cudaMemPoolCreate(&g_pools[dev], &props);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        e = cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        g_pool_ready[dev] = true;
    }
    *pool = g_pools[dev];
    return cudaSuccess;
}

int mhdev_malloc(void **p, size_t bytes, void *stream)
{
    cudaMemPool_t pool;
    cudaError_t e = pool_for_current_device(&pool);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMallocFromPoolAsync(p, bytes ? bytes : 16, pool, static_cast<cudaStream_t>(stream));
}
void mhdev_free(void *p, void *stream) { if (p) cudaFreeAsync(p, static_cast<cudaStream_t>(stream)); }
int mhdev_trim(void)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (dev >= 0 && dev < 64 && g_pool_ready[dev]) {
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return (int)e;
        return (int)cudaMemPoolTrimTo(g_pools[dev], 0);
    }
    return 0;
}
int mhdev_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
}
int mhdev_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream));
}
int mhdev_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
}
int mhdev_memset(void *dst, int value, size_t bytes, void *stream)
{
    return (int)cudaMemsetAsync(dst, value, bytes, static_cast<cudaStream_t>(stream));
}
int mhdev_stream_create(void **stream)
{
    cudaStream_t s;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    *stream = e == cudaSuccess ? static_cast<void *>(s) : nullptr;
    return (int)e;
}
void mhdev_stream_destroy(void *stream) { if (stream) cudaStreamDestroy(static_cast<cudaStream_t>(stream)); }
int mhdev_stream_sync(void *stream) { return (int)cudaStreamSynchronize(static_cast<cudaStream_t>(stream)); }
int mhdev_event_create(void **ev)
{
    cudaEvent_t e;
    cudaError_t r = cudaEventCreate(&e);
    *ev = r == cudaSuccess ? static_cast<void *>(e) : nullptr;
    return (int)r;
}
void mhdev_event_destroy(void *ev) { if (ev) cudaEventDestroy(static_cast<cudaEvent_t>(ev)); }
int mhdev_stream_wait_event(void *stream, void *ev) { return (int)cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(ev), 0); }
int mhdev_event_record(void *ev, void *stream) { return (int)cudaEventRecord(static_cast<cudaEvent_t>(ev), static_cast<cudaStream_t>(stream)); }
int mhdev_event_elapsed_ms(void *e0, void *e1, float *ms)
{
    cudaError_t r = cudaEventSynchronize(static_cast<cudaEvent_t>(e1));
    if (r != cudaSuccess) return (int)r;
    return (int)cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(e0), static_cast<cudaEvent_t>(e1));
}
// Small pinned read-back slots (64 bytes each) from ONE page-locked arena the library allocates once per process:
// cudaMallocHost / cudaFreeHost per context were measured at 20-40 ms and up to 0.8 s on this driver (they
// synchronise and re-map), which a one-shot call cannot afford.  A context takes a slot when it is created and gives
// it back when it is destroyed; if the arena is exhausted the slot is plain (pageable) memory -- still correct, the
// read-backs are followed by a stream synchronisation.
static std::mutex g_scratch_mutex;
static unsigned char *g_scratch_arena;
static constexpr int kScratchSlots = 1024, kScratchBytes = 64;
static unsigned char g_scratch_used[kScratchSlots];

int mhdev_scratch_acquire(void **p)
{
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    if (!g_scratch_arena) {
        void *a = nullptr;
        if (cudaMallocHost(&a, (size_t)kScratchSlots * kScratchBytes) == cudaSuccess) g_scratch_arena = static_cast<unsigned char *>(a);
        else (void)cudaGetLastError();
    }
    if (g_scratch_arena)
        for (int i = 0; i < kScratchSlots; i++)
            if (!g_scratch_used[i]) {
                g_scratch_used[i] = 1;
                *p = g_scratch_arena + (size_t)i * kScratchBytes;
                memset(*p, 0, kScratchBytes);
                return 0;
            }
    *p = calloc(1, kScratchBytes);
    return *p ? 0 : (int)cudaErrorMemoryAllocation;
}

void mhdev_scratch_release(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    unsigned char *q = static_cast<unsigned char *>(p);
    if (g_scratch_arena && q >= g_scratch_arena && q < g_scratch_arena + (size_t)kScratchSlots * kScratchBytes)
        g_scratch_used[(q - g_scratch_arena) / kScratchBytes] = 0;
    else
        free(p);
}

int mhdev_host_alloc(void **p, size_t bytes) { return (int)cudaMallocHost(p, bytes ? bytes : 16); }
void mhdev_host_free(void *p) { if (p) cudaFreeHost(p); }
int mhdev_host_register(void *p, size_t bytes) { return (int)cudaHostRegister(p, bytes, cudaHostRegisterPortable); }
void mhdev_host_unregister(void *p) { if (p) cudaHostUnregister(p); }
const char *mhdev_error_string(int code) { return cudaGetErrorString(static_cast<cudaError_t>(code)); }

} // extern "C"
