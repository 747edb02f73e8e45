// mh_delta.cuh -- evaluation of a proposal through memos: MH_EVAL_MEMO (exact) and MH_EVAL_DELTA.
//
// A proposal moves one object (translate, rotate) or two (swap).  Full evaluation redoes
// n^2 + C n pair terms for it; here only what the moved objects touch is recomputed:
//
//   symmetry   sum_i (5 - min_j key(i,j)) -- every row keeps its minimum and ONE column that attains
//              it (KM memo).  A move of object m changes row m (rescanned by the whole group) and
//              column m of every other row: the row minimum becomes min(old, key(i,m)) unless the
//              remembered column was m itself, in which case that row is rescanned too.  min is exact,
//              so the memo always equals what a full evaluation computes -- bit for bit.
//   clearance  an INTEGER sum (mh_costs.cuh), updated by the pairs the move touches: (k, m) for every
//              clearance k, and (k', i) for every clearance k' whose source object is m.  Integer addition is
//              associative, so the updated sum IS the from-scratch sum of the new layout.
//   surface    the moved objects' own rectangles and the clearances with the same INDEX (quirk Q7).
//   focal      the memoised cosines;  visual balance: area * position.
//   pair-wise  relationships that name a moved object; the others come from the PR memo.
//
// MH_EVAL_MEMO re-adds the cheap O(n) float sums from their memos in the full evaluation's order, so every
// total is bit-identical to the plain scan's.  MH_EVAL_DELTA keeps those float sums as running sums instead
// (new - old), which drift in the last bits and are rebuilt every kRefresh iterations: statistically, not
// bitwise, equivalent (tests: KS against the oracle, running total against a fresh evaluation).  The costs a
// caller receives always come from mh_score_kernel, i.e. from a full evaluation of the emitted layout.
#pragma once
#include "mh_costs.cuh"

namespace mh {

// The loops of delta_eval are deliberately NOT unrolled: the kernel's per-iteration code path must stay
// inside the 32 KB instruction cache (measured at n = 50: 8.4e8 proposals/s rolled against 7.2e8 with
// the compiler's 4x unrolling -- with the unrolled code the warps of an SM evict each other's lines).
#ifndef MH_DU_CLR
#define MH_DU_CLR 1
#endif
#ifndef MH_DU_ROW
#define MH_DU_ROW 1
#endif
#ifndef MH_DU_SCAN
#define MH_DU_SCAN 1
#endif
#ifndef MH_DU_SUM
#define MH_DU_SUM 1
#endif
constexpr int kDuClr = MH_DU_CLR, kDuRow = MH_DU_ROW, kDuScan = MH_DU_SCAN, kDuSum = MH_DU_SUM;
constexpr int kRefresh = 128; // iterations between full rebuilds of the memo and the running sums

// Which memos a chain keeps (template parameter MODE of mh_delta_kernel):
//   kModeDelta   MH_EVAL_DELTA: KM + PR, running sums of the additive float terms, integer clearance sum;
//   kModeExact   MH_EVAL_MEMO : KM + PR + SV, float sums re-added from the memos, integer clearance sum.
// (Round 1 had a third form with a memo of clearance ROW sums for 16- and 32-lane groups -- float sums cannot be
// updated term by term, so whole rows were re-added whenever one of their overlaps could have changed, and on
// the piled-up rooms the sampler produces (it MAXIMISES totalCosts, quirk Q10, which rewards overlap) a third of
// the rows were flagged per move.  The integer sum makes the term itself updatable and the row memo is gone.)
constexpr int kModeDelta = 0, kModeExact = 1;

template <int G> struct DeltaState {
    static constexpr int CPW = 32 / G;
    float2 *KM; // [2][n][CPW] {row minimum of key, a column attaining it (int bits; -1 = none below 5)}
    float2 *PR; // [R][CPW]    {distance penalty, angle penalty} of every relationship
    float *SV;  // [n + C][CPW] exact mode: area outside the room of object i's rectangle, then of clearance k's
    int n, nC, nR;
    __host__ __device__ static int words(int n, int C, int R, int mode)
    {
        const int w = CPW * (4 * n + 2 * R + (mode != kModeDelta ? n + C : 0));
        return (w + 3) & ~3;                                    // the next warp's float4 state must stay 16-byte aligned
    }
    __device__ __forceinline__ void bind(float *base, int n_, int C, int R)
    {
        n = n_;
        nC = C;
        nR = R;
        KM = reinterpret_cast<float2 *>(base);
        PR = KM + 2 * n * CPW;
        SV = reinterpret_cast<float *>(PR + R * CPW);
    }
    __device__ __forceinline__ float2 &km(int sel, int i, int c) const
    {
        MH_CHECK((sel == 0 || sel == 1) && i >= 0 && i < n && c >= 0 && c < CPW);
        return KM[(sel * n + i) * CPW + c];
    }
    __device__ __forceinline__ float2 &pr(int r, int c) const
    {
        MH_CHECK(r >= 0 && r < nR && c >= 0 && c < CPW);
        return PR[r * CPW + c];
    }
    __device__ __forceinline__ float &sv(int i, int c) const
    {
        MH_CHECK(i >= 0 && i < n + nC && c >= 0 && c < CPW);
        return SV[i * CPW + c];
    }
};

// Committed running sums of the additive terms (positive magnitudes, as in RawTerms).
struct RunSums {
    float pw, pa, vbx, vby, focal, surf;
    long long clr_q;   // the clearance term: an exact integer sum in both modes
};

// Lexicographic (key, column) minimum over the G lanes of a group.
constexpr bool kDeltaStr = true; // the delta kernel uses the interleaved lane mapping (LaneMap<G, true>)

template <int G, bool STR> __device__ __forceinline__ void group_argmin(float &k, int &arg)
{
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, k, m * LaneMap<G, STR>::xor_step);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, m * LaneMap<G, STR>::xor_step);
        if (ok < k || (ok == k && (unsigned)oa < (unsigned)arg)) {
            k = ok;
            arg = oa;
        }
    }
}

template <int G, bool STR> __device__ __forceinline__ int group_min_int(int v)
{
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1)
        v = min(v, __shfl_xor_sync(0xffffffffu, v, m * LaneMap<G, STR>::xor_step));
    return v;
}

// Two lexicographic (key, column) minima at once: after the first exchange the lanes with even g carry
// pair 0 and the lanes with odd g pair 1, so both reductions share one butterfly.  On return lanes with
// even g hold the group's (k0, a0) and lanes with odd g its (k1, a1); for G = 1 nothing moves.
template <int G, bool STR> __device__ __forceinline__ void group_argmin2(const int g, float &k0, int &a0, float &k1, int &a1)
{
    if (G == 1) return;
    constexpr int step = LaneMap<G, STR>::xor_step;
    const bool odd = g & 1;
    float k = odd ? k1 : k0, sk = odd ? k0 : k1;
    int arg = odd ? a1 : a0, sa = odd ? a0 : a1;
    {
        const float ok = __shfl_xor_sync(0xffffffffu, sk, step);
        const int oa = __shfl_xor_sync(0xffffffffu, sa, step);
        if (ok < k || (ok == k && (unsigned)oa < (unsigned)arg)) { k = ok; arg = oa; }
    }
#pragma unroll
    for (int m = 2; m < G; m <<= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, k, m * step);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, m * step);
        if (ok < k || (ok == k && (unsigned)oa < (unsigned)arg)) { k = ok; arg = oa; }
    }
    k0 = k1 = k;
    a0 = a1 = arg;
}

// Relationship penalties a proposal recomputed, kept in registers until the accept decision.  The
// relationships that name a moved object are dealt round-robin to the lanes of the group, so a lane
// rarely holds more than one; beyond two the commit recomputes.
struct RelStash {
    int r0, r1;
    float2 v0, v1;
    int overflow; // some lane of the group held more than two: the commit recomputes the touched ones
};

// Minimum of key(row, j) over j = j0, j0+step, ... and a column attaining it.
template <int CPW>
__device__ __forceinline__ void sym_scan(const RowRef &rr, const float4 *Pc, int n, int j0, int step, float pi_f, float &k, int &arg)
{
    k = 5.0f;
    arg = -1;
    for (int j = j0; j < n; j += step) {
        const float kk = sym_key(rr, Pc[j * CPW], pi_f);
        if (kk < k) {
            k = kk;
            arg = j;
        }
    }
}

// Build the symmetry memo of the CURRENT layout: every row scans all columns.
template <int G>
__device__ __forceinline__ void sym_memo_build(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, int c, int g, int sel)
{
    constexpr int CPW = WarpState<G>::CPW;
    const mhProblemHeader *h = P.h;
    const int n = h->n;
    const float pi_f = 0.5f * h->two_pi;
    const float4 *Pc = S.P4 + c;
    for (int i = g; i < n; i += G) {
        float k;
        int arg;
        sym_scan<CPW>(sym_row(h, Pc[i * CPW]), Pc, n, 0, 1, pi_f, k, arg);
        D.km(sel, i, c) = make_float2(k, __int_as_float(arg));
    }
    __syncwarp();
}

// Column update of row i (an object that did not move) for the proposal that moved columns a and b
// (b = a's duplicate when only one moved): the row minimum becomes min(old, key(i, a), key(i, b)) -- exact,
// since every other column is unchanged -- unless the remembered column itself moved and neither moved
// column does at least as well as the old minimum: then the row is flagged (bit p) for a rescan.
template <int G>
__device__ __forceinline__ void sym_col_update(const mhProblemHeader *h, const DeltaState<G> &D, const int c, const int sel, const int i,
                                               const int p, const float4 pi, const int a, const int b, const bool mvb, const bool any_b,
                                               const float4 na, const float4 nbx, const float pi_f, unsigned &flags)
{
    const float2 km = D.km(sel, i, c);
    float k = km.x;
    int arg = __float_as_int(km.y);
    const bool hit = arg == a || (mvb && arg == b);
    const RowRef rr = sym_row(h, pi);
    const float k1 = sym_key(rr, na, pi_f);
    if (k1 < k || (k1 == k && hit)) { k = k1; arg = a; }
    if (any_b) {
        const float k2 = sym_key(rr, nbx, pi_f);
        if (mvb && k2 < k) { k = k2; arg = b; }
    }
    if (hit && k == km.x && arg == __float_as_int(km.y)) flags |= 1u << p;
    else D.km(1 - sel, i, c) = make_float2(k, __int_as_float(arg));
}

// Rows rescanned by the whole group, two per pass over the columns: the moved rows first, then the
// flagged ones; each group works through its own queue, the warp pays for the longest.  Then the
// symmetry sum of the proposal's memo (buffer 1-sel), each lane adding its rows in the order of the full
// scan, so that the value is bit-identical to eval_terms' (min is exact).  Every lane of the warp must
// call this; S.P4 holds the proposal.
template <int G>
__device__ __forceinline__ float sym_rescan_sum(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                                const int sel, const int a, const int b, unsigned flags)
{
    constexpr int CPW = WarpState<G>::CPW;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NONE = 0x7fffffff;
    const mhProblemHeader *h = P.h;
    const int n = h->n;
    const float pi_f = 0.5f * h->two_pi;
    const float4 *Pc = S.P4 + c;
    const bool mvb = b >= 0;
    auto take = [&](bool consume) {                              // next flagged row of this group (NONE: queue empty)
        const int mine = flags ? g + (__ffs(flags) - 1) * G : NONE;
        const int row = group_min_int<G, kDeltaStr>(mine);
        if (consume && mine == row && mine != NONE) flags &= flags - 1;
        return row;
    };
    int row0 = a, row1 = take(!mvb);
    if (mvb) row1 = b;
    for (;;) {
        const bool v0 = row0 != NONE, v1 = row1 != NONE;
        MH_CHECK((!v0 || (row0 >= 0 && row0 < n)) && (!v1 || (row1 >= 0 && row1 < n)));
        const RowRef r0 = sym_row(h, Pc[(v0 ? row0 : a) * CPW]);
        const RowRef r1 = sym_row(h, Pc[(v1 ? row1 : a) * CPW]);
        float k0 = 5.0f, k1 = 5.0f;
        int a0 = -1, a1 = -1;
#pragma unroll kDuScan
        for (int j = g; j < n; j += G) {
            const float4 q = Pc[j * CPW];
            const float x0 = sym_key(r0, q, pi_f), x1 = sym_key(r1, q, pi_f);
            if (x0 < k0) { k0 = x0; a0 = j; }
            if (x1 < k1) { k1 = x1; a1 = j; }
        }
        group_argmin2<G, kDeltaStr>(g, k0, a0, k1, a1);
        if (G == 1) {
            if (v0) D.km(1 - sel, row0, c) = make_float2(k0, __int_as_float(a0));
            if (v1) D.km(1 - sel, row1, c) = make_float2(k1, __int_as_float(a1));
        } else if (g < 2) {                                      // g = 0 holds row0's minimum, g = 1 row1's
            const int row = g ? row1 : row0;
            if (row != NONE) D.km(1 - sel, row, c) = make_float2(k0, __int_as_float(a0));
        }
        row0 = take(true);
        row1 = take(true);
        if (!__any_sync(FULL, row0 != NONE)) break;
    }
    __syncwarp();
    float s = 0.f;
#pragma unroll kDuSum
    for (int i = g; i < n; i += G)
        s += 5.0f - D.km(1 - sel, i, c).x;
    return group_sum<G, kDeltaStr>(s);
}

// The clearance sum of the current layout from scratch, WITHOUT the per-warp array of clearance rectangles the
// plain scan keeps (the memo kernels have no use for it between rebuilds, and its shared memory buys resident
// warps): every lane walks its rows and rebuilds each clearance rectangle on the fly.  Integer sum: the value is
// the plain scan's whatever the order.  Runs once per launch (and per refresh in delta mode).
template <int G>
__device__ __forceinline__ long long clearance_sum_q(const SmemProblem &P, const float4 *Pc, const int g)
{
    constexpr int CPW = 32 / G;
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C;
    long long acc = 0;
#pragma unroll 1
    for (int i = g; i < n; i += G) {
        const float4 pi = Pc[i * CPW];
        const int4 bi = box_at_q(P.obj_boxq[i], P.obj_v0xq[i], to_fixed(h, pi.x), to_fixed(h, pi.y));
#pragma unroll 1
        for (int k = 0; k < C; k++) {
            const float2 ps = *reinterpret_cast<const float2 *>(&Pc[P.clr_src[k] * CPW]);
            acc = overlap_add_q(bi, box_at_q(P.clr_boxq[k], P.clr_v0xq[k], to_fixed(h, ps.x), to_fixed(h, ps.y)), acc);
        }
    }
    return group_sum_q<G, kDeltaStr>(acc);
}

// Rebuild the memo and the running sums of the CURRENT layout from scratch; returns its total.
template <int G, int MODE>
__device__ __forceinline__ float delta_rebuild(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, int c, int g, int sel,
                                               RunSums &cur)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    const mhProblemHeader *h = P.h;
    const int R = h->R;
    const float4 *Pc = S.P4 + c;
    RawTerms t;
    eval_terms<G, false, kDeltaStr, false, false, true>(P, S, c, g, t);   // every term but the clearance (no S.CB in this kernel)
    t.clr_q = clearance_sum_q<G>(P, Pc, g);
    t.clr = clearance_value(h, t.clr_q);
    cur.pw = t.pw; cur.pa = t.pa; cur.vbx = t.vbx; cur.vby = t.vby; cur.focal = t.focal; cur.clr_q = t.clr_q; cur.surf = t.surf;
    sym_memo_build<G>(P, S, D, c, g, sel);
    for (int r = g; r < R; r += G) {
        float pd, pa;
        rel_pen<CPW>(P, Pc, r, pd, pa);
        D.pr(r, c) = make_float2(pd, pa);
    }
    if (MODE != kModeDelta) {                                   // surface values
        for (int i = g; i < h->n; i += G) {
            const float4 pi = Pc[i * CPW];
            D.sv(i, c) = outside_room(box_at(P.obj_box[i], P.obj_v0x[i], pi.x, pi.y), h);
            if (i < h->C) D.sv(h->n + i, c) = outside_room(box_at(P.clr_box[i], P.clr_v0x[i], pi.x, pi.y), h);   // Q7
        }
    }
    __syncwarp();
    return combine(h, t).total;
}

// ---------------------------------------------------------------------------------------------------
// Change of the integer clearance sum when object a (and b; -1 = none) moves from oa / ob to na / nb.
//   part 1 (clearance_delta_moved_rows): the moved objects' own rectangles against EVERY clearance -- a
//           clearance may itself have moved with its source;
//   part 2 (inside the callers' pass over the rows): the clearances SOURCED at a moved object against the
//           rectangles of the objects that did not move.
// Pc already holds the proposal; the old rectangles are rebuilt from oa / ob.  Returns this lane's share.
template <int G>
__device__ __forceinline__ long long clearance_delta_moved_rows(const SmemProblem &P, const float4 *Pc, const int g, const int a, const int b,
                                                                const bool any_b, const float4 oa, const float4 ob, const float4 na,
                                                                const float4 nb)
{
    constexpr int CPW = 32 / G;
    const mhProblemHeader *h = P.h;
    const int C = h->C;
    const bool mvb = b >= 0;
    const int4 zero4 = make_int4(0, 0, 0, 0);
    const int4 ka = P.obj_boxq[a];
    const int va = P.obj_v0xq[a];
    const int4 box_oa = box_at_q(ka, va, to_fixed(h, oa.x), to_fixed(h, oa.y));
    const int4 box_na = box_at_q(ka, va, to_fixed(h, na.x), to_fixed(h, na.y));
    int4 box_ob = zero4, box_nb = zero4;                        // an empty rectangle: its overlaps are exactly zero
    if (mvb) {
        const int4 kb = P.obj_boxq[b];
        const int vb = P.obj_v0xq[b];
        box_ob = box_at_q(kb, vb, to_fixed(h, ob.x), to_fixed(h, ob.y));
        box_nb = box_at_q(kb, vb, to_fixed(h, nb.x), to_fixed(h, nb.y));
    }
    long long d = 0;
#pragma unroll kDuClr
    for (int k = g; k < C; k += G) {
        const int src = P.clr_src[k];
        const float2 pn = *reinterpret_cast<const float2 *>(&Pc[src * CPW]);
        float2 po = pn;                                         // where the clearance's source was before the move
        if (src == a) po = make_float2(oa.x, oa.y);
        if (src == b) po = make_float2(ob.x, ob.y);
        const int4 kq = P.clr_boxq[k];
        const int vq = P.clr_v0xq[k];
        const int4 cb_new = box_at_q(kq, vq, to_fixed(h, pn.x), to_fixed(h, pn.y));
        const int4 cb_old = box_at_q(kq, vq, to_fixed(h, po.x), to_fixed(h, po.y));
        d += overlap_q(box_na, cb_new) - overlap_q(box_oa, cb_old);
        if (any_b) d += overlap_q(box_nb, cb_new) - overlap_q(box_ob, cb_old);
    }
    return d;
}

// The clearances sourced at the moved objects, two per pass (t0, t0 + 1 of the concatenated adjacency lists of a and
// b): old and new rectangle of each; an absent one is an empty rectangle.
struct MovedClearances {
    int4 mo0, mn0, mo1, mn1;
};

__device__ __forceinline__ MovedClearances moved_clearances(const SmemProblem &P, const int a, const int b, const float4 oa, const float4 ob,
                                                            const float4 na, const float4 nb, const int t0)
{
    const mhProblemHeader *h = P.h;
    const bool mvb = b >= 0;
    const int ca0 = P.clr_adj_off[a], na_c = P.clr_adj_off[a + 1] - ca0;
    const int cb0 = mvb ? P.clr_adj_off[b] : 0, nb_c = mvb ? P.clr_adj_off[b + 1] - cb0 : 0;
    const int tot = na_c + nb_c;
    MovedClearances m;
    m.mo0 = m.mn0 = m.mo1 = m.mn1 = make_int4(0, 0, 0, 0);
    if (t0 < tot) {
        const bool fa = t0 < na_c;
        const int k = P.clr_adj[fa ? ca0 + t0 : cb0 + t0 - na_c];
        const float4 po = fa ? oa : ob, pn = fa ? na : nb;
        m.mo0 = box_at_q(P.clr_boxq[k], P.clr_v0xq[k], to_fixed(h, po.x), to_fixed(h, po.y));
        m.mn0 = box_at_q(P.clr_boxq[k], P.clr_v0xq[k], to_fixed(h, pn.x), to_fixed(h, pn.y));
    }
    if (t0 + 1 < tot) {
        const bool fa = t0 + 1 < na_c;
        const int k = P.clr_adj[fa ? ca0 + t0 + 1 : cb0 + t0 + 1 - na_c];
        const float4 po = fa ? oa : ob, pn = fa ? na : nb;
        m.mo1 = box_at_q(P.clr_boxq[k], P.clr_v0xq[k], to_fixed(h, po.x), to_fixed(h, po.y));
        m.mn1 = box_at_q(P.clr_boxq[k], P.clr_v0xq[k], to_fixed(h, pn.x), to_fixed(h, pn.y));
    }
    return m;
}

// how many clearances the moved objects carry: this chain's count
__device__ __forceinline__ int moved_clearance_count(const SmemProblem &P, const int a, const int b)
{
    int tot = P.clr_adj_off[a + 1] - P.clr_adj_off[a];
    if (b >= 0) tot += P.clr_adj_off[b + 1] - P.clr_adj_off[b];
    return tot;
}

// Evaluate the proposal that moved object a (and b; -1 = none) from oa/ob to na/nb; S.P4 already
// holds the new state.  Writes the proposal's KM memo into buffer 1-sel, returns its total and its
// running sums.  Every lane of the warp must call this.
//
// The code is written for a warp that holds several chains whose moves differ: there is no branch
// on the move type.  A one-object move is evaluated as a two-object move whose second object has an
// empty rectangle (its overlaps are exactly zero) and repeats the first object's symmetry column;
// loop trip counts that depend on the move (clearances sourced at a moved object, relationships that
// name one, rows to rescan) are the maximum over the warp, lanes without work evaluate a dummy and
// drop the result.
template <int G>
__device__ __forceinline__ float delta_eval(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                            const int sel, const int a, const int b, const float4 oa, const float4 ob, const float4 na,
                                            const float4 nb, const RunSums &cur, RunSums &star, RelStash &stash)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    constexpr unsigned FULL = 0xffffffffu;
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C;
    const float4 *Pc = S.P4 + c;
    const bool mvb = b >= 0;
    const bool any_b = __any_sync(FULL, mvb);
    const float pi_f = 0.5f * h->two_pi;
    MH_CHECK(a >= 0 && a < n && b >= -1 && b < n && b != a && (sel == 0 || sel == 1));

    float d_pw = 0.f, d_pa = 0.f, d_vbx = 0.f, d_vby = 0.f, d_focal = 0.f, d_surf = 0.f;

    // ---- the moved objects' own rectangles and the clearances with the same INDEX (quirk Q7), one job per
    //      lane: 0 = object a, 1 = object b, 2 = clearance a, 3 = clearance b --------------------------------
    for (int job = g; job < 4; job += G) {
        const bool second = job & 1, isclr = job >= 2;
        const int m = second ? b : a;
        const bool valid = m >= 0 && (!isclr || m < C);
        const int mm = valid ? m : 0;
        const float4 kb = isclr ? P.clr_box[mm] : P.obj_box[mm];
        const float v0 = isclr ? P.clr_v0x[mm] : P.obj_v0x[mm];
        const float4 o = second ? ob : oa, w = second ? nb : na;
        const float ds = outside_room(box_at(kb, v0, w.x, w.y), h) - outside_room(box_at(kb, v0, o.x, o.y), h);
        if (valid) {
            d_surf += ds;
            if (!isclr) {
                const float area = P.obj_area[mm];
                d_focal += w.w - o.w;
                d_vbx += area * (w.x - o.x);
                d_vby += area * (w.y - o.y);
            }
        }
    }

    // ---- clearance, part 1: every clearance against the moved objects --------------------------------------------
    long long d_clr = clearance_delta_moved_rows<G>(P, Pc, g, a, b, any_b, oa, ob, na, nb);

    // ---- relationships that name a moved object (CSR by object; one that names both is taken from a's
    //      list only), dealt round-robin to the lanes of the group -------------------------------------------
    stash.r0 = stash.r1 = -1;
    stash.overflow = 0;
    {
        const int ra0 = P.rel_adj_off[a], na_r = P.rel_adj_off[a + 1] - ra0;
        const int rb0 = mvb ? P.rel_adj_off[b] : 0, nb_r = mvb ? P.rel_adj_off[b + 1] - rb0 : 0;
        const int tot = na_r + nb_r;
        const int tmax = __reduce_max_sync(FULL, tot);
        int slot = 0;
        for (int t = g; t < tmax; t += G, slot++) {
            int r = -1;
            if (t < na_r) {
                r = P.rel_adj[ra0 + t];
            } else if (t < tot) {
                r = P.rel_adj[rb0 + t - na_r];
                const int4 id = P.rel_idx[r];
                if (id.x == a || id.y == a || id.z == a || id.w == a) r = -1;
            }
            float pd, pe;
            rel_pen<CPW>(P, Pc, r >= 0 ? r : 0, pd, pe);
            if (r >= 0) {
                const float2 old = D.pr(r, c);
                d_pw += pd - old.x;
                d_pa += pe - old.y;
                if (slot == 0) { stash.r0 = r; stash.v0 = make_float2(pd, pe); }
                else if (slot == 1) { stash.r1 = r; stash.v1 = make_float2(pd, pe); }
                else stash.overflow = 1;
            }
        }
        // the commit's recomputation deals the relationships differently: the whole group must take part
        if (tmax > 2 * G) stash.overflow = -group_min_int<G, kDeltaStr>(-stash.overflow);
    }

    // ---- one pass over the rows that did not move: (i) symmetry -- column update of the row minimum
    //      (exact: min(old, key(i, a), key(i, b)); a row whose remembered column moved is flagged for a
    //      rescan); (ii) clearance, part 2: the row's rectangle against the clearances sourced at a moved
    //      object, two such clearances per pass (a further pass only if some chain of the warp moved more) ------
    unsigned flags = 0;
    {
        const int tmax = __reduce_max_sync(FULL, moved_clearance_count(P, a, b));
        const float4 nbx = mvb ? nb : na;
        int t0 = 0;
        do {
            const MovedClearances mc = moved_clearances(P, a, b, oa, ob, na, nb, t0);
            int p = 0;
#pragma unroll kDuRow
            for (int i = g; i < n; i += G, p++) {
                const float4 pi = Pc[i * CPW];
                const bool moved = i == a || i == b;
                if (t0 == 0 && !moved)                          // (rows a, b are rescanned below)
                    sym_col_update<G>(h, D, c, sel, i, p, pi, a, b, mvb, any_b, na, nbx, pi_f, flags);
                if (tmax > 0) {
                    const int4 bi = box_at_q(P.obj_boxq[i], P.obj_v0xq[i], to_fixed(h, pi.x), to_fixed(h, pi.y));
                    long long dd = overlap_q(bi, mc.mn0) - overlap_q(bi, mc.mo0);
                    if (tmax > 1) dd += overlap_q(bi, mc.mn1) - overlap_q(bi, mc.mo1);
                    if (!moved) d_clr += dd;
                }
            }
            t0 += 2;
        } while (t0 < tmax);
    }

    // ---- symmetry: rescans of the moved and the flagged rows, then the sum ------------------------------
    const float sym_total = sym_rescan_sum<G>(P, S, D, c, g, sel, a, b, flags);

    // ---- totals ---------------------------------------------------------------------------------------
    star.pw = cur.pw + group_sum<G, kDeltaStr>(d_pw);
    star.pa = cur.pa + group_sum<G, kDeltaStr>(d_pa);
    star.vbx = cur.vbx + group_sum<G, kDeltaStr>(d_vbx);
    star.vby = cur.vby + group_sum<G, kDeltaStr>(d_vby);
    star.focal = cur.focal + group_sum<G, kDeltaStr>(d_focal);
    star.clr_q = cur.clr_q + group_sum_q<G, kDeltaStr>(d_clr);
    star.surf = cur.surf + group_sum<G, kDeltaStr>(d_surf);
    RawTerms t;
    t.pw = star.pw; t.pa = star.pa; t.vbx = star.vbx; t.vby = star.vby; t.focal = star.focal; t.surf = star.surf;
    t.clr_q = star.clr_q;
    t.clr = clearance_value(h, star.clr_q);
    t.sym = sym_total;
    t.off = 0.f;
    return combine(h, t).total;
}

// The proposal was accepted: bring the relationship memo up to the new layout (the KM memo is switched by
// flipping `sel`; the clearance sum travels in RunSums).
template <int G>
__device__ __forceinline__ void delta_commit(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                             const int a, const int b, const RelStash &stash)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    const float4 *Pc = S.P4 + c;
    if (stash.r0 >= 0) D.pr(stash.r0, c) = stash.v0;
    if (stash.r1 >= 0) D.pr(stash.r1, c) = stash.v1;
    if (stash.overflow) {                                       // rare: more than 2 G touched relationships
        for (int which = 0; which < 2; which++) {
            const int m = which ? b : a;
            if (m < 0) continue;
            for (int t = P.rel_adj_off[m] + g; t < P.rel_adj_off[m + 1]; t += G) {
                const int r = P.rel_adj[t];
                float pd, pe;
                rel_pen<CPW>(P, Pc, r, pd, pe);
                D.pr(r, c) = make_float2(pd, pe);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Relationship memo of the memo form (exact_eval): PR[r] holds the penalties of every relationship in the CURRENT layout.  A
// proposal re-evaluates only the relationships that name a moved object (CSR by object; one that names both
// is taken from a's list only), dealt round-robin to the lanes of the group, and overwrites their entries;
// then every lane adds ITS relationships r = g, g+G, ... from the memo -- the values and the order of
// eval_terms' relationship loop, hence the same bits.  A rejected proposal puts the old entries back.
// (The same memo inside the plain scan kernel, for rooms below 28 objects, was measured slower than evaluating
// all R relationships: 1.7e9 against 2.4e9 proposals/s at n = 16 -- with 16 chains per warp the warp pays
// for the chain with most touched relationships, and the kernel started to spill.)
struct RelMemoStash {
    int r0, r1;      // relationships this lane overwrote ...
    float2 o0, o1;   // ... and their penalties in the current layout
    int overflow;    // some lane of the group overwrote more than two: rejection recomputes
};

// PRc = the chain's memo (entry r at PRc[r * CPW]); Pc = its state, holding the proposal; a < 0: no move.
// Every lane of the warp must call this.  Returns the two sums reduced over the group.
template <int G, bool STR>
__device__ __forceinline__ void rel_memo_eval(const SmemProblem &P, const float4 *Pc, float2 *PRc, const int g, const int a, const int b,
                                              RelMemoStash &stash, float &pw_total, float &pa_total)
{
    constexpr int CPW = 32 / G;
    constexpr unsigned FULL = 0xffffffffu;
    const int R = P.h->R;
    const bool mva = a >= 0, mvb = b >= 0;
    stash.r0 = stash.r1 = -1;
    stash.overflow = 0;
    {
        const int ra0 = mva ? P.rel_adj_off[a] : 0, na_r = mva ? P.rel_adj_off[a + 1] - ra0 : 0;
        const int rb0 = mvb ? P.rel_adj_off[b] : 0, nb_r = mvb ? P.rel_adj_off[b + 1] - rb0 : 0;
        const int tot = na_r + nb_r;
        const int tmax = __reduce_max_sync(FULL, tot);
        int slot = 0;
        for (int tt = g; tt < tmax; tt += G, slot++) {
            int r = -1;
            if (tt < na_r) {
                r = P.rel_adj[ra0 + tt];
            } else if (tt < tot) {
                r = P.rel_adj[rb0 + tt - na_r];
                const int4 id = P.rel_idx[r];
                if (id.x == a || id.y == a || id.z == a || id.w == a) r = -1;
            }
            MH_CHECK(r < R && (R > 0 || tmax == 0));
            float pd, pe;
            rel_pen<CPW>(P, Pc, r >= 0 ? r : 0, pd, pe);
            if (r >= 0) {
                const float2 old = PRc[r * CPW];
                PRc[r * CPW] = make_float2(pd, pe);
                if (slot == 0) { stash.r0 = r; stash.o0 = old; }
                else if (slot == 1) { stash.r1 = r; stash.o1 = old; }
                else stash.overflow = 1;
            }
        }
        if (tmax > 2 * G) stash.overflow = -group_min_int<G, STR>(-stash.overflow);
    }
    __syncwarp();
    float pw = 0.f, pa = 0.f;
#pragma unroll 1
    for (int r = g; r < R; r += G) {
        const float2 v = PRc[r * CPW];
        pw += v.x;
        pa += v.y;
    }
    pw_total = group_sum<G, STR>(pw);
    pa_total = group_sum<G, STR>(pa);
}

// The proposal was rejected and Pc holds the current layout again: put the memo back.
template <int G>
__device__ __forceinline__ void rel_memo_restore(const SmemProblem &P, const float4 *Pc, float2 *PRc, const int g, const int a, const int b,
                                                 const RelMemoStash &stash)
{
    constexpr int CPW = 32 / G;
    if (stash.r0 >= 0) PRc[stash.r0 * CPW] = stash.o0;
    if (stash.r1 >= 0) PRc[stash.r1 * CPW] = stash.o1;
    if (stash.overflow) {
        for (int which = 0; which < 2; which++) {
            const int m = which ? b : a;
            if (m < 0) continue;
            for (int t = P.rel_adj_off[m] + g; t < P.rel_adj_off[m + 1]; t += G) {
                const int r = P.rel_adj[t];
                float pd, pe;
                rel_pen<CPW>(P, Pc, r, pd, pe);
                PRc[r * CPW] = make_float2(pd, pe);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// MH_EVAL_MEMO: full evaluation, bit for bit, at a fraction of the work.  Every term of the proposal is
// what eval_terms computes for it:
//   symmetry       the exact memo above (row minima; min is exact), summed in eval_terms' order;
//   relationships  only those that name a moved object are recomputed (into the PR memo); every lane
//                  then adds ITS relationships r = g, g+G, ... from the memo, as eval_terms does;
//   surface        per-rectangle values in the SV memo (only the moved objects' change), re-added in
//                  eval_terms' order: clearances k = g, g+G, ..., then objects i = g, g+G, ...;
//   visual balance, focal   re-added from the state itself (two FMAs and an add per object);
//   clearance      the integer sum, updated by the pairs the move touches (order-free, hence exact).
// Nothing is a floating-point running sum, so nothing drifts and there is no periodic rebuild.
template <int G> struct ExactStash {
    static constexpr int JOBS = (4 + G - 1) / G;   // surface jobs per lane (4 in all: objects a, b, clearances a, b)
    RelMemoStash rel;     // the PR entries this lane overwrote
    float sv_old[JOBS];   // the SV entries this lane overwrote
};

// SV slot of surface job 0..3 (object a, object b, clearance a, clearance b; quirk Q7: clearance k sits at
// object k's position), or -1 if there is no such rectangle.
__device__ __forceinline__ int sv_slot(const int job, const int a, const int b, const int n, const int C)
{
    const int m = (job & 1) ? b : a;
    if (m < 0) return -1;
    if (job >= 2) return m < C ? n + m : -1;
    return m;
}

// clr_cur: the integer clearance sum of the current layout; clr_star receives the proposal's.
template <int G>
__device__ __forceinline__ float exact_eval(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                            const int sel, const int a, const int b, const float4 oa, const float4 ob, const float4 na,
                                            const float4 nb, const long long clr_cur, long long &clr_star, ExactStash<G> &stash)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    constexpr unsigned FULL = 0xffffffffu;
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C;
    const float4 *Pc = S.P4 + c;
    const bool mvb = b >= 0;
    const bool any_b = __any_sync(FULL, mvb);
    const float pi_f = 0.5f * h->two_pi;
    RawTerms t;
    MH_CHECK(a >= 0 && a < n && b >= -1 && b < n && b != a && (sel == 0 || sel == 1));

    // ---- surface values of the moved rectangles, one job per lane -------------------------------------------
#pragma unroll
    for (int q = 0; q < ExactStash<G>::JOBS; q++) {
        const int job = g + q * G;
        const int slot = job < 4 ? sv_slot(job, a, b, n, C) : -1;
        if (slot >= 0) {
            const int m = (job & 1) ? b : a;
            const float4 pm = (job & 1) ? nb : na;
            const float4 kb = job >= 2 ? P.clr_box[m] : P.obj_box[m];
            const float v0 = job >= 2 ? P.clr_v0x[m] : P.obj_v0x[m];
            stash.sv_old[q] = D.sv(slot, c);
            D.sv(slot, c) = outside_room(box_at(kb, v0, pm.x, pm.y), h);
        }
    }

    // ---- relationships: memo update and the two sums (synchronises the warp: SV, written above, is read by
    //      other lanes below) -----------------------------------------------------------------------------------
    rel_memo_eval<G, kDeltaStr>(P, Pc, D.PR + c, g, a, b, stash.rel, t.pw, t.pa);

    // ---- clearance, part 1: every clearance against the moved objects ----------------------------------------
    long long d_clr = clearance_delta_moved_rows<G>(P, Pc, g, a, b, any_b, oa, ob, na, nb);

    // ---- one pass over the lane's rows: the float sums in eval_terms' order; for the rows that did not move the
    //      symmetry column update and clearance part 2 (the clearances sourced at a moved object) ------------------
    unsigned flags = 0;
    {
        float surf = 0.f, vbx = 0.f, vby = 0.f, focal = 0.f;
#pragma unroll kDuSum
        for (int k = g; k < C; k += G)
            surf += D.sv(n + k, c);
        const int tmax = __reduce_max_sync(FULL, moved_clearance_count(P, a, b));
        const float4 nbx = mvb ? nb : na;
        int t0 = 0;
        do {
            const MovedClearances mc = moved_clearances(P, a, b, oa, ob, na, nb, t0);
            int p = 0;
#pragma unroll kDuRow
            for (int i = g; i < n; i += G, p++) {
                const float4 pi = Pc[i * CPW];
                const bool moved = i == a || i == b;
                if (t0 == 0) {
                    const float area = P.obj_area[i];
                    vbx = fmaf(area, pi.x, vbx);
                    vby = fmaf(area, pi.y, vby);
                    focal += pi.w;
                    surf += D.sv(i, c);
                    if (!moved) sym_col_update<G>(h, D, c, sel, i, p, pi, a, b, mvb, any_b, na, nbx, pi_f, flags);
                }
                if (tmax > 0) {
                    const int4 bi = box_at_q(P.obj_boxq[i], P.obj_v0xq[i], to_fixed(h, pi.x), to_fixed(h, pi.y));
                    long long dd = overlap_q(bi, mc.mn0) - overlap_q(bi, mc.mo0);
                    if (tmax > 1) dd += overlap_q(bi, mc.mn1) - overlap_q(bi, mc.mo1);
                    if (!moved) d_clr += dd;
                }
            }
            t0 += 2;
        } while (t0 < tmax);
        t.vbx = group_sum<G, kDeltaStr>(vbx);
        t.vby = group_sum<G, kDeltaStr>(vby);
        t.focal = group_sum<G, kDeltaStr>(focal);
        t.surf = group_sum<G, kDeltaStr>(surf);
        t.off = 0.f;
    }
    clr_star = clr_cur + group_sum_q<G, kDeltaStr>(d_clr);
    t.clr_q = clr_star;
    t.clr = clearance_value(h, clr_star);

    // ---- symmetry ------------------------------------------------------------------------------------------------
    t.sym = sym_rescan_sum<G>(P, S, D, c, g, sel, a, b, flags);
    return combine(h, t).total;
}

// The proposal was rejected (S.P4 is restored): put the surface values and the relationship memo back.
template <int G>
__device__ __forceinline__ void exact_reject(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                             const int a, const int b, const ExactStash<G> &stash)
{
    const int n = P.h->n, C = P.h->C;
    const float4 *Pc = S.P4 + c;
#pragma unroll
    for (int q = 0; q < ExactStash<G>::JOBS; q++) {
        const int job = g + q * G;
        const int slot = job < 4 ? sv_slot(job, a, b, n, C) : -1;
        if (slot >= 0) D.sv(slot, c) = stash.sv_old[q];
    }
    rel_memo_restore<G>(P, Pc, D.PR + c, g, a, b, stash.rel);
}

} // namespace mh
