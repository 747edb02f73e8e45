// mh_delta.cuh -- incremental ("delta") evaluation of a proposal: MH_EVAL_DELTA.
//
// A proposal moves one object (translate, rotate) or two (swap).  Full evaluation redoes
// n^2 + C n pair terms for it; here only what the moved objects touch is recomputed:
//
//   symmetry   sum_i (5 - min_j key(i,j)) -- every row keeps its minimum and ONE column that attains
//              it (KM memo).  A move of object m changes row m (rescanned by the whole group) and
//              column m of every other row: the row minimum becomes min(old, key(i,m)) unless the
//              remembered column was m itself, in which case that row is rescanned too.  min is exact,
//              so the memo always equals what a full evaluation computes -- bit for bit.
//   clearance  pairs (k, m) for every clearance k, and pairs (k', i) for every clearance k' whose
//              source object is m: new overlap - old overlap.
//   surface    the moved objects' own rectangles and the clearances with the same INDEX (quirk Q7).
//   focal      difference of the memoised cosines;  visual balance: area * displacement.
//   pair-wise  relationships that name a moved object; their old penalties come from the PR memo.
//
// The additive terms are kept as running sums, so their rounding differs from a full evaluation
// and drifts; every kRefresh iterations (and at every launch) the memo and the sums are rebuilt
// from scratch.  Delta mode is therefore statistically, not bitwise, equivalent to full evaluation
// (tests: KS against the oracle, running total against a fresh evaluation); the costs a caller
// receives always come from mh_score_kernel, i.e. from a full evaluation of the emitted layout.
#pragma once
#include "mh_costs.cuh"

namespace mh {

constexpr int kRefresh = 128; // iterations between full rebuilds of the memo and the running sums

template <int G> struct DeltaState {
    static constexpr int CPW = 32 / G;
    float2 *KM; // [2][n][CPW] {row minimum of key, a column attaining it (int bits; -1 = none below 5)}
    float2 *PR; // [R][CPW]    {distance penalty, angle penalty} of every relationship
    int n;
    // with_pr = false: only the symmetry memo (MH_EVAL_MEMO)
    __host__ __device__ static int words(int n, int R, bool with_pr = true) { return CPW * (4 * n + (with_pr ? 2 * R : 0)); }
    __device__ __forceinline__ void bind(float *base, int n_, int /*R*/)
    {
        n = n_;
        KM = reinterpret_cast<float2 *>(base);
        PR = KM + 2 * n * CPW;
    }
    __device__ __forceinline__ float2 &km(int sel, int i, int c) const { return KM[(sel * n + i) * CPW + c]; }
    __device__ __forceinline__ float2 &pr(int r, int c) const { return PR[r * CPW + c]; }
};

// Committed running sums of the additive terms (positive magnitudes, as in RawTerms).
struct RunSums {
    float pw, pa, vbx, vby, focal, clr, surf;
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

// Relationship penalties a proposal recomputed, kept in registers until the accept decision.
struct RelStash {
    int r0, r1, r2, r3;
    float2 v0, v1, v2, v3;
    int overflow; // more than four touched relationships in this lane: commit recomputes
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

// Symmetry term of the proposal that moved objects a and b (-1 = none) to na / nb (S.P4 already holds
// them), from the memo of the current layout in KM[sel]; writes the proposal's memo to KM[1-sel] and
// returns sum_i (5 - min_j key(i,j)) reduced over the group.  The row minima are EXACT (min is exact),
// and each lane adds its rows in increasing row order, exactly as the full scan of eval_terms does, so
// the value is bit-identical to a full evaluation.  Every lane of the warp must call this.
template <int G, bool STR>
__device__ __forceinline__ float sym_memo_eval(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                               const int sel, const int a, const int b, const float4 na, const float4 nb)
{
    constexpr int CPW = WarpState<G>::CPW;
    const mhProblemHeader *h = P.h;
    const int n = h->n;
    const float pi_f = 0.5f * h->two_pi;
    const float4 *Pc = S.P4 + c;
    const bool mva = a >= 0, mvb = b >= 0;
    auto inM = [&](int i) { return i == a || (mvb && i == b); };

    // ---- column update of the rows that did not move ---------------------------------------------------
    unsigned flags = 0;
    {
        int p = 0;
        for (int i = g; i < n; i += G, p++) {
            if (mva && inM(i)) continue;                        // rescanned below
            const float2 km = D.km(sel, i, c);
            float k = km.x;
            int arg = __float_as_int(km.y);
            if (mva) {
                if (arg == a || (mvb && arg == b)) {            // the remembered best column moved
                    flags |= 1u << p;
                    continue;
                }
                const RowRef rr = sym_row(h, Pc[i * CPW]);
                const float k1 = sym_key(rr, na, pi_f);
                if (k1 < k) { k = k1; arg = a; }
                if (mvb) {
                    const float k2 = sym_key(rr, nb, pi_f);
                    if (k2 < k) { k = k2; arg = b; }
                }
            }
            D.km(1 - sel, i, c) = make_float2(k, __int_as_float(arg));
        }
    }
    MH_PHASE_SYNC(3);
    // ---- rows rescanned by the whole group.  The moved rows a and b share one pass over the columns;
    //      rows whose remembered column moved are taken one per trip, each group picking its own next
    //      row, so the warp pays for the longest group queue. -----------------------------------------------
    if (__any_sync(0xffffffffu, mva)) {
        const RowRef ra = sym_row(h, mva ? na : Pc[0]), rb = sym_row(h, mvb ? nb : Pc[0]);
        float ka = 5.0f, kb2 = 5.0f;
        int aa = -1, ab = -1;
        for (int j = g; j < n; j += G) {
            const float4 q = Pc[j * CPW];
            const float k1 = sym_key(ra, q, pi_f), k2 = sym_key(rb, q, pi_f);
            if (k1 < ka) { ka = k1; aa = j; }
            if (k2 < kb2) { kb2 = k2; ab = j; }
        }
        group_argmin<G, STR>(ka, aa);
        group_argmin<G, STR>(kb2, ab);
        if (g == 0) {
            if (mva) D.km(1 - sel, a, c) = make_float2(ka, __int_as_float(aa));
            if (mvb) D.km(1 - sel, b, c) = make_float2(kb2, __int_as_float(ab));
        }
    }
    for (;;) {
        const int mine = flags ? g + (__ffs(flags) - 1) * G : 0x7fffffff;
        const int row = group_min_int<G, STR>(mine);
        if (!__any_sync(0xffffffffu, row != 0x7fffffff)) break;
        const bool act = row != 0x7fffffff;
        if (act && mine == row) flags &= flags - 1;
        float k;
        int arg;
        sym_scan<CPW>(sym_row(h, Pc[(act ? row : 0) * CPW]), Pc, n, g, G, pi_f, k, arg);
        group_argmin<G, STR>(k, arg);
        if (act && g == 0) D.km(1 - sel, row, c) = make_float2(k, __int_as_float(arg));
    }
    __syncwarp();
    // ---- the sum, in the row order of the full scan --------------------------------------------------------
    float s = 0.f;
    for (int i = g; i < n; i += G)
        s += 5.0f - D.km(1 - sel, i, c).x;
    return group_sum<G, STR>(s);
}

// Rebuild the memo and the running sums of the CURRENT layout from scratch; returns its total.
template <int G>
__device__ __forceinline__ float delta_rebuild(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, int c, int g, int sel,
                                               RunSums &cur)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    const mhProblemHeader *h = P.h;
    const int R = h->R;
    RawTerms t;
    eval_terms<G, false, kDeltaStr>(P, S, c, g, t); // also refreshes S.CB
    cur.pw = t.pw; cur.pa = t.pa; cur.vbx = t.vbx; cur.vby = t.vby; cur.focal = t.focal; cur.clr = t.clr; cur.surf = t.surf;
    const float4 *Pc = S.P4 + c;
    sym_memo_build<G>(P, S, D, c, g, sel);
    for (int r = g; r < R; r += G) {
        float pd, pa;
        rel_pen<CPW>(P, Pc, r, pd, pa);
        D.pr(r, c) = make_float2(pd, pa);
    }
    __syncwarp();
    return combine(h, t).total;
}

// Evaluate the proposal that moved objects a (and b; -1 = none) from oa/ob to na/nb; S.P4 already
// holds the new state.  Writes the proposal's KM memo into buffer 1-sel, returns its total and
// its running sums.  Every lane of the warp must call this.
template <int G>
__device__ __forceinline__ float delta_eval(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                            const int sel, const int a, const int b, const float4 oa, const float4 ob, const float4 na,
                                            const float4 nb, const RunSums &cur, RunSums &star, RelStash &stash)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C, R = h->R;
    const float4 *Pc = S.P4 + c, *CBc = S.CB + c;
    const bool mva = a >= 0, mvb = b >= 0;
    auto inM = [&](int i) { return i == a || (mvb && i == b); };

    float d_pw = 0.f, d_pa = 0.f, d_vbx = 0.f, d_vby = 0.f, d_focal = 0.f, d_clr = 0.f, d_surf = 0.f;

    // ---- the moved objects' own rectangles ----------------------------------------------------------
    float4 box_oa = make_float4(0.f, 0.f, 0.f, 0.f), box_na = box_oa, box_ob = box_oa, box_nb = box_oa;
    if (mva) {
        const float4 kb = P.obj_box[a];
        const float v0 = P.obj_v0x[a];
        box_oa = box_at(kb, v0, oa.x, oa.y);
        box_na = box_at(kb, v0, na.x, na.y);
        if (g == 0) {
            const float area = P.obj_area[a];
            d_surf += outside_room(box_na, h) - outside_room(box_oa, h);
            d_focal += na.w - oa.w;
            d_vbx += area * (na.x - oa.x);
            d_vby += area * (na.y - oa.y);
        }
    }
    if (mvb) {
        const float4 kb = P.obj_box[b];
        const float v0 = P.obj_v0x[b];
        box_ob = box_at(kb, v0, ob.x, ob.y);
        box_nb = box_at(kb, v0, nb.x, nb.y);
        if (g == 0) {
            const float area = P.obj_area[b];
            d_surf += outside_room(box_nb, h) - outside_room(box_ob, h);
            d_focal += nb.w - ob.w;
            d_vbx += area * (nb.x - ob.x);
            d_vby += area * (nb.y - ob.y);
        }
    }

    MH_PHASE_SYNC(3);
    // ---- clearances: pairs (k, moved object) for every k; Q7 surface of clearance INDEX a / b ------
    if (mva) {
        for (int k = g; k < C; k += G) {
            const int src = P.clr_src[k];
            const float4 kb = P.clr_box[k];
            const float v0 = P.clr_v0x[k];
            const float4 cb_old = CBc[k * CPW];
            float4 cb_new = cb_old;
            if (inM(src)) {
                const float4 ps = Pc[src * CPW];
                cb_new = box_at(kb, v0, ps.x, ps.y);
            }
            d_clr += overlap(box_na, cb_new) - overlap(box_oa, cb_old);
            if (mvb) d_clr += overlap(box_nb, cb_new) - overlap(box_ob, cb_old);
            if (k == a) d_surf += outside_room(box_at(kb, v0, na.x, na.y), h) - outside_room(box_at(kb, v0, oa.x, oa.y), h);
            if (mvb && k == b) d_surf += outside_room(box_at(kb, v0, nb.x, nb.y), h) - outside_room(box_at(kb, v0, ob.x, ob.y), h);
        }
        // ---- clearances sourced at a moved object against every object that did not move -----------
        for (int which = 0; which < 2; which++) {
            const int m = which ? b : a;
            if (m < 0) continue;
            const int t0 = P.clr_adj_off[m], t1 = P.clr_adj_off[m + 1];
            const float4 pm = which ? nb : na;
            for (int t = t0; t < t1; t++) {
                const int k = P.clr_adj[t];
                const float4 cb_old = CBc[k * CPW];
                const float4 cb_new = box_at(P.clr_box[k], P.clr_v0x[k], pm.x, pm.y);
                for (int i = g; i < n; i += G) {
                    if (inM(i)) continue;
                    const float4 q = Pc[i * CPW];
                    const float4 bi = box_at(P.obj_box[i], P.obj_v0x[i], q.x, q.y);
                    d_clr += overlap(bi, cb_new) - overlap(bi, cb_old);
                }
            }
        }
    }

    MH_PHASE_SYNC(3);
    // ---- relationships that name a moved object: first collect them per lane, then evaluate slot by
    //      slot, so that the warp pays for the deepest lane queue and not for every loop trip in
    //      which some lane happens to hold one ------------------------------------------------------------
    stash.r0 = stash.r1 = stash.r2 = stash.r3 = -1;
    stash.overflow = 0;
    if (mva) {
        int nq = 0;
        for (int r = g; r < R; r += G) {
            const int4 id = P.rel_idx[r];
            if (inM(id.x) || inM(id.y) || inM(id.z) || inM(id.w)) {
                if (nq == 0) stash.r0 = r;
                else if (nq == 1) stash.r1 = r;
                else if (nq == 2) stash.r2 = r;
                else if (nq == 3) stash.r3 = r;
                else {                                          // rare: evaluate in place, commit recomputes
                    float pd, pe;
                    rel_pen<CPW>(P, Pc, r, pd, pe);
                    const float2 old = D.pr(r, c);
                    d_pw += pd - old.x;
                    d_pa += pe - old.y;
                    stash.overflow = 1;
                }
                nq++;
            }
        }
    }
#define MH_REL_SLOT(RQ, VQ)                                   \
    if (RQ >= 0) {                                            \
        float pd, pe;                                         \
        rel_pen<CPW>(P, Pc, RQ, pd, pe);                      \
        const float2 old = D.pr(RQ, c);                       \
        d_pw += pd - old.x;                                   \
        d_pa += pe - old.y;                                   \
        VQ = make_float2(pd, pe);                             \
    }
    MH_REL_SLOT(stash.r0, stash.v0)
    MH_REL_SLOT(stash.r1, stash.v1)
    MH_REL_SLOT(stash.r2, stash.v2)
    MH_REL_SLOT(stash.r3, stash.v3)
#undef MH_REL_SLOT

    MH_PHASE_SYNC(3);
    // ---- symmetry: exact memo (sym_memo_eval) ---------------------------------------------------------------
    const float sym_total = sym_memo_eval<G, kDeltaStr>(P, S, D, c, g, sel, a, b, na, nb);

    // ---- totals ---------------------------------------------------------------------------------------
    star.pw = cur.pw + group_sum<G, kDeltaStr>(d_pw);
    star.pa = cur.pa + group_sum<G, kDeltaStr>(d_pa);
    star.vbx = cur.vbx + group_sum<G, kDeltaStr>(d_vbx);
    star.vby = cur.vby + group_sum<G, kDeltaStr>(d_vby);
    star.focal = cur.focal + group_sum<G, kDeltaStr>(d_focal);
    star.clr = cur.clr + group_sum<G, kDeltaStr>(d_clr);
    star.surf = cur.surf + group_sum<G, kDeltaStr>(d_surf);
    RawTerms t;
    t.pw = star.pw; t.pa = star.pa; t.vbx = star.vbx; t.vby = star.vby; t.focal = star.focal; t.clr = star.clr; t.surf = star.surf;
    t.sym = sym_total;
    t.off = 0.f;
    return combine(h, t).total;
}

// The proposal was accepted: bring the clearance rectangles and the relationship memo up to the
// new layout (the KM memo is switched by flipping `sel`).
template <int G>
__device__ __forceinline__ void delta_commit(const SmemProblem &P, const WarpState<G> &S, const DeltaState<G> &D, const int c, const int g,
                                             const int a, const int b, const RelStash &stash)
{
    using WS = WarpState<G>;
    constexpr int CPW = WS::CPW;
    const int R = P.h->R;
    const float4 *Pc = S.P4 + c;
    if (a < 0) return;
    const bool mvb = b >= 0;
    auto inM = [&](int i) { return i == a || (mvb && i == b); };
    for (int which = 0; which < 2; which++) {
        const int m = which ? b : a;
        if (m < 0) continue;
        const float4 pm = Pc[m * CPW];
        for (int t = P.clr_adj_off[m] + g; t < P.clr_adj_off[m + 1]; t += G) {
            const int k = P.clr_adj[t];
            S.CB[WS::at(k, c)] = box_at(P.clr_box[k], P.clr_v0x[k], pm.x, pm.y);
        }
    }
    if (stash.r0 >= 0) D.pr(stash.r0, c) = stash.v0;
    if (stash.r1 >= 0) D.pr(stash.r1, c) = stash.v1;
    if (stash.r2 >= 0) D.pr(stash.r2, c) = stash.v2;
    if (stash.r3 >= 0) D.pr(stash.r3, c) = stash.v3;
    if (stash.overflow) {
        for (int r = g; r < R; r += G) {
            const int4 id = P.rel_idx[r];
            if (inM(id.x) || inM(id.y) || inM(id.z) || inM(id.w)) {
                float pd, pe;
                rel_pen<CPW>(P, Pc, r, pd, pe);
                D.pr(r, c) = make_float2(pd, pe);
            }
        }
    }
}

} // namespace mh
