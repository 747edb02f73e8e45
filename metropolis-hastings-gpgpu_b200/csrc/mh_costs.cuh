// mh_costs.cuh -- the Merrell et al. cost terms of the reference (Kernel.cu:162-550), written
// for a group of G lanes that owns one layout ("chain").
//
// Data layout.  A warp holds CPW = 32/G chains.  Chain state lives in shared memory as one
// float4 per object, {x, y, rotY, f} with f = the object's memoised focal-point cosine,
// interleaved by chain:   P4[j*LD + c], LD >= CPW.   In the pair loops every lane of a group reads
// the same 16 bytes (a broadcast) with ONE LDS.128 whose address advances by the compile-time
// constant LD*16 per column, so an unrolled loop needs no address arithmetic; the CPW groups
// of a warp read CPW consecutive float4s.  In the lane-strided O(n) loops lane (c, g) touches
// object g + G*k -> float4 (g+Gk)*LD + c.  With the interleaved lane mapping (the memo kernels) LD = CPW
// and the 32 lanes cover 32 consecutive float4s.  With contiguous groups (the scan kernels) LD = CPW would
// put the G objects a group reads at a stride of CPW*16 bytes -- for CPW = 8 exactly one bank row apart, a
// 4-way bank conflict (ncu round 1: 25 % of the kernel's shared-memory wavefronts) -- so the rows are PADDED:
// LD = 20, 10, 5, 3 for G = 2, 4, 8, 16 makes the 8 lanes of every quarter-warp (the unit a 128-bit shared load
// is served in) fall into 8 different 16-byte bank groups.
//
// The O(n^2) and O(C n) terms are row-parallel: lane g owns rows i = g, g+G, ... and walks
// all columns, keeping the running max / sum in registers; one xor-shuffle tree per term per
// evaluation reduces over the group.
//
// Precision: float32 throughout (the reference mixes float and double per expression,
// SURVEY.md section 8a); every quirk that changes VALUES is kept: PI = 3.1416 (Q4), the
// off-limits term outside the total (Q5), the untranslated first vertex in the AABB min (Q6),
// clearance i moved by object i in the surface term (Q7), translate-only rectangles (Q8), the
// always-true angle condition (Q9), centroid/2 (Q11), one-sided angle wraps (Q18), the
// distance x angle product (Q20).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mh_abi.h"

#ifndef MH_SYM_UNROLL
#define MH_SYM_UNROLL 8
#endif

// compute-sanitizer is not available on the development pool: a build with -DMH_DEBUG_BOUNDS checks every
// index that comes from data (Philox draws, adjacency lists, memo columns) and every memo accessor, and
// traps on a violation; tools/build_variant.sh dbg -DMH_DEBUG_BOUNDS + MH_LIB run the GPU tests against it.
#ifdef MH_DEBUG_BOUNDS
#include <stdio.h>
#define MH_CHECK(cond)                                                                          \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            printf("MH_CHECK failed: %s at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                           \
        }                                                                                       \
    } while (0)
#else
#define MH_CHECK(cond) ((void)0)
#endif

#ifndef MH_RECIP_DENOM
#define MH_RECIP_DENOM 1
#endif

namespace mh {

constexpr int kSymUnroll = MH_SYM_UNROLL; // columns per trip of the symmetry loop
#ifndef MH_FUSE_UNROLL
#define MH_FUSE_UNROLL 2
#endif
constexpr int kFuseUnroll = MH_FUSE_UNROLL; // trips of the fused symmetry+clearance loop unrolled together

struct SmemProblem {
    const mhProblemHeader *h;
    const float4 *obj_box;
    const float *obj_v0x;
    const float *obj_area;
    const int *obj_frozen;
    const float4 *clr_box;
    const float *clr_v0x;
    const int *clr_src;
    const int *clr_adj_off;
    const int *clr_adj;
    const int *rel_adj_off;
    const int *rel_adj;
    const int4 *rel_idx;
    const float4 *rel_rng;
    const float4 *rel_aux;
    const int4 *obj_boxq;   // fixed-point copies of obj_box / obj_v0x / clr_box / clr_v0x (clearance term)
    const int *obj_v0xq;
    const int4 *clr_boxq;
    const int *clr_v0xq;
};

// base: the staged blob in shared memory (arrays); hdr: where the header's scalars are read from -- the blob itself,
// or the copy in the kernel's parameter space (constant bank) that the chain kernels carry.
__device__ __forceinline__ SmemProblem bind_problem(const float *base, const mhProblemHeader *hdr = nullptr)
{
    SmemProblem P;
    P.h = hdr ? hdr : reinterpret_cast<const mhProblemHeader *>(base);
    P.obj_box = reinterpret_cast<const float4 *>(base + P.h->off_obj_box);
    P.obj_v0x = base + P.h->off_obj_v0x;
    P.obj_area = base + P.h->off_obj_area;
    P.obj_frozen = reinterpret_cast<const int *>(base + P.h->off_obj_frozen);
    P.clr_box = reinterpret_cast<const float4 *>(base + P.h->off_clr_box);
    P.clr_v0x = base + P.h->off_clr_v0x;
    P.clr_src = reinterpret_cast<const int *>(base + P.h->off_clr_src);
    P.clr_adj_off = reinterpret_cast<const int *>(base + P.h->off_clr_adj_off);
    P.clr_adj = reinterpret_cast<const int *>(base + P.h->off_clr_adj);
    P.rel_adj_off = reinterpret_cast<const int *>(base + P.h->off_rel_adj_off);
    P.rel_adj = reinterpret_cast<const int *>(base + P.h->off_rel_adj);
    P.rel_idx = reinterpret_cast<const int4 *>(base + P.h->off_rel_idx);
    P.rel_rng = reinterpret_cast<const float4 *>(base + P.h->off_rel_rng);
    P.rel_aux = reinterpret_cast<const float4 *>(base + P.h->off_rel_aux);
    P.obj_boxq = reinterpret_cast<const int4 *>(base + P.h->off_obj_boxq);
    P.obj_v0xq = reinterpret_cast<const int *>(base + P.h->off_obj_v0xq);
    P.clr_boxq = reinterpret_cast<const int4 *>(base + P.h->off_clr_boxq);
    P.clr_v0xq = reinterpret_cast<const int *>(base + P.h->off_clr_v0xq);
    return P;
}

// Row stride (in float4) of the padded layout for contiguous lane groups: see "Data layout" above.
__host__ __device__ constexpr int padded_ld(int G) { return G == 2 ? 20 : G == 4 ? 10 : G == 8 ? 5 : G == 16 ? 3 : 32 / G; }

// Per-warp chain state in shared memory.  PAD: padded rows (the scan kernels, contiguous lane groups).
template <int G, bool PAD = false> struct WarpState {
    static constexpr int CPW = 32 / G;
    static constexpr int LD = PAD ? padded_ld(G) : CPW;
    float4 *P4;     // [n][LD] {x, y, rotY, focal cosine}
    int4 *CB;       // [C][LD] clearance AABBs (fixed point) of the layout under evaluation
    __device__ __forceinline__ static int at(int j, int c) { return j * LD + c; }
    // words of shared memory one warp needs
    __host__ __device__ static int words(int n, int C) { return LD * (4 * n + 4 * C); }
    __device__ __forceinline__ void bind(float *base, int n, int C)
    {
        P4 = reinterpret_cast<float4 *>(base);
        CB = reinterpret_cast<int4 *>(P4 + n * LD);
    }
};

// Sums are kept as POSITIVE magnitudes; the reference's "result -= ..." sign is applied once
// in combine().
struct RawTerms {
    float pw;    // sum of pair-wise distance penalties        (Kernel.cu:210-233)
    float pa;    // sum of pair-wise angle penalties           (Kernel.cu:236-263)
    float vbx;   // sum area_i * x_i                           (Kernel.cu:197-203)
    float vby;
    float focal; // sum cos(phi_i)                             (Kernel.cu:266-281)
    float sym;   // sum_i max(0, max_j ...)                    (Kernel.cu:283-318)
    float clr;   // sum of clearance x off-limit overlaps      (Kernel.cu:404-434) = clr_q * 2^(-2 clr_k)
    long long clr_q; // the same as the integer it is computed as
    float surf;  // sum of areas outside the room              (Kernel.cu:437-483)
    float off;   // sum of off-limit x off-limit overlaps      (Kernel.cu:485-514)
};

struct Costs8 {
    float total, pair, visual, focal, sym, clr, off, surf; // field order of resultCosts
};

// Lane -> (chain c, lane-in-group g).  Contiguous groups (c = lane / G) make the broadcast loads of
// the pair loops cheapest (a quarter-warp touches few distinct float4s); the delta kernel, whose
// loads are lane-strided, uses the interleaved mapping (c = lane % CPW), for which 32 lanes read 32
// consecutive float4s without bank conflicts.  STR selects the interleaved mapping.
template <int G, bool STR> struct LaneMap {
    static constexpr int CPW = 32 / G;
    __device__ __forceinline__ static int chain(int lane) { return STR ? lane % CPW : lane / G; }
    __device__ __forceinline__ static int lane_in_group(int lane) { return STR ? lane / CPW : lane % G; }
    __device__ __forceinline__ static int first_lane(int c) { return STR ? c : c * G; }
    static constexpr int xor_step = STR ? CPW : 1; // lane distance between neighbouring lanes of a group
};

template <int G, bool STR = false> __device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, m * LaneMap<G, STR>::xor_step);
    return v;
}

__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float pos(float v) { return fmaxf(v, 0.0f); }

// AABB of a rectangle translated by (tx, ty): Kernel.cu:366-401 with the four-vertex min/max
// hoisted (min(a+t, b+t) == min(a,b)+t in IEEE arithmetic).  Q6: the first vertex's x enters
// the minimum untranslated.
__device__ __forceinline__ float4 box_at(float4 k, float v0x, float tx, float ty)
{
    return make_float4(fminf(v0x, k.x + tx), k.y + ty, k.z + tx, k.w + ty);
}

// Kernel.cu:321-340: zero unless both extents are positive.
__device__ __forceinline__ float overlap(float4 a, float4 b)
{
    const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    const float h = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    return pos(w) * pos(h);
}

// ---- the clearance term is an integer sum ---------------------------------------------------------------------
// ClearanceCosts (Kernel.cu:404-434) = sum over C x n pairs of AABB overlaps.  It is evaluated in FIXED POINT:
// coordinates in units of 2^-k (k = mhProblemHeader.clr_k, chosen on the host so that nothing can overflow;
// 5e-7 .. 6e-8 length units at the BASELINE rooms, finer than the float32 ulp of the coordinates it replaces),
// areas accumulated in a 64-bit integer.  Per pair that is the same instruction count as the float form (4 min/max,
// 2 subtractions, 2 max-with-zero, one multiply-add), but integer addition is ASSOCIATIVE: the sum does not depend on
// the order of its terms.  A proposal can therefore update the sum by the few pairs it touches (the moved objects'
// rows and the columns of the clearances they carry) and hold, bit for bit, what a from-scratch evaluation of the
// new layout computes -- no running-sum drift, no periodic rebuild, and the memo kernel equals the plain scan exactly.
// (no clamp: every position a chain can hold lies inside the bound the host derived the scale from -- translations
// snap to the room, swaps permute -- and the conversion itself saturates)
__device__ __forceinline__ int to_fixed(const mhProblemHeader *h, float v)
{
    return __float2int_rn(v * h->clr_scale);                   // 2^k: the product is exact
}

// AABB of a rectangle translated by the fixed-point position (tx, ty); Q6 as in box_at.
__device__ __forceinline__ int4 box_at_q(int4 k, int v0x, int tx, int ty)
{
    return make_int4(min(v0x, k.x + tx), k.y + ty, k.z + tx, k.w + ty);
}

// Kernel.cu:321-340 in fixed point: zero unless both extents are positive.
__device__ __forceinline__ long long overlap_q(int4 a, int4 b)
{
    const int w = min(a.z, b.z) - max(a.x, b.x);
    const int h = min(a.w, b.w) - max(a.y, b.y);
    return (long long)max(w, 0) * (long long)max(h, 0);
}

__device__ __forceinline__ long long overlap_add_q(int4 a, int4 b, long long acc)
{
    const int w = min(a.z, b.z) - max(a.x, b.x);
    const int h = min(a.w, b.w) - max(a.y, b.y);
    return acc + (long long)max(w, 0) * (long long)max(h, 0);   // one IMAD.WIDE with a 64-bit addend
}

// Sum over all clearance rectangles (CBc[k * LD]) of their overlap with box a: one row of ClearanceCosts.
template <int LD> __device__ __forceinline__ long long clearance_row_q(const int4 a, const int4 *CBc, const int C)
{
    long long acc = 0;
    int k = 0;
#pragma unroll 2
    for (; k + 2 <= C; k += 2) {
        const int4 b0 = CBc[(k + 0) * LD], b1 = CBc[(k + 1) * LD];
        acc = overlap_add_q(a, b0, acc);
        acc = overlap_add_q(a, b1, acc);
    }
    for (; k < C; k++)
        acc = overlap_add_q(a, CBc[k * LD], acc);
    return acc;
}

template <int G, bool STR = false> __device__ __forceinline__ long long group_sum_q(long long v)
{
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, m * LaneMap<G, STR>::xor_step);
    return v;
}

// the float the cost function sees: ONE rounding of the exact integer sum (the scale is a power of two)
__device__ __forceinline__ float clearance_value(const mhProblemHeader *h, long long q) { return __ll2float_rn(q) * h->clr_unit; }

// Area of box b outside the room: the four complement rectangles of Kernel.cu:343-364 with
// +-DBL_MAX narrowed to +-inf by fmaxf/fminf (Kernel.cu:325-328), i.e. no clamp on that side.
__device__ __forceinline__ float outside_room(float4 b, const mhProblemHeader *h)
{
    const float rx0 = h->room_minx, ry0 = h->room_miny, rx1 = h->room_maxx, ry1 = h->room_maxy;
    const float w = pos(b.z - b.x);
    const float hm = pos(fminf(b.w, ry1) - fmaxf(b.y, ry0));
    float e = w * pos(fminf(b.w, ry0) - b.y);                 // below the room, all x
    e = __fadd_rn(e, pos(fminf(b.z, rx0) - b.x) * hm);        // left of it
    e = __fadd_rn(e, w * pos(b.w - fmaxf(b.y, ry1)));         // above
    e = __fadd_rn(e, pos(b.z - fmaxf(b.x, rx1)) * hm);        // right
    return e;
}

__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// cos(phi) of one object, phi = atan2(fy - y, fx - x) - rot + PI/2 (Kernel.cu:185-188, 271-277).
// A pure function of that object's state, so it is memoised in the .w lane of its float4 and
// recomputed only for the one or two objects a proposal moves.
// MH_FOCAL_DIRECT (default): cos(theta - psi) with theta = atan2(dy, dx), psi = rot - PI/2 is
// (dx cos psi + dy sin psi) / |d| -- one sincosf and one rsqrtf instead of atan2f and cosf: about half the
// instructions and half the code on the per-iteration path (which is what this kernel is sensitive to); the value
// differs from the reference's composition of two libm calls by rounding only (<= 3e-7 absolute).
#ifndef MH_FOCAL_DIRECT
#define MH_FOCAL_DIRECT 1
#endif
__device__ __noinline__ float focal_cos_impl(float fx, float fy, float half_pi, float x, float y, float rot)
{
#if MH_FOCAL_DIRECT
    const float dx = fx - x, dy = fy - y;
    const float d2 = fmaf(dx, dx, dy * dy);
    float sn, cs;
    sincosf(rot - half_pi, &sn, &cs);
    return d2 > 0.f ? fmaf(dx, cs, dy * sn) * rsqrtf(d2) : cs;   // atan2(0, 0) = 0: cos(-psi)
#else
    return cosf(atan2f(fy - y, fx - x) - rot + half_pi);
#endif
}
__device__ __forceinline__ float focal_cos(const mhProblemHeader *h, float x, float y, float rot)
{
    return focal_cos_impl(h->focal_x, h->focal_y, h->half_pi, x, y, rot);
}

// Symmetry (Kernel.cu:290-314).  Object i is reflected across the focal axis; the best match over
// all j is max_j (5 - sqrt(d) - 0.4 |dt|) = 5 - min_j key(i, j), key = sqrt(d) + 0.4 |dt|, floored at 0.
// Q18: dt = rot_j - rr wraps one-sidedly (dt > PI -> dt - 2 PI); for every dt the wrapped magnitude
// equals min(|dt|, |dt - 2 PI|) = ||dt - PI| - PI|, two adds with |.| operand modifiers and no
// compare/select.  sqrt(Distance) = (d^2)^(1/4) = rsqrt(rsqrt(d^2)): MUFU.RSQ issues faster than
// MUFU.SQRT on sm_100 (measured).
struct RowRef {
    float rx, ry, rrp; // reflected position; reflected rotation + PI
};

__device__ __forceinline__ RowRef sym_row(const mhProblemHeader *h, const float4 pi)
{
    const float s = 2.0f * (h->fdotu - (pi.x * h->ux + pi.y * h->uy));
    RowRef r;
    r.rx = pi.x + s * h->ux;
    r.ry = pi.y + s * h->uy;
    float rr = h->two_focal_rot - pi.z;
    if (rr < -h->pi_cmp) rr += h->two_pi;                      // Q18: one-sided wrap of the reflection
    r.rrp = rr + 0.5f * h->two_pi;                              // rot_j - rrp = dt - PI
    return r;
}

__device__ __forceinline__ float sym_key(const RowRef &r, const float4 q, const float pi_f)
{
    const float dx = q.x - r.rx, dy = q.y - r.ry;
    const float sd = rsqrt_approx(rsqrt_approx(fmaf(dx, dx, dy * dy)));
    const float e = q.z - r.rrp;
    const float w = fabsf(e) - pi_f;
    return fmaf(0.4f, fabsf(w), sd);
}

// atan2(y, x) in [-pi, pi] for the bearing of the pair-wise angle term (Kernel.cu:170-182): octant reduction, one
// approximate division, a degree-7 minimax polynomial of atan(t)/t in t^2 on [0, 1] (3.8e-8 before rounding), about
// 20 instructions where atan2f takes about 45 -- and the plain scan evaluates it R times per proposal.  Absolute
// error <= 4e-7 (1.5 ulp of pi; checked against a float64 atan2 over 4e6 points, tools/bearing_check.py), which is
// what matters here: the angle only ever enters differences against angleMin / angleMax.  atan2(0, 0) = 0 as in libm.
__device__ __forceinline__ float bearing_atan2(const float y, const float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    const float s = t * t;
    float p = -0.004054440185427666f;
    p = fmaf(p, s, 0.021862473338842392f);
    p = fmaf(p, s, -0.0559115894138813f);
    p = fmaf(p, s, 0.09642140567302704f);
    p = fmaf(p, s, -0.1390860676765442f);
    p = fmaf(p, s, 0.19946561753749847f);
    p = fmaf(p, s, -0.33329859375953674f);
    p = fmaf(p, s, 0.9999993443489075f);
    float a = t * p;
    if (ay > ax) a = 1.57079632679489662f - a;
    if (x < 0.f) a = 3.14159265358979324f - a;
    return copysignf(a, y);
}

// Penalties of relationship r (Kernel.cu:210-263) as positive magnitudes: pd = distance penalty of
// rss[r]'s pair, pa = angle penalty of rsa[r]'s pair.  Pc = this chain's float4 state, stride LD.
// (not inlined: it is called from several places and its ~110 instructions would otherwise be
// replicated, and the delta/memo kernels are bound by instruction fetch, not by issue; arguments and
// result travel in registers)
template <int LD>
__device__ __noinline__ float2 rel_pen_impl(const int4 *rel_idx, const float4 *rel_rng, const float4 *rel_aux, const float two_pi,
                                            const float4 *Pc, const int r)
{
    float pd, pa;
    const int4 id = rel_idx[r];       // distance pair (x, y) from rss[r], angle pair (z, w) from rsa[r]
    const float4 rg = rel_rng[r];     // 1/start, end, angleMin, angleMax
    const float4 ax = rel_aux[r];     // start, 1/norm, wraps
    MH_CHECK(r >= 0 && id.x >= 0 && id.y >= 0 && id.z >= 0 && id.w >= 0);
    const float4 ps = Pc[id.x * LD], pt = Pc[id.y * LD];
    pd = 0.f;
    pa = 0.f;
    {
        const float dX = ps.x - pt.x, dY = ps.y - pt.y;
        const float d2 = fmaf(dX, dX, dY * dY);
        const float d = sqrt_approx(d2);
        if (d < ax.x) {                                         // too close (Kernel.cu:219-223)
            const float f = d * rg.x;
            pd = f * f;
        } else if (d > rg.y) {                                  // too far (Kernel.cu:225-229)
            const float f = rg.y * rsqrt_approx(d2);
            pd = f * f;
        }
    }
    const float4 as = (id.z == id.x) ? ps : Pc[id.z * LD], at = (id.w == id.y) ? pt : Pc[id.w * LD];
    const float dX = as.x - at.x, dY = as.y - at.y;
    // bearing of source seen from target, relative to the target's rotation (Kernel.cu:170-182)
    float tp = bearing_atan2(dY, dX);
    if (tp < 0.f) tp = two_pi + tp;
    float th = tp - at.z;
    if (th < 0.f) th = two_pi + th;
    const float pen = fminf(fabsf(th - rg.z), fabsf(th - rg.w)) * ax.y;
    if (ax.z != 0.f) {                                          // range crosses zero (Kernel.cu:245-250)
        float f = rg.z + th;
        if (f >= two_pi) {                                      // fmodf (Q17); one subtraction covers [0, 4 PI)
            f -= two_pi;
            if (f >= two_pi) f = fmodf(f, two_pi);
        } else if (f < 0.f) {
            f = fmodf(f, two_pi);
        }
        if (f > rg.w) pa = pen;
    } else if (rg.z < th || th < rg.w) {                        // Q9: almost always true
        pa = pen;
    }
    return make_float2(pd, pa);
}

template <int LD>
__device__ __forceinline__ void rel_pen(const SmemProblem &P, const float4 *Pc, const int r, float &pd, float &pa)
{
    const float2 v = rel_pen_impl<LD>(P.rel_idx, P.rel_rng, P.rel_aux, P.h->two_pi, Pc, r);
    pd = v.x;
    pa = v.y;
}

// All terms of one layout.  Every lane of the warp must call this (it synchronises the warp);
// on return every lane of a group holds the group's totals.
template <int G, bool WITH_OFFLIMITS, bool STR = false, bool SKIP_SYM = false, bool SKIP_REL = false, bool SKIP_CLR = false>
__device__ __forceinline__ void eval_terms(const SmemProblem &P, const WarpState<G, !STR> &S, const int c, const int g, RawTerms &t)
{
    using WS = WarpState<G, !STR>;                              // contiguous groups (the scan kernels): padded rows
    constexpr int LD = WS::LD;
    const mhProblemHeader *h = P.h;
    const int n = h->n, C = h->C, R = h->R;
    const int Cc = SKIP_CLR ? 0 : C;                            // clearance columns walked by the row loops
    float surf = 0.f, sym = 0.f, vbx = 0.f, vby = 0.f, focal = 0.f, off = 0.f, pw = 0.f, pa = 0.f;
    long long clr = 0;                                          // fixed point: see "the clearance term is an integer sum"

    // ---- clearance rectangles: at their source object for ClearanceCosts, at object k for
    //      SurfaceAreaCosts (Q7) -------------------------------------------------------------
    for (int k = g; k < C; k += G) {
        const float4 kb = P.clr_box[k];
        const float v0 = P.clr_v0x[k];
        const float2 ps = *reinterpret_cast<const float2 *>(&S.P4[WS::at(P.clr_src[k], c)]);
        const float2 pk = *reinterpret_cast<const float2 *>(&S.P4[WS::at(k, c)]);
        if (!SKIP_CLR) S.CB[WS::at(k, c)] = box_at_q(P.clr_boxq[k], P.clr_v0xq[k], to_fixed(h, ps.x), to_fixed(h, ps.y));
        surf += outside_room(box_at(kb, v0, pk.x, pk.y), h);
    }
    __syncwarp();

    // ---- rows: two objects per lane per pass (i and i+G), so that every clearance rectangle and
    //      every column of the symmetry scan is loaded once for two rows; a last odd row runs the
    //      one-row form of the same code ---------------------------------------------------------------
    const float pi_f = 0.5f * h->two_pi;
    const float4 *Pc = S.P4 + c;
    const int4 *CBc = S.CB + c;
    int i = g;
#ifndef MH_NO_ROW_BLOCKING
    for (; !SKIP_SYM && i + G < n; i += 2 * G) {      // (the memo form has no column scan to share: one-row code only)
        const int i2 = i + G;
        const float4 p1 = Pc[i * LD], p2 = Pc[i2 * LD];
        const float ar1 = P.obj_area[i], ar2 = P.obj_area[i2];
        vbx = fmaf(ar1, p1.x, vbx);
        vby = fmaf(ar1, p1.y, vby);
        vbx = fmaf(ar2, p2.x, vbx);
        vby = fmaf(ar2, p2.y, vby);
        focal += p1.w;
        focal += p2.w;
        const float4 a1 = box_at(P.obj_box[i], P.obj_v0x[i], p1.x, p1.y);
        const float4 a2 = box_at(P.obj_box[i2], P.obj_v0x[i2], p2.x, p2.y);
        surf += outside_room(a1, h);
        surf += outside_room(a2, h);
        {
            // The symmetry scan is bound by the MUFU pipe and the clearance overlaps by the ALU pipe
            // (FMNMX), so the two loops are FUSED: each trip takes two symmetry columns and one
            // clearance rectangle for both rows, and the scheduler interleaves the two instruction
            // mixes inside one warp instead of relying on other warps being in the other phase.
            const RowRef r1 = sym_row(h, p1), r2 = sym_row(h, p2);
            const int4 a1q = box_at_q(P.obj_boxq[i], P.obj_v0xq[i], to_fixed(h, p1.x), to_fixed(h, p1.y));
            const int4 a2q = box_at_q(P.obj_boxq[i2], P.obj_v0xq[i2], to_fixed(h, p2.x), to_fixed(h, p2.y));
            float k1 = 5.0f, k2 = 5.0f;
            long long acc1 = 0, acc2 = 0;
            int j = 0, k = 0;
#ifndef MH_NO_LOOP_FUSION
            if (!SKIP_SYM) {
#pragma unroll kFuseUnroll
                for (; j + 2 <= n && k < Cc; j += 2, k++) {
                    const float4 q0 = Pc[j * LD], q1 = Pc[(j + 1) * LD];
                    const int4 b0 = CBc[k * LD];
                    k1 = fminf(k1, sym_key(r1, q0, pi_f));
                    k2 = fminf(k2, sym_key(r2, q0, pi_f));
                    acc1 = overlap_add_q(a1q, b0, acc1);
                    k1 = fminf(k1, sym_key(r1, q1, pi_f));
                    k2 = fminf(k2, sym_key(r2, q1, pi_f));
                    acc2 = overlap_add_q(a2q, b0, acc2);
                }
            }
#endif
#pragma unroll 2
            for (; k + 2 <= Cc; k += 2) {
                const int4 b0 = CBc[(k + 0) * LD], b1 = CBc[(k + 1) * LD];
                acc1 = overlap_add_q(a1q, b0, acc1);
                acc2 = overlap_add_q(a2q, b0, acc2);
                acc1 = overlap_add_q(a1q, b1, acc1);
                acc2 = overlap_add_q(a2q, b1, acc2);
            }
            for (; k < Cc; k++) {
                const int4 b0 = CBc[k * LD];
                acc1 = overlap_add_q(a1q, b0, acc1);
                acc2 = overlap_add_q(a2q, b0, acc2);
            }
            clr += acc1 + acc2;
            if (!SKIP_SYM) {
#pragma unroll 4
                for (; j < n; j++) {
                    const float4 q = Pc[j * LD];
                    k1 = fminf(k1, sym_key(r1, q, pi_f));
                    k2 = fminf(k2, sym_key(r2, q, pi_f));
                }
                sym += 5.0f - k1;
                sym += 5.0f - k2;
            }
        }
        if (WITH_OFFLIMITS) {                                   // Kernel.cu:488-511, pairs i < j
            float acc = 0.f;
            for (int j = i + 1; j < n; j++) {
                const float4 q = Pc[j * LD];
                acc += overlap(a1, box_at(P.obj_box[j], P.obj_v0x[j], q.x, q.y));
            }
            off += acc;
            acc = 0.f;
            for (int j = i2 + 1; j < n; j++) {
                const float4 q = Pc[j * LD];
                acc += overlap(a2, box_at(P.obj_box[j], P.obj_v0x[j], q.x, q.y));
            }
            off += acc;
        }
    }
#endif
    for (; i < n; i += G) {
        const float4 pi = Pc[i * LD];
        // visual balance partial sums (Kernel.cu:199-202); memoised focal cosine
        const float area = P.obj_area[i];
        vbx = fmaf(area, pi.x, vbx);
        vby = fmaf(area, pi.y, vby);
        focal += pi.w;
        // own off-limit rectangle: outside the room, against every clearance (Kernel.cu:404-434)
        const float4 a = box_at(P.obj_box[i], P.obj_v0x[i], pi.x, pi.y);
        surf += outside_room(a, h);
        if (!SKIP_CLR) clr += clearance_row_q<LD>(box_at_q(P.obj_boxq[i], P.obj_v0xq[i], to_fixed(h, pi.x), to_fixed(h, pi.y)), CBc, C);
        // symmetry: best match of the reflection of object i over all columns (see sym_key)
        if (!SKIP_SYM) {
            const RowRef rr = sym_row(h, pi);
            float kmin = 5.0f;
#pragma unroll kSymUnroll
            for (int j = 0; j < n; j++)
                kmin = fminf(kmin, sym_key(rr, Pc[j * LD], pi_f));
            sym += 5.0f - kmin;
        }
        if (WITH_OFFLIMITS) {                                   // Kernel.cu:488-511, pairs i < j
            float acc = 0.f;
            for (int j = i + 1; j < n; j++) {
                const float4 q = Pc[j * LD];
                acc += overlap(a, box_at(P.obj_box[j], P.obj_v0x[j], q.x, q.y));
            }
            off += acc;
        }
    }

    // ---- relationships --------------------------------------------------------------------------
    if (!SKIP_REL) {
        for (int r = g; r < R; r += G) {
            float pd, pe;
            rel_pen<LD>(P, Pc, r, pd, pe);
            pw += pd;
            pa += pe;
        }
    }

    t.pw = SKIP_REL ? 0.f : group_sum<G, STR>(pw);
    t.pa = SKIP_REL ? 0.f : group_sum<G, STR>(pa);
    t.vbx = group_sum<G, STR>(vbx);
    t.vby = group_sum<G, STR>(vby);
    t.focal = group_sum<G, STR>(focal);
    t.sym = group_sum<G, STR>(sym);
    t.clr_q = SKIP_CLR ? 0ll : group_sum_q<G, STR>(clr);
    t.clr = clearance_value(h, t.clr_q);
    t.surf = group_sum<G, STR>(surf);
    t.off = WITH_OFFLIMITS ? group_sum<G, STR>(off) : 0.f;
}

// Costs() proper (Kernel.cu:516-550): each raw term is a float, weighted in float, and the
// total is the left-to-right float sum of six of them (Q5).  __fmul_rn/__fadd_rn keep the
// compiler from fusing a weight multiply into the running sum, which the reference cannot do
// because it stores every weighted term first.
__device__ __forceinline__ Costs8 combine(const mhProblemHeader *h, const RawTerms &t)
{
    Costs8 c;
    c.pair = __fmul_rn(h->w_pair, __fmul_rn(t.pw, t.pa)); // (-pw)*(-pa): Q20
#if MH_RECIP_DENOM
    const float ax = t.vbx * h->inv_denom, ay = t.vby * h->inv_denom;   // (1 / sum of areas is a constant of the room)
#else
    const float ax = t.vbx / h->denom, ay = t.vby / h->denom;
#endif
    const float dX = ax - h->cx2, dY = ay - h->cy2;       // Q11
    c.visual = __fmul_rn(h->w_visual, -sqrtf(fmaf(dX, dX, dY * dY)));
    c.focal = __fmul_rn(h->w_focal, -t.focal);
    c.sym = __fmul_rn(h->w_sym, -t.sym);
    c.off = __fmul_rn(h->w_off, -t.off);
    c.clr = __fmul_rn(h->w_clear, -t.clr);
    c.surf = __fmul_rn(h->w_surf, -t.surf);
    float tot = __fadd_rn(c.pair, c.visual);
    tot = __fadd_rn(tot, c.focal);
    tot = __fadd_rn(tot, c.sym);
    tot = __fadd_rn(tot, c.clr);
    tot = __fadd_rn(tot, c.surf);
    c.total = tot;
    return c;
}

} // namespace mh
