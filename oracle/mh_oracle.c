/*
 * mh_oracle.c -- CPU restatement of the reference's per-chain Metropolis-Hastings path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (libKernel.so, its host code, the
 * Python binding) may include, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the reported CPU baseline.
 *
 * What it restates (K.cu = /root/reference/KernelFolder/Kernel/Kernel.cu):
 *   - the cost function Costs() and its seven terms, K.cu:162-550, with every precision
 *     quirk of the mixed float/double C++ source kept (see the comment on each function);
 *   - propose() / Accept(), K.cu:566-713, and the chain loop K.cu:777-827 as the serial
 *     chain "Semantics S" of SURVEY.md section 8a (one logical proposal per iteration);
 *   - the output narrowing K.cu:834-842.
 * What it does NOT restate: the cuRAND XORWOW bit stream (third-party, toolkit header
 * curand_kernel.h, pinned "CUDA 8.0" by Kernel.vcxproj:44; the reference seeds it with
 * time(NULL), K.cu:943, so no reference output pins it).  Random numbers come from the
 * counter-based Philox4x32-10 stream specified below, which the CUDA kernel shares; the
 * uniform and Box-Muller transforms follow cuRAND's published ones (curand_uniform.h:69-72,
 * curand_normal.h:70-92).
 *
 * Parity pin: the cost functions are checked bit-for-bit on this host against the
 * reference's own source compiled as host C++ (oracle/_ref/libref_costs_host.so, built by
 * oracle/Makefile from the reference tree in place) -- tests/test_oracle.py::
 * test_bit_exact_against_reference_host_build -- and
 * against the committed golden vectors that library produced (tests/golden/).  The RNG is
 * pinned to the Random123 known-answer vectors for Philox4x32-10.  The proposal stream as
 * a whole is "parity unpinned" against the reference by construction (wall-clock seed); what
 * propose() and Accept() DO with their draws is pinned distributionally against the reference's
 * own device functions (oracle/ref_gpu_harness.cu: RefProposeGPU / RefAcceptGPU;
 * tests/test_propose_parity.py).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (oracle/Makefile).  -ffp-contract=off is REQUIRED so
 * that no a*b+c is fused: the reference's host build does not fuse either.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/mh_layout.h"

#define PI MH_PI     /* 3.1416, K.cu:31 (quirk Q4) */
#define BETA MH_BETA /* K.cu:33 */
#define S_SIGMA_T MH_S_SIGMA_T /* K.cu:39 */

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11; same constants as curand_philox4x32_x.h:88-91).
 * ------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

ORACLE_API void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream spec shared with the CUDA kernel:
 *   key     = (seed lo, seed hi)
 *   counter = (iteration lo, draw block + (iteration hi << 16), chain lo, chain hi)
 *   block 0 : w0 -> move type, w1 -> first object, w2/w3 -> Box-Muller pair (translate,
 *             rotate) or w2 -> second object (swap)
 *   block 1 : w0 -> acceptance uniform
 *   block 2+t: t-th re-draw while the picked object is frozen: w0 -> first object,
 *             w1 -> second object (K.cu:601, 637, 662, 666)                                  */
static void draw_block(uint64_t seed, uint64_t chain, uint64_t it, uint32_t block, uint32_t w[4])
{
    uint32_t ctr[4] = { (uint32_t)it, block + ((uint32_t)(it >> 32) << 16), (uint32_t)chain,
                        (uint32_t)(chain >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    oracle_philox4x32_10(ctr, key, w);
}

/* curand_uniform.h:69-72: (0, 1].  The multiply is by 2^-32, hence exact, hence the result
 * is the same with or without FMA contraction. */
static float uniform_from_bits(uint32_t x) { return (float)x * 0x1p-32f + 0x1p-33f; }

/* curand_normal.h:70-92 (_curand_box_muller): first normal = s*sin(v), second = s*cos(v).
 * v is written as an explicit fmaf so host and device agree on it bit for bit. */
static void box_muller(uint32_t x, uint32_t y, float *n0, float *n1)
{
    const float two_pi_2pow32_inv = 1.46291807e-09f;
    float u = uniform_from_bits(x);
    float v = fmaf((float)y, two_pi_2pow32_inv, two_pi_2pow32_inv / 2.0f);
    float s = sqrtf(-2.0f * logf(u));
    *n0 = s * sinf(v);
    *n1 = s * cosf(v);
}

ORACLE_API float oracle_uniform(uint32_t x) { return uniform_from_bits(x); }
ORACLE_API void oracle_box_muller(uint32_t x, uint32_t y, float *n0, float *n1) { box_muller(x, y, n0, n1); }

/* K.cu:566-574, with u supplied.  p_rand is a float; the scale is a double; the sum with
 * `min` is float.  Q13 (index n when u == 1 and n-1 >= 32) is NOT kept: clamp. */
static int random_int_in_range(float u, int max, int min)
{
    float p_rand = u;
    p_rand = (float)((double)p_rand * (max - min + 0.999999));
    p_rand = p_rand + (float)min;
    int v = (int)truncf(p_rand);
    return v > max ? max : v;
}
ORACLE_API int oracle_random_int(float u, int max, int min) { return random_int_in_range(u, max, min); }

/* ------------------------------------------------------------------------------------------
 * Cost terms.  Each mirrors the reference expression by expression, including which
 * sub-expressions are float and which are double in the C++ source.
 * ------------------------------------------------------------------------------------------ */

/* K.cu:162-167.  Arguments are float (callers' doubles are narrowed at the call). */
static double Distance(float xi, float yi, float xj, float yj)
{
    double dX = xi - xj; /* float subtraction, then widened */
    double dY = yi - yj;
    return sqrt(dX * dX + dY * dY);
}

/* K.cu:170-182.  Bearing in [0, 2*PI) with PI = 3.1416; atan2 in double. */
static double theta(float xi, float yi, float xj, float yj, float ti)
{
    double dX = xi - xj;
    double dY = yi - yj;
    double theta_p = atan2(dY, dX);
    theta_p = (theta_p < 0) ? 2 * PI + theta_p : theta_p;
    double th = theta_p - ti;
    return (th < 0) ? 2 * PI + th : th;
}

/* K.cu:185-188.  atan2 of float arguments resolves to the float overload in C++ (and in
 * CUDA device code); the subtraction of tj is float; adding PI/2.0 promotes to double; the
 * return narrows to float. */
static float phi(float xi, float yi, float xj, float yj, float tj)
{
    return (float)((double)(atan2f(yi - yj, xi - xj) - tj) + PI / 2.0);
}

/* K.cu:191-207.  float accumulators, each step computed through a double product (Q11:
 * the centroid is halved). */
static double VisualBalanceCosts(const Surface *srf, const positionAndRotation *cfg)
{
    float nx = 0, ny = 0, denom = 0;
    for (int i = 0; i < srf->nObjs; i++) {
        float area = (float)(cfg[i].length * cfg[i].width);
        nx = (float)((double)nx + (double)area * cfg[i].x);
        ny = (float)((double)ny + (double)area * cfg[i].y);
        denom += area;
    }
    return -1.0 * Distance(nx / denom, ny / denom, (float)(srf->centroidX / 2), (float)(srf->centroidY / 2));
}

/* K.cu:210-233 */
static double PairWiseCosts(const Surface *srf, const positionAndRotation *cfg, const relationshipStruct *rs)
{
    double result = 0;
    for (int i = 0; i < srf->nRelationships; i++) {
        const positionAndRotation *s = &cfg[rs[i].SourceIndex], *t = &cfg[rs[i].TargetIndex];
        double distance = Distance((float)s->x, (float)s->y, (float)t->x, (float)t->y);
        if (distance < rs[i].TargetRange.targetRangeStart) {
            double fraction = distance / rs[i].TargetRange.targetRangeStart;
            result -= (fraction * fraction);
        } else if (distance > rs[i].TargetRange.targetRangeEnd) {
            double fraction = rs[i].TargetRange.targetRangeEnd / distance;
            result -= (fraction * fraction);
        }
    }
    return result;
}

/* K.cu:236-263.  Q17: fmodf on narrowed operands.  Q9: the second condition is almost
 * always true. */
static double PairWiseAngleCosts(const Surface *srf, const positionAndRotation *cfg, const relationshipAngleStruct *rs)
{
    double result = 0;
    for (int i = 0; i < srf->nRelationships; i++) {
        const positionAndRotation *s = &cfg[rs[i].SourceIndex], *t = &cfg[rs[i].TargetIndex];
        double distance = theta((float)s->x, (float)s->y, (float)t->x, (float)t->y, (float)t->rotY);
        if (rs[i].angleMin > rs[i].angleMax) {
            double norm = (2 * PI - (rs[i].angleMax + (2 * PI - rs[i].angleMin))) / 2.0;
            if (fmodf((float)(rs[i].angleMin + distance), (float)(2 * PI)) > rs[i].angleMax)
                result -= fmin(fabs(distance - rs[i].angleMin), fabs(distance - rs[i].angleMax)) / norm;
        } else if (rs[i].angleMin < distance || distance < rs[i].angleMax) {
            double norm = (2 * PI - (rs[i].angleMax - rs[i].angleMin)) / 2.0;
            result -= fmin(fabs(distance - rs[i].angleMin), fabs(distance - rs[i].angleMax)) / norm;
        }
    }
    return result;
}

/* K.cu:266-281.  cos of a float argument is the float overload; the sum is double. */
static double FocalPointCosts(const Surface *srf, const positionAndRotation *cfg)
{
    double sum = 0;
    for (int i = 0; i < srf->nObjs; i++) {
        float phi_fi = phi((float)srf->focalX, (float)srf->focalY, (float)cfg[i].x, (float)cfg[i].y, (float)cfg[i].rotY);
        sum -= (double)cosf(phi_fi);
    }
    return sum;
}

/* K.cu:283-318.  Q18: one-sided angle wraps.  gamma_ij = 1 and j runs over ALL objects
 * including i. */
static float SymmetryCosts(const Surface *srf, const positionAndRotation *cfg)
{
    float sum = 0;
    for (int i = 0; i < srf->nObjs; i++) {
        float maxVal = 0;
        float ux = (float)cos(srf->focalRot);
        float uy = (float)sin(srf->focalRot);
        float s = (float)(2 * (srf->focalX * ux + srf->focalY * uy - (cfg[i].x * ux + cfg[i].y * uy)));
        float rx_i = (float)(cfg[i].x + (double)(s * ux));
        float ry_i = (float)(cfg[i].y + (double)(s * uy));
        float rRot_i = (float)(2 * srf->focalRot - cfg[i].rotY);
        if (rRot_i < -PI)
            rRot_i = (float)(rRot_i + 2 * PI);
        for (int j = 0; j < srf->nObjs; j++) {
            int gamma_ij = 1;
            float dp = (float)Distance((float)cfg[j].x, (float)cfg[j].y, rx_i, ry_i);
            float dt = (float)(cfg[j].rotY - rRot_i);
            if (dt > PI)
                dt = (float)(dt - 2 * PI);
            /* 5 - sqrt(dp) is int - float -> float; 0.4 * fabs(dt) is double. */
            float val = (float)(gamma_ij * ((double)(5 - sqrtf(dp)) - 0.4 * (double)fabsf(dt)));
            maxVal = fmaxf(maxVal, val);
        }
        sum -= maxVal;
    }
    return sum;
}

/* K.cu:321-340.  fmaxf/fminf narrow the double coordinates to float (+-DBL_MAX -> +-inf). */
static float calculateIntersectionArea(vertex rect1Min, vertex rect1Max, vertex rect2Min, vertex rect2Max)
{
    float x5 = fmaxf((float)rect1Min.x, (float)rect2Min.x);
    float y5 = fmaxf((float)rect1Min.y, (float)rect2Min.y);
    float x6 = fminf((float)rect1Max.x, (float)rect2Max.x);
    float y6 = fminf((float)rect1Max.y, (float)rect2Max.y);
    if (x5 >= x6 || y5 >= y6)
        return 0.0f;
    return (x6 - x5) * (y6 - y5);
}

/* K.cu:343-364.  Index 0 = min corner, 1 = max corner. */
static void createComplementRectangle(vertex srfRectMin, vertex srfRectMax, vertex *c1, vertex *c2, vertex *c3, vertex *c4)
{
    c1[0].x = -DBL_MAX; c1[0].y = -DBL_MAX;    c1[1].x = DBL_MAX;      c1[1].y = srfRectMin.y;
    c2[0].x = -DBL_MAX; c2[0].y = srfRectMin.y; c2[1].x = srfRectMin.x; c2[1].y = srfRectMax.y;
    c3[0].x = -DBL_MAX; c3[0].y = srfRectMax.y; c3[1].x = DBL_MAX;      c3[1].y = DBL_MAX;
    c4[0].x = srfRectMax.x; c4[0].y = srfRectMin.y; c4[1].x = DBL_MAX;  c4[1].y = srfRectMax.y;
}

/* K.cu:366-382.  Q6: the first vertex's x is taken WITHOUT the translation.  Q8: the four
 * consecutive vertices from startIndexVertices, translated only, never rotated. */
static vertex minValue(const vertex *vertices, int start, float xt, float yt)
{
    vertex r;
    r.x = DBL_MAX; r.y = DBL_MAX; r.z = 0;
    r.x = (r.x > vertices[start].x + xt) ? vertices[start].x : r.x;
    r.x = (r.x > vertices[start + 1].x + xt) ? vertices[start + 1].x + xt : r.x;
    r.x = (r.x > vertices[start + 2].x + xt) ? vertices[start + 2].x + xt : r.x;
    r.x = (r.x > vertices[start + 3].x + xt) ? vertices[start + 3].x + xt : r.x;
    r.y = (r.y > vertices[start].y + yt) ? vertices[start].y + yt : r.y;
    r.y = (r.y > vertices[start + 1].y + yt) ? vertices[start + 1].y + yt : r.y;
    r.y = (r.y > vertices[start + 2].y + yt) ? vertices[start + 2].y + yt : r.y;
    r.y = (r.y > vertices[start + 3].y + yt) ? vertices[start + 3].y + yt : r.y;
    return r;
}

/* K.cu:384-401 */
static vertex maxValue(const vertex *vertices, int start, float xt, float yt)
{
    vertex r;
    r.x = -DBL_MAX; r.y = -DBL_MAX; r.z = 0;
    for (int k = 0; k < 4; k++)
        r.x = (r.x < vertices[start + k].x + xt) ? vertices[start + k].x + xt : r.x;
    for (int k = 0; k < 4; k++)
        r.y = (r.y < vertices[start + k].y + yt) ? vertices[start + k].y + yt : r.y;
    return r;
}

/* K.cu:404-434.  Clearance i (translated by its SourceIndex object) against every
 * off-limit rectangle j, including its own object's. */
static float ClearanceCosts(const Surface *srf, const positionAndRotation *cfg, const vertex *vertices,
                            const rectangle *clearances, const rectangle *offlimits)
{
    float error = 0.0f;
    for (int i = 0; i < srf->nClearances; i++) {
        int src = clearances[i].SourceIndex;
        vertex r1min = minValue(vertices, clearances[i].point1Index, (float)cfg[src].x, (float)cfg[src].y);
        vertex r1max = maxValue(vertices, clearances[i].point1Index, (float)cfg[src].x, (float)cfg[src].y);
        for (int j = 0; j < srf->nObjs; j++) {
            vertex r2min = minValue(vertices, offlimits[j].point1Index, (float)cfg[j].x, (float)cfg[j].y);
            vertex r2max = maxValue(vertices, offlimits[j].point1Index, (float)cfg[j].x, (float)cfg[j].y);
            error -= calculateIntersectionArea(r1min, r1max, r2min, r2max);
        }
    }
    return error;
}

/* K.cu:437-483.  Q7: clearance i is translated by cfg[i], not by its source object. */
static float SurfaceAreaCosts(const Surface *srf, const positionAndRotation *cfg, const vertex *vertices,
                              const rectangle *clearances, const rectangle *offlimits, const vertex *surfaceRectangle)
{
    float error = 0.0f;
    vertex c1[2], c2[2], c3[2], c4[2];
    vertex smin = minValue(surfaceRectangle, 0, 0, 0);
    vertex smax = maxValue(surfaceRectangle, 0, 0, 0);
    createComplementRectangle(smin, smax, c1, c2, c3, c4);
    for (int i = 0; i < srf->nClearances; i++) {
        vertex rmin = minValue(vertices, clearances[i].point1Index, (float)cfg[i].x, (float)cfg[i].y);
        vertex rmax = maxValue(vertices, clearances[i].point1Index, (float)cfg[i].x, (float)cfg[i].y);
        error -= calculateIntersectionArea(rmin, rmax, c1[0], c1[1]);
        error -= calculateIntersectionArea(rmin, rmax, c2[0], c2[1]);
        error -= calculateIntersectionArea(rmin, rmax, c3[0], c3[1]);
        error -= calculateIntersectionArea(rmin, rmax, c4[0], c4[1]);
    }
    for (int j = 0; j < srf->nObjs; j++) {
        vertex rmin = minValue(vertices, offlimits[j].point1Index, (float)cfg[j].x, (float)cfg[j].y);
        vertex rmax = maxValue(vertices, offlimits[j].point1Index, (float)cfg[j].x, (float)cfg[j].y);
        error -= calculateIntersectionArea(rmin, rmax, c1[0], c1[1]);
        error -= calculateIntersectionArea(rmin, rmax, c2[0], c2[1]);
        error -= calculateIntersectionArea(rmin, rmax, c3[0], c3[1]);
        error -= calculateIntersectionArea(rmin, rmax, c4[0], c4[1]);
    }
    return error;
}

/* K.cu:485-514 */
static float OffLimitsCosts(const Surface *srf, const positionAndRotation *cfg, const vertex *vertices, const rectangle *offlimits)
{
    float error = 0.0f;
    for (int i = 0; i < srf->nObjs; i++) {
        vertex r1min = minValue(vertices, offlimits[i].point1Index, (float)cfg[i].x, (float)cfg[i].y);
        vertex r1max = maxValue(vertices, offlimits[i].point1Index, (float)cfg[i].x, (float)cfg[i].y);
        for (int j = i + 1; j < srf->nObjs; j++) {
            vertex r2min = minValue(vertices, offlimits[j].point1Index, (float)cfg[j].x, (float)cfg[j].y);
            vertex r2max = maxValue(vertices, offlimits[j].point1Index, (float)cfg[j].x, (float)cfg[j].y);
            error -= calculateIntersectionArea(r1min, r1max, r2min, r2max);
        }
    }
    return error;
}

/* Unweighted terms, for diagnostics: [0] pair-wise distance sum, [1] pair-wise angle sum,
 * [2] visual balance, [3] focal point, [4] symmetry, [5] off-limits, [6] clearance,
 * [7] surface area -- each already narrowed the way Costs() narrows it, except [0],[1]. */
typedef struct oracleRawTerms {
    double v[8];
} oracleRawTerms;

/* K.cu:516-550.  Q20: pair-wise = distance sum x angle sum.  Q5: the total leaves the
 * off-limits term out.  Every term is narrowed to float before it is weighted. */
static void Costs(const Surface *srf, resultCosts *costs, const positionAndRotation *cfg, const relationshipStruct *rs,
                  const relationshipAngleStruct *ra, const vertex *vertices, const rectangle *clearances,
                  const rectangle *offlimits, const vertex *surfaceRectangle, oracleRawTerms *raw, int with_offlimits)
{
    double pw = PairWiseCosts(srf, cfg, rs);
    double pa = PairWiseAngleCosts(srf, cfg, ra);
    float pairWiseCosts = (float)(pw * pa);
    costs->PairWiseCosts = srf->WeightPairWise * pairWiseCosts;

    float visualBalanceCosts = (float)VisualBalanceCosts(srf, cfg);
    costs->VisualBalanceCosts = srf->WeightVisualBalance * visualBalanceCosts;

    float focalPointCosts = (float)FocalPointCosts(srf, cfg);
    costs->FocalPointCosts = srf->WeightFocalPoint * focalPointCosts;

    float symmertryCosts = SymmetryCosts(srf, cfg);
    costs->SymmetryCosts = srf->WeightSymmetry * symmertryCosts;

    /* with_offlimits == 0 is used by the timed CPU baseline only when explicitly asked to
     * skip the term the total never reads; the default evaluates it like the reference. */
    float offlimitsCosts = with_offlimits ? OffLimitsCosts(srf, cfg, vertices, offlimits) : 0.0f;
    costs->OffLimitsCosts = srf->WeightOffLimits * offlimitsCosts;

    float clearanceCosts = ClearanceCosts(srf, cfg, vertices, clearances, offlimits);
    costs->ClearanceCosts = srf->WeightClearance * clearanceCosts;

    float surfaceAreaCosts = SurfaceAreaCosts(srf, cfg, vertices, clearances, offlimits, surfaceRectangle);
    costs->SurfaceAreaCosts = srf->WeightSurfaceArea * surfaceAreaCosts;

    float totalCosts = costs->PairWiseCosts + costs->VisualBalanceCosts + costs->FocalPointCosts + costs->SymmetryCosts +
                       costs->ClearanceCosts + costs->SurfaceAreaCosts;
    costs->totalCosts = totalCosts;
    if (raw) {
        raw->v[0] = pw; raw->v[1] = pa; raw->v[2] = visualBalanceCosts; raw->v[3] = focalPointCosts;
        raw->v[4] = symmertryCosts; raw->v[5] = offlimitsCosts; raw->v[6] = clearanceCosts; raw->v[7] = surfaceAreaCosts;
    }
}

ORACLE_API void oracle_costs(const Surface *srf, const positionAndRotation *cfg, const relationshipStruct *rs,
                             const relationshipAngleStruct *ra, const vertex *vertices, const rectangle *clearances,
                             const rectangle *offlimits, const vertex *surfaceRectangle, resultCosts *out,
                             double *raw8 /* may be NULL */)
{
    oracleRawTerms raw;
    Costs(srf, out, cfg, rs, ra, vertices, clearances, offlimits, surfaceRectangle, &raw, 1);
    if (raw8)
        memcpy(raw8, raw.v, sizeof raw.v);
}

/* Batch form: layouts[l*n + i], out[l]. */
ORACLE_API void oracle_costs_batch(const Surface *srf, const positionAndRotation *layouts, int nLayouts,
                                   const relationshipStruct *rs, const relationshipAngleStruct *ra, const vertex *vertices,
                                   const rectangle *clearances, const rectangle *offlimits, const vertex *surfaceRectangle,
                                   resultCosts *out)
{
#pragma omp parallel for schedule(static)
    for (int l = 0; l < nLayouts; l++)
        Costs(srf, &out[l], layouts + (size_t)l * srf->nObjs, rs, ra, vertices, clearances, offlimits, surfaceRectangle, NULL, 1);
}

/* Angle distance of every relationship and how close it sits to one of the branch
 * boundaries of K.cu:245-254, where the penalty jumps: a parity test must not compare a
 * float32 and a float64 evaluation on a layout that straddles such a jump.  Returns the
 * smallest margin over all relationships. */
ORACLE_API double oracle_angle_branch_margin(const Surface *srf, const positionAndRotation *cfg, const relationshipAngleStruct *rs)
{
    double margin = 1e30;
    for (int i = 0; i < srf->nRelationships; i++) {
        const positionAndRotation *s = &cfg[rs[i].SourceIndex], *t = &cfg[rs[i].TargetIndex];
        double dX = (float)s->x - (float)t->x, dY = (float)s->y - (float)t->y;
        double tp = atan2(dY, dX);
        double d0 = (tp < 0) ? 2 * PI + tp : tp;
        double d1 = d0 - (float)t->rotY;
        double d = (d1 < 0) ? 2 * PI + d1 : d1;
        double m = fmin(fabs(tp), fabs(d1)); /* the two sign wraps of theta() */
        if (rs[i].angleMin > rs[i].angleMax) {
            double f = fmod(rs[i].angleMin + d, 2 * PI);
            m = fmin(m, fabs(f - rs[i].angleMax));          /* K.cu:248 threshold       */
            m = fmin(m, fmin(f, 2 * PI - f));               /* the fmod wrap itself      */
        }
        margin = fmin(margin, m);
    }
    return margin;
}

/* ------------------------------------------------------------------------------------------
 * The chain (Semantics S).
 * ------------------------------------------------------------------------------------------ */
typedef struct oracleProblem {
    const Surface *srf;
    const relationshipStruct *rs;
    const relationshipAngleStruct *ra;
    const positionAndRotation *cfg;
    const rectangle *clearances;
    const rectangle *offlimits;
    const vertex *vertices;
    const vertex *surfaceRectangle;
} oracleProblem;

typedef struct oracleTraceEntry { /* same layout as mhTraceEntry */
    int32_t move, obj1, obj2, accepted;
    float star_total, cur_total, u, beta;
} oracleTraceEntry;

typedef struct oracleRunOptions {
    uint64_t seed;
    uint64_t chain_offset;
    uint64_t iteration_offset;
    double beta_start, beta_end;
    int32_t schedule;        /* 0 constant, 1 geometric, 2 linear (mh_kernel.h MH_SCHEDULE_*) */
    int32_t schedule_length; /* 0 -> iterations */
    int32_t result_mode;     /* 0 final, 1 best */
    int32_t with_offlimits;  /* 1 = evaluate the dead off-limits term every proposal like K.cu:534 */
    int32_t threads;         /* 0 -> omp default */
    int32_t tempering_rungs; /* 0 = off */
    int32_t exchange_interval;
    int32_t _pad;
} oracleRunOptions;

static double beta_at(const oracleRunOptions *o, uint64_t it, int iterations)
{
    double b0 = o->beta_start > 0 ? o->beta_start : BETA;
    double b1 = o->beta_end > 0 ? o->beta_end : b0;
    int len = o->schedule_length > 0 ? o->schedule_length : iterations;
    if (o->schedule == 0 || len <= 1)
        return b0;
    double t = (double)(it < (uint64_t)(len - 1) ? it : (uint64_t)(len - 1)) / (double)(len - 1);
    /* single precision like the kernel: beta is a float there */
    if (o->schedule == 1)
        return (double)((float)b0 * exp2f((float)t * log2f((float)(b1 / b0))));
    return (double)((float)b0 + (float)(b1 - b0) * (float)t);
}

static int any_free(const positionAndRotation *cfg, int n)
{
    for (int i = 0; i < n; i++)
        if (!cfg[i].frozen)
            return 1;
    return 0;
}

/* K.cu:576-704 with the Philox stream.  Returns the move type; obj1/obj2 = touched objects
 * (-1 when none).  Q14: with no free object nothing moves (the reference would spin). */
static int propose(const oracleProblem *P, positionAndRotation *cfgStar, uint64_t seed, uint64_t chain, uint64_t it,
                   int *obj1_out, int *obj2_out)
{
    const Surface *srf = P->srf;
    uint32_t w[4], rw[4];
    draw_block(seed, chain, it, 0, w);
    int p = random_int_in_range(uniform_from_bits(w[0]), 2, 0);
    *obj1_out = -1;
    *obj2_out = -1;

    vertex smin = minValue(P->surfaceRectangle, 0, 0, 0);
    vertex smax = maxValue(P->surfaceRectangle, 0, 0, 0);
    float width = (float)(smax.x - smin.x);
    float height = (float)(smax.y - smin.y);
    float stdXAxis = width / 16; /* Q19 */
    float stdYAxis = height / 16;
    int n = srf->nObjs;
    int movable = any_free(cfgStar, n);
    uint32_t redraw = 2;

    if (p == 0) {
        if (!movable) return p;
        int obj = random_int_in_range(uniform_from_bits(w[1]), n - 1, 0);
        while (cfgStar[obj].frozen) {
            draw_block(seed, chain, it, redraw++, rw);
            obj = random_int_in_range(uniform_from_bits(rw[0]), n - 1, 0);
        }
        float n0, n1;
        box_muller(w[2], w[3], &n0, &n1);
        float dx = n0 * stdXAxis;
        float dy = n1 * stdYAxis;
        if (cfgStar[obj].x + dx > smax.x) cfgStar[obj].x = smax.x;
        else if (cfgStar[obj].x + dx < smin.x) cfgStar[obj].x = smin.x;
        else cfgStar[obj].x += dx;
        if (cfgStar[obj].y + dy > smax.y) cfgStar[obj].y = smax.y;
        else if (cfgStar[obj].y + dy < smin.y) cfgStar[obj].y = smin.y;
        else cfgStar[obj].y += dy;
        *obj1_out = obj;
    } else if (p == 1) {
        if (!movable) return p;
        int obj = random_int_in_range(uniform_from_bits(w[1]), n - 1, 0);
        while (cfgStar[obj].frozen) {
            draw_block(seed, chain, it, redraw++, rw);
            obj = random_int_in_range(uniform_from_bits(rw[0]), n - 1, 0);
        }
        float n0, n1;
        box_muller(w[2], w[3], &n0, &n1);
        float dRot = n0;
        dRot = (float)(dRot * S_SIGMA_T);
        cfgStar[obj].rotY += dRot;
        if (cfgStar[obj].rotY < 0) cfgStar[obj].rotY += 2 * PI;
        else if (cfgStar[obj].rotY > 2 * PI) cfgStar[obj].rotY -= 2 * PI;
        *obj1_out = obj;
    } else {
        if (n < 2 || !movable) return p;
        int obj1 = random_int_in_range(uniform_from_bits(w[1]), n - 1, 0);
        int obj2 = random_int_in_range(uniform_from_bits(w[2]), n - 1, 0);
        while (cfgStar[obj1].frozen || cfgStar[obj2].frozen) {
            draw_block(seed, chain, it, redraw++, rw);
            if (cfgStar[obj1].frozen) obj1 = random_int_in_range(uniform_from_bits(rw[0]), n - 1, 0);
            if (cfgStar[obj2].frozen) obj2 = random_int_in_range(uniform_from_bits(rw[1]), n - 1, 0);
        }
        /* Q12: obj1's fields pass through float temporaries */
        float x = (float)cfgStar[obj1].x, y = (float)cfgStar[obj1].y, z = (float)cfgStar[obj1].z;
        float rotX = (float)cfgStar[obj1].rotX, rotY = (float)cfgStar[obj1].rotY, rotZ = (float)cfgStar[obj1].rotZ;
        cfgStar[obj1].x = cfgStar[obj2].x; cfgStar[obj1].y = cfgStar[obj2].y; cfgStar[obj1].z = cfgStar[obj2].z;
        cfgStar[obj1].rotX = cfgStar[obj2].rotX; cfgStar[obj1].rotY = cfgStar[obj2].rotY; cfgStar[obj1].rotZ = cfgStar[obj2].rotZ;
        cfgStar[obj2].x = x; cfgStar[obj2].y = y; cfgStar[obj2].z = z;
        cfgStar[obj2].rotX = rotX; cfgStar[obj2].rotY = rotY; cfgStar[obj2].rotZ = rotZ;
        *obj1_out = obj1;
        *obj2_out = obj2;
    }
    return p;
}

/* K.cu:706-713 with beta a parameter (reference: BETA = 2.0).  Q10: maximises totalCosts. */
static int Accept(double costStar, double costCur, float u, double beta)
{
    return u < fminf(1.0f, (float)exp(beta * (costStar - costCur)));
}

typedef struct chainState {
    positionAndRotation *cur, *star, *best;
    resultCosts curCosts, bestCosts;
    double beta; /* tempering: the rung's current beta */
} chainState;

static void chain_step(const oracleProblem *P, chainState *S, const oracleRunOptions *o, uint64_t chain, uint64_t it,
                       double beta, oracleTraceEntry *tr)
{
    int n = P->srf->nObjs;
    memcpy(S->star, S->cur, sizeof(positionAndRotation) * (size_t)n); /* K.cu:792 */
    int o1, o2;
    int p = propose(P, S->star, o->seed, chain, it, &o1, &o2); /* K.cu:798 */
    resultCosts starCosts;
    Costs(P->srf, &starCosts, S->star, P->rs, P->ra, P->vertices, P->clearances, P->offlimits, P->surfaceRectangle, NULL,
          o->with_offlimits); /* K.cu:804 */
    uint32_t w[4];
    draw_block(o->seed, chain, it, 1, w);
    float u = uniform_from_bits(w[0]);
    int acc = Accept(starCosts.totalCosts, S->curCosts.totalCosts, u, beta); /* K.cu:819 */
    if (acc) {
        memcpy(S->cur, S->star, sizeof(positionAndRotation) * (size_t)n); /* K.cu:824 */
        S->curCosts = starCosts;
        if (S->best && S->curCosts.totalCosts > S->bestCosts.totalCosts) {
            memcpy(S->best, S->cur, sizeof(positionAndRotation) * (size_t)n);
            S->bestCosts = S->curCosts;
        }
    }
    if (tr) {
        tr->move = p; tr->obj1 = o1; tr->obj2 = o2; tr->accepted = acc;
        tr->star_total = starCosts.totalCosts; tr->cur_total = S->curCosts.totalCosts; tr->u = u; tr->beta = (float)beta;
    }
}

static void emit(const oracleProblem *P, const positionAndRotation *cfg, const resultCosts *c, point *points, resultCosts *costs)
{
    int n = P->srf->nObjs;
    for (int i = 0; i < n; i++) { /* K.cu:834-842 */
        points[i].x = (float)cfg[i].x; points[i].y = (float)cfg[i].y; points[i].z = (float)cfg[i].z;
        points[i].rotX = (float)cfg[i].rotX; points[i].rotY = (float)cfg[i].rotY; points[i].rotZ = (float)cfg[i].rotZ;
    }
    if (costs) {
        /* Q3 is not kept: report the costs of the emitted layout, off-limits term included. */
        Costs(P->srf, costs, cfg, P->rs, P->ra, P->vertices, P->clearances, P->offlimits, P->surfaceRectangle, NULL, 1);
        (void)c;
    }
}

/* Run nChains chains (global ids chain_offset ..) for `iterations` steps each.
 * points[nChains*n], costs[nChains]; trace (may be NULL) [it*nChains + chain].
 * Without tempering chains are independent and run one per OpenMP thread.  Returns the
 * number of threads used. */
ORACLE_API int oracle_run(const Surface *srf, const relationshipStruct *rs, const relationshipAngleStruct *ra,
                          const positionAndRotation *cfg, const rectangle *clearances, const rectangle *offlimits,
                          const vertex *vertices, const vertex *surfaceRectangle, int nChains, int iterations,
                          const oracleRunOptions *opt, point *points, resultCosts *costs, oracleTraceEntry *trace)
{
    oracleProblem P = { srf, rs, ra, cfg, clearances, offlimits, vertices, surfaceRectangle };
    oracleRunOptions o = *opt;
    int n = srf->nObjs;
    int threads = 1;
#ifdef _OPENMP
    threads = o.threads > 0 ? o.threads : omp_get_max_threads();
    if (threads > nChains) threads = nChains > 0 ? nChains : 1;
#endif
    if (o.tempering_rungs > 1) {
        /* Parallel tempering, serial per ladder (extension; no reference counterpart).
         * Ladder l = chains [l*T, (l+1)*T) by GLOBAL id; rung r starts with beta_r.  Every
         * exchange_interval iterations, neighbouring chain pairs (r, r+1), r = epoch parity,
         * swap their betas with probability min(1, exp((b_r - b_{r+1}) * (E_r - E_{r+1})))
         * with E = -totalCosts; u from Philox block 0xFFFF of the pair's lower chain. */
        int T = o.tempering_rungs;
        int ex = o.exchange_interval > 0 ? o.exchange_interval : 100;
        if (o.chain_offset % (uint64_t)T != 0 || nChains % T != 0) return -1;
        double b0 = o.beta_start > 0 ? o.beta_start : BETA, b1 = o.beta_end > 0 ? o.beta_end : b0;
        int nl = nChains / T;
#pragma omp parallel for schedule(dynamic) num_threads(threads)
        for (int l = 0; l < nl; l++) {
            chainState *S = (chainState *)calloc((size_t)T, sizeof *S);
            for (int r = 0; r < T; r++) {
                S[r].cur = (positionAndRotation *)malloc(sizeof(positionAndRotation) * (size_t)n);
                S[r].star = (positionAndRotation *)malloc(sizeof(positionAndRotation) * (size_t)n);
                S[r].best = (positionAndRotation *)malloc(sizeof(positionAndRotation) * (size_t)n);
                memcpy(S[r].cur, cfg, sizeof(positionAndRotation) * (size_t)n);
                Costs(srf, &S[r].curCosts, S[r].cur, rs, ra, vertices, clearances, offlimits, surfaceRectangle, NULL, 1);
                memcpy(S[r].best, S[r].cur, sizeof(positionAndRotation) * (size_t)n);
                S[r].bestCosts = S[r].curCosts;
                float t = T > 1 ? (float)r / (float)(T - 1) : 0.0f;
                S[r].beta = (double)((float)b0 * exp2f(t * log2f((float)(b1 / b0))));
            }
            for (int it = 0; it < iterations; it++) {
                uint64_t git = o.iteration_offset + (uint64_t)it;
                for (int r = 0; r < T; r++) {
                    int c = l * T + r;
                    chain_step(&P, &S[r], &o, o.chain_offset + (uint64_t)c, git, S[r].beta,
                               trace ? &trace[(size_t)it * nChains + c] : NULL);
                }
                if ((git + 1) % (uint64_t)ex == 0) {
                    uint64_t epoch = (git + 1) / (uint64_t)ex;
                    for (int r = (int)(epoch & 1); r + 1 < T; r += 2) {
                        uint32_t w[4];
                        draw_block(o.seed, o.chain_offset + (uint64_t)(l * T + r), git, 0xFFFFu, w);
                        float u = uniform_from_bits(w[0]);
                        double Ea = -(double)S[r].curCosts.totalCosts, Eb = -(double)S[r + 1].curCosts.totalCosts;
                        float pacc = fminf(1.0f, (float)exp((S[r].beta - S[r + 1].beta) * (Ea - Eb)));
                        if (u < pacc) {
                            double tb = S[r].beta; S[r].beta = S[r + 1].beta; S[r + 1].beta = tb;
                        }
                    }
                }
            }
            for (int r = 0; r < T; r++) {
                int c = l * T + r;
                if (o.result_mode == 1) emit(&P, S[r].best, &S[r].bestCosts, points + (size_t)c * n, costs ? &costs[c] : NULL);
                else emit(&P, S[r].cur, &S[r].curCosts, points + (size_t)c * n, costs ? &costs[c] : NULL);
                free(S[r].cur); free(S[r].star); free(S[r].best);
            }
            free(S);
        }
        return threads;
    }

#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (int c = 0; c < nChains; c++) {
        chainState S;
        memset(&S, 0, sizeof S);
        S.cur = (positionAndRotation *)malloc(sizeof(positionAndRotation) * (size_t)n);
        S.star = (positionAndRotation *)malloc(sizeof(positionAndRotation) * (size_t)n);
        S.best = o.result_mode == 1 ? (positionAndRotation *)malloc(sizeof(positionAndRotation) * (size_t)n) : NULL;
        memcpy(S.cur, cfg, sizeof(positionAndRotation) * (size_t)n); /* K.cu:777 */
        Costs(srf, &S.curCosts, S.cur, rs, ra, vertices, clearances, offlimits, surfaceRectangle, NULL, 1); /* K.cu:778 */
        if (S.best) {
            memcpy(S.best, S.cur, sizeof(positionAndRotation) * (size_t)n);
            S.bestCosts = S.curCosts;
        }
        uint64_t chain = o.chain_offset + (uint64_t)c;
        for (int it = 0; it < iterations; it++) { /* K.cu:785 */
            uint64_t git = o.iteration_offset + (uint64_t)it;
            chain_step(&P, &S, &o, chain, git, beta_at(&o, git, iterations),
                       trace ? &trace[(size_t)it * nChains + c] : NULL);
        }
        if (S.best) emit(&P, S.best, &S.bestCosts, points + (size_t)c * n, costs ? &costs[c] : NULL);
        else emit(&P, S.cur, &S.curCosts, points + (size_t)c * n, costs ? &costs[c] : NULL);
        free(S.cur); free(S.star); free(S.best);
    }
    return threads;
}

/* Timed form for bench.py: returns wall seconds of the chain loops only. */
ORACLE_API double oracle_run_timed(const Surface *srf, const relationshipStruct *rs, const relationshipAngleStruct *ra,
                                   const positionAndRotation *cfg, const rectangle *clearances, const rectangle *offlimits,
                                   const vertex *vertices, const vertex *surfaceRectangle, int nChains, int iterations,
                                   const oracleRunOptions *opt, point *points, resultCosts *costs, int *threads_used)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int th = oracle_run(srf, rs, ra, cfg, clearances, offlimits, vertices, surfaceRectangle, nChains, iterations, opt, points,
                        costs, NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (threads_used) *threads_used = th;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

ORACLE_API int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
