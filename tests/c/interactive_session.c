/*
 * The interactive use the reference's caller is built for (Readme.md:2: a layout tool asks for suggestions on every
 * user action), written against the additive part of the C ABI from plain C: ONE device-resident context for the
 * room (KernelCreate replaces the reference's per-call 12x cudaMalloc + curand init, Kernel.cu:879-943), stepped
 * several times (KernelRun), and after every step the handful of DIFFERENT best suggestions (KernelTopKDistinct) with
 * their layouts fetched by index.  Same room as call_kernel_wrapper.c (Kernel.cu:1007-1194).  Optional third argument:
 * number of devices to spread the chains over (mhOptions.n_devices; a repeated ordinal 0 on a one-GPU box).
 * tests/test_multi_device.py builds it with gcc, runs it and compares its output with the Python binding's.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mh_kernel.h"

#define N 32
#define NREL 1
#define NCLR 2
#define K 4

int main(int argc, char **argv)
{
    const int chains = argc > 1 ? atoi(argv[1]) : 256, steps = argc > 2 ? atoi(argv[2]) : 3, ndev = argc > 3 ? atoi(argv[3]) : 0;
    Surface srf;
    memset(&srf, 0, sizeof srf);
    srf.nObjs = N; srf.nRelationships = NREL; srf.nClearances = NCLR;
    srf.WeightFocalPoint = -2.0f; srf.WeightPairWise = -2.0f; srf.WeightVisualBalance = 1.5f; srf.WeightSymmetry = -2.0f;
    srf.WeightClearance = -2.0f; srf.WeightSurfaceArea = -2.0f; srf.WeightOffLimits = 0.0f;
    srf.focalX = 5.0; srf.focalY = 5.0;
    vertex room[4] = { { 10, 10, 0 }, { 10, 0, 0 }, { 0, 0, 0 }, { 0, 10, 0 } };
    const double vx[16] = { 2, 2, 0, 0, 3, 3, 1, 1, 2, 2, 0, 0, 3, 3, 1, 1 };
    const double vy[16] = { 2, 0, 0, 2, 2, 0, 0, 2, 2, 0, 0, 2, 2, 0, 0, 2 };
    vertex vtx[16];
    for (int i = 0; i < 16; i++) { vtx[i].x = vx[i]; vtx[i].y = vy[i]; vtx[i].z = 0; }
    rectangle clearances[NCLR] = { { 0, 1, 2, 3, 0 }, { 4, 5, 6, 7, 1 } };
    rectangle offlimits[N];
    positionAndRotation cfg[N];
    memset(cfg, 0, sizeof cfg);
    for (int i = 0; i < N; i++) {
        rectangle even = { 8, 9, 10, 11, 0 }, odd = { 12, 13, 14, 15, 1 };
        offlimits[i] = (i % 2 == 0) ? even : odd;
        cfg[i].x = i * 2.0; cfg[i].y = i * 2.0; cfg[i].length = 1.0; cfg[i].width = 1.0;
    }
    relationshipStruct rss[NREL];
    memset(rss, 0, sizeof rss);
    rss[0].TargetRange.targetRangeStart = 2.0; rss[0].TargetRange.targetRangeEnd = 4.0;
    rss[0].DegreesOfAtrraction = 2.0; rss[0].SourceIndex = 0; rss[0].TargetIndex = 1;
    relationshipAngleStruct rsa[NREL];
    memset(rsa, 0, sizeof rsa);
    rsa[0].angleMin = MH_PI / 4; rsa[0].angleMax = 5 * MH_PI / 8; rsa[0].SourceIndex = 0; rsa[0].TargetIndex = 1;

    mhOptions opt;
    memset(&opt, 0, sizeof opt);                                /* all zero = the reference's behaviour ... */
    opt.struct_size = (uint32_t)sizeof opt;
    opt.seed = 2026;                                            /* ... with a fixed seed */
    opt.result_mode = MH_RESULT_BEST;
    if (ndev > 1) {
        opt.n_devices = ndev > MH_MAX_DEVICES ? MH_MAX_DEVICES : ndev;
        const int visible = KernelDeviceCount();
        for (int i = 0; i < opt.n_devices; i++) opt.devices[i] = visible > 0 ? i % visible : 0;
    }
    mhContext *ctx = KernelCreate(rss, rsa, cfg, clearances, offlimits, vtx, room, &srf, chains, &opt);
    if (!ctx) { fprintf(stderr, "KernelCreate failed: %s\n", KernelLastError()); return 1; }
    point *pts = (point *)malloc(sizeof(point) * (size_t)chains * N);
    resultCosts *costs = (resultCosts *)malloc(sizeof(resultCosts) * (size_t)chains);
    if (!pts || !costs) return 3;
    for (int s = 0; s < steps; s++) {
        int idx[K], best = -1;
        float tot[K], bt = 0.f;
        if (KernelRun(ctx, 150)) { fprintf(stderr, "KernelRun failed: %s\n", KernelLastError()); return 1; }
        const int m = KernelTopKDistinct(ctx, K, 0.5f, 0.25f, idx, tot);
        if (m < 1 || KernelBest(ctx, &best, &bt) || best != idx[0] || bt != tot[0]) {
            fprintf(stderr, "ranking failed: %s\n", KernelLastError());
            return 1;
        }
        if (KernelResults(ctx, pts, costs)) { fprintf(stderr, "KernelResults failed: %s\n", KernelLastError()); return 1; }
        for (int j = 0; j < m; j++) {
            const point *p = pts + (size_t)idx[j] * N;
            if (costs[idx[j]].totalCosts != tot[j]) return 4;   /* the ranking's totals are the results' totals */
            printf("step %d pick %d chain %d total %a first %a %a %a\n", s, j, idx[j], tot[j], p[0].x, p[0].y, p[0].rotY);
        }
    }
    double ms = 0;
    long long launches = 0;
    if (KernelStats(ctx, &ms, &launches) || launches < steps) return 5;
    KernelDestroy(ctx);
    free(pts); free(costs);
    return 0;
}
