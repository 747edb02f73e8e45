/*
 * A plain C caller of the drop-in: fills the wire structs by hand, the way the reference's own
 * smoke driver does (same room as Kernel.cu:1007-1194: 32 objects on a diagonal, 2 clearances,
 * 1 relationship, 10x10 surface, 100 iterations), calls KernelWrapper through the C header and
 * prints what comes back.  tests/test_gpu_parity.py builds it with gcc against libKernel.so and
 * compares the output with the Python binding: the ABI is exercised from real C, not only ctypes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mh_kernel.h"

#define N 32
#define NREL 1
#define NCLR 2

int main(int argc, char **argv)
{
    int chains = argc > 1 ? atoi(argv[1]) : 1, iterations = argc > 2 ? atoi(argv[2]) : 100;
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
    gpuConfig g = { chains, 0, 64, 0, 0, iterations };

    result *res = KernelWrapper(rss, rsa, cfg, clearances, offlimits, vtx, room, &srf, &g);
    if (!res) {
        fprintf(stderr, "KernelWrapper failed: %s\n", KernelLastError());
        return 1;
    }
    for (int i = 0; i < chains; i++) {
        const resultCosts *c = &res[i].costs;
        printf("costs %d %a %a %a %a %a %a %a %a\n", i, c->totalCosts, c->PairWiseCosts, c->VisualBalanceCosts, c->FocalPointCosts,
               c->SymmetryCosts, c->ClearanceCosts, c->OffLimitsCosts, c->SurfaceAreaCosts);
        for (int j = 0; j < N; j++) {
            const point *p = &res[i].points[j];
            printf("point %d %d %a %a %a %a %a %a\n", i, j, p->x, p->y, p->z, p->rotX, p->rotY, p->rotZ);
        }
    }
    if (res[0].points + (size_t)(chains - 1) * N != res[chains - 1].points) return 2; /* one block (Kernel.cu:981) */
    free(res[0].points); /* plain malloc: the caller may free() without KernelFree */
    free(res);
    return 0;
}
