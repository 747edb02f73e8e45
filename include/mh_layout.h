/*
 * mh_layout.h -- wire format of the Metropolis-Hastings layout optimiser.
 *
 * These are the blittable structs a caller (the NarrativeWorldCreator C# P/Invoke
 * wrapper, or any C program) passes to KernelWrapper.  The reference keeps them only
 * inside its translation unit (Kernel.h is empty); the field order, types, sizes and
 * offsets below restate /root/reference/KernelFolder/Kernel/Kernel.cu:43-149 and are
 * pinned with static assertions so that a layout drift is a compile error.
 *
 * Plain C99 / C++11; no CUDA types.
 */
#ifndef MH_LAYOUT_H
#define MH_LAYOUT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define MH_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define MH_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* Kernel.cu:43-48.  Rectangle corner; z is carried but never read. */
typedef struct vertex {
    double x;
    double y;
    double z;
} vertex;

/* Kernel.cu:50-57.  A rectangle is the four CONSECUTIVE vertices starting at
 * point1Index (Kernel.cu:371-379); point2..4Index are carried but never read. */
typedef struct rectangle {
    int32_t point1Index;
    int32_t point2Index;
    int32_t point3Index;
    int32_t point4Index;
    int32_t SourceIndex;
} rectangle;

/* Kernel.cu:59-72.  Per-object state.  `frozen` is read as ONE byte at offset 48
 * (a C# bool marshalled as 4-byte BOOL lands in the same low byte). */
typedef struct positionAndRotation {
    double x;
    double y;
    double z;
    double rotX;
    double rotY;
    double rotZ;
    uint8_t frozen;
    uint8_t _pad[7];
    double length;
    double width;
} positionAndRotation;

/* Kernel.cu:74-77 */
typedef struct targetRangeStruct {
    double targetRangeStart;
    double targetRangeEnd;
} targetRangeStruct;

/* Kernel.cu:79-85.  DegreesOfAtrraction (sic) is carried but never read. */
typedef struct relationshipStruct {
    targetRangeStruct TargetRange;
    int32_t SourceIndex;
    int32_t TargetIndex;
    double DegreesOfAtrraction;
} relationshipStruct;

/* Kernel.cu:87-92 */
typedef struct relationshipAngleStruct {
    double angleMin;
    double angleMax;
    int32_t SourceIndex;
    int32_t TargetIndex;
} relationshipAngleStruct;

/* Kernel.cu:94-117.  Counts, the seven weights, centroid and focal point. */
typedef struct Surface {
    int32_t nObjs;
    int32_t nRelationships;
    int32_t nClearances;
    float WeightFocalPoint;
    float WeightPairWise;
    float WeightVisualBalance;
    float WeightSymmetry;
    float WeightOffLimits;
    float WeightClearance;
    float WeightSurfaceArea;
    double centroidX;
    double centroidY;
    double focalX;
    double focalY;
    double focalRot;
} Surface;

/* Kernel.cu:119-127.  gridxDim = number of chains (= number of results),
 * iterations = MH steps per chain.  blockxDim was the reference's threads per block, a
 * tuning knob without algorithmic meaning: accepted and ignored.  The other three are
 * unused by the reference as well (Kernel.cu:946-947). */
typedef struct gpuConfig {
    int32_t gridxDim;
    int32_t gridyDim;
    int32_t blockxDim;
    int32_t blockyDim;
    int32_t blockzDim;
    int32_t iterations;
} gpuConfig;

/* Kernel.cu:129-132 */
typedef struct point {
    float x, y, z, rotX, rotY, rotZ;
} point;

/* Kernel.cu:134-144 */
typedef struct resultCosts {
    float totalCosts;
    float PairWiseCosts;
    float VisualBalanceCosts;
    float FocalPointCosts;
    float SymmetryCosts;
    float ClearanceCosts;
    float OffLimitsCosts;
    float SurfaceAreaCosts;
} resultCosts;

/* Kernel.cu:146-149 */
typedef struct result {
    point *points;
    resultCosts costs;
} result;

/* ---- layout pins (identical on MSVC x64 and SysV x86-64) ---- */
MH_STATIC_ASSERT(sizeof(vertex) == 24, "vertex");
MH_STATIC_ASSERT(sizeof(rectangle) == 20, "rectangle");
MH_STATIC_ASSERT(offsetof(rectangle, SourceIndex) == 16, "rectangle.SourceIndex");
MH_STATIC_ASSERT(sizeof(positionAndRotation) == 72, "positionAndRotation");
MH_STATIC_ASSERT(offsetof(positionAndRotation, rotY) == 32, "positionAndRotation.rotY");
MH_STATIC_ASSERT(offsetof(positionAndRotation, frozen) == 48, "positionAndRotation.frozen");
MH_STATIC_ASSERT(offsetof(positionAndRotation, length) == 56, "positionAndRotation.length");
MH_STATIC_ASSERT(offsetof(positionAndRotation, width) == 64, "positionAndRotation.width");
MH_STATIC_ASSERT(sizeof(targetRangeStruct) == 16, "targetRangeStruct");
MH_STATIC_ASSERT(sizeof(relationshipStruct) == 32, "relationshipStruct");
MH_STATIC_ASSERT(offsetof(relationshipStruct, SourceIndex) == 16, "relationshipStruct.SourceIndex");
MH_STATIC_ASSERT(offsetof(relationshipStruct, TargetIndex) == 20, "relationshipStruct.TargetIndex");
MH_STATIC_ASSERT(offsetof(relationshipStruct, DegreesOfAtrraction) == 24, "relationshipStruct.Degrees");
MH_STATIC_ASSERT(sizeof(relationshipAngleStruct) == 24, "relationshipAngleStruct");
MH_STATIC_ASSERT(offsetof(relationshipAngleStruct, SourceIndex) == 16, "relationshipAngleStruct.SourceIndex");
MH_STATIC_ASSERT(sizeof(Surface) == 80, "Surface");
MH_STATIC_ASSERT(offsetof(Surface, WeightFocalPoint) == 12, "Surface.WeightFocalPoint");
MH_STATIC_ASSERT(offsetof(Surface, WeightSurfaceArea) == 36, "Surface.WeightSurfaceArea");
MH_STATIC_ASSERT(offsetof(Surface, centroidX) == 40, "Surface.centroidX");
MH_STATIC_ASSERT(offsetof(Surface, focalRot) == 72, "Surface.focalRot");
MH_STATIC_ASSERT(sizeof(gpuConfig) == 24, "gpuConfig");
MH_STATIC_ASSERT(offsetof(gpuConfig, iterations) == 20, "gpuConfig.iterations");
MH_STATIC_ASSERT(sizeof(point) == 24, "point");
MH_STATIC_ASSERT(sizeof(resultCosts) == 32, "resultCosts");
MH_STATIC_ASSERT(sizeof(result) == 40, "result");
MH_STATIC_ASSERT(offsetof(result, costs) == 8, "result.costs");

/* Constants of the reference (Kernel.cu:31-39).  PI is 3.1416 on purpose (quirk Q4). */
#define MH_PI 3.1416
#define MH_BETA 2.0
#define MH_S_SIGMA_T (15.0 / 90.0 * MH_PI)

#endif /* MH_LAYOUT_H */
