"""numpy mirrors of the wire structs in include/mh_layout.h.

Field order, sizes and offsets restate /root/reference/KernelFolder/Kernel/Kernel.cu:43-149
(SURVEY.md section 8b); `check_layout()` re-asserts every size so that a drift between this
file and the C header is caught by the CPU test suite.
"""
import numpy as np

PI = 3.1416  # Kernel.cu:31 (quirk Q4)


def _dt(fields, itemsize):
    names = [f[0] for f in fields]
    formats = [(f[1], f[3]) if len(f) > 3 else f[1] for f in fields]
    offsets = [f[2] for f in fields]
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": itemsize})


vertex = _dt([("x", "<f8", 0), ("y", "<f8", 8), ("z", "<f8", 16)], 24)
rectangle = _dt(
    [("point1Index", "<i4", 0), ("point2Index", "<i4", 4), ("point3Index", "<i4", 8), ("point4Index", "<i4", 12), ("SourceIndex", "<i4", 16)],
    20,
)
positionAndRotation = _dt(
    [("x", "<f8", 0), ("y", "<f8", 8), ("z", "<f8", 16), ("rotX", "<f8", 24), ("rotY", "<f8", 32), ("rotZ", "<f8", 40),
     ("frozen", "u1", 48), ("length", "<f8", 56), ("width", "<f8", 64)],
    72,
)
relationshipStruct = _dt(
    [("targetRangeStart", "<f8", 0), ("targetRangeEnd", "<f8", 8), ("SourceIndex", "<i4", 16), ("TargetIndex", "<i4", 20),
     ("DegreesOfAtrraction", "<f8", 24)],
    32,
)
relationshipAngleStruct = _dt([("angleMin", "<f8", 0), ("angleMax", "<f8", 8), ("SourceIndex", "<i4", 16), ("TargetIndex", "<i4", 20)], 24)
Surface = _dt(
    [("nObjs", "<i4", 0), ("nRelationships", "<i4", 4), ("nClearances", "<i4", 8), ("WeightFocalPoint", "<f4", 12),
     ("WeightPairWise", "<f4", 16), ("WeightVisualBalance", "<f4", 20), ("WeightSymmetry", "<f4", 24), ("WeightOffLimits", "<f4", 28),
     ("WeightClearance", "<f4", 32), ("WeightSurfaceArea", "<f4", 36), ("centroidX", "<f8", 40), ("centroidY", "<f8", 48),
     ("focalX", "<f8", 56), ("focalY", "<f8", 64), ("focalRot", "<f8", 72)],
    80,
)
gpuConfig = _dt(
    [("gridxDim", "<i4", 0), ("gridyDim", "<i4", 4), ("blockxDim", "<i4", 8), ("blockyDim", "<i4", 12), ("blockzDim", "<i4", 16),
     ("iterations", "<i4", 20)],
    24,
)
point = _dt([("x", "<f4", 0), ("y", "<f4", 4), ("z", "<f4", 8), ("rotX", "<f4", 12), ("rotY", "<f4", 16), ("rotZ", "<f4", 20)], 24)
COST_FIELDS = ("totalCosts", "PairWiseCosts", "VisualBalanceCosts", "FocalPointCosts", "SymmetryCosts", "ClearanceCosts",
               "OffLimitsCosts", "SurfaceAreaCosts")
resultCosts = _dt([(f, "<f4", 4 * i) for i, f in enumerate(COST_FIELDS)], 32)
result = _dt([("points", "<u8", 0), ("costs", resultCosts, 8)], 40)

# include/mh_kernel.h
mhOptions = _dt(
    [("struct_size", "<u4", 0), ("flags", "<u4", 4), ("seed", "<u8", 8), ("chain_offset", "<u8", 16), ("iteration_offset", "<u8", 24),
     ("beta_start", "<f8", 32), ("beta_end", "<f8", 40), ("schedule", "<i4", 48), ("schedule_length", "<i4", 52),
     ("result_mode", "<i4", 56), ("eval_mode", "<i4", 60), ("lanes_per_chain", "<i4", 64), ("device", "<i4", 68),
     ("tempering_rungs", "<i4", 72), ("exchange_interval", "<i4", 76), ("chain_stride", "<u8", 80),
     ("total_chains", "<u8", 88), ("n_devices", "<i4", 96), ("devices", "<i4", 100, (8,)), ("reserved0", "<i4", 132)],
    136,
)
MH_OPT_EXPLICIT_DEVICE = 1
MH_MAX_DEVICES = 8
mhTraceEntry = _dt(
    [("move", "<i4", 0), ("obj1", "<i4", 4), ("obj2", "<i4", 8), ("accepted", "<i4", 12), ("star_total", "<f4", 16),
     ("cur_total", "<f4", 20), ("u", "<f4", 24), ("beta", "<f4", 28)],
    32,
)

SIZES = {"vertex": 24, "rectangle": 20, "positionAndRotation": 72, "relationshipStruct": 32, "relationshipAngleStruct": 24,
         "Surface": 80, "gpuConfig": 24, "point": 24, "resultCosts": 32, "result": 40}


def check_layout():
    g = globals()
    for name, size in SIZES.items():
        assert g[name].itemsize == size, (name, g[name].itemsize, size)
    assert positionAndRotation.fields["frozen"][1] == 48
    assert Surface.fields["focalRot"][1] == 72
    return True
