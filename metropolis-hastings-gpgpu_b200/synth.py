"""Synthetic rooms and the reference's own smoke fixture.

One generator shared by the tests, the oracle drivers and bench.py, so that every arm is fed
byte-identical inputs.  PRNG: SplitMix64, u = (x >> 11) * 2**-53, seed 0x5EED0000 + config id.
"""
import math
from dataclasses import dataclass

import numpy as np

from . import layout as L

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def u(self):
        return (self.next() >> 11) * (1.0 / (1 << 53))

    def uniform(self, a, b):
        return a + (b - a) * self.u()

    def randint(self, n):
        return min(int(self.u() * n), n - 1)


@dataclass
class Room:
    """The nine KernelWrapper arguments minus gpuConfig, as numpy struct arrays."""
    srf: np.ndarray
    rss: np.ndarray
    rsa: np.ndarray
    cfg: np.ndarray
    clearances: np.ndarray
    offlimits: np.ndarray
    vertices: np.ndarray
    surfaceRectangle: np.ndarray
    name: str = ""

    @property
    def n(self):
        return int(self.srf["nObjs"][0])

    @property
    def C(self):
        return int(self.srf["nClearances"][0])

    @property
    def R(self):
        return int(self.srf["nRelationships"][0])

    def flops_per_proposal(self, live=False):
        """F = 20n^2 + 10Cn + 70n + 50C + 34R + 51 (contract figure).
        live=True leaves out the 5n(n-1) of the off-limits term, which the acceptance test
        never reads (quirk Q5) and which the kernel therefore evaluates once per result."""
        n, C, R = self.n, self.C, self.R
        f = 20 * n * n + 10 * C * n + 70 * n + 50 * C + 34 * R + 51
        return f - 5 * n * (n - 1) if live else f


# (n, C, R, W, H) of BASELINE.json configs 1-4; config 5 uses the room of config 3
CONFIGS = {1: (8, 4, 2, 4.0, 4.0), 2: (16, 8, 16, 5.0, 4.0), 3: (50, 25, 50, 8.0, 6.0), 4: (200, 100, 200, 20.0, 15.0)}
CONFIG_NAMES = {1: "cfg1 single room n=8", 2: "cfg2 bedroom n=16", 3: "cfg3 living room n=50", 4: "cfg4 hall n=200"}


def make_room(n, C, R, W, H, seed, name=""):
    assert C <= n, "quirk Q7 needs C <= n"
    g = SplitMix64(seed)
    srf = np.zeros(1, L.Surface)
    srf["nObjs"], srf["nRelationships"], srf["nClearances"] = n, R, C
    srf["WeightFocalPoint"] = -2.0
    srf["WeightPairWise"] = -2.0
    srf["WeightVisualBalance"] = 1.5
    srf["WeightSymmetry"] = -2.0
    srf["WeightOffLimits"] = -2.0
    srf["WeightClearance"] = -2.0
    srf["WeightSurfaceArea"] = -2.0
    srf["centroidX"], srf["centroidY"] = W / 2, H / 2
    srf["focalX"], srf["focalY"], srf["focalRot"] = W / 2, 0.0, math.pi / 2

    sr = np.zeros(4, L.vertex)
    sr["x"] = [W, W, 0, 0]
    sr["y"] = [H, 0, 0, H]

    cfg = np.zeros(n, L.positionAndRotation)
    for i in range(n):
        cfg["length"][i] = g.uniform(0.4, 2.0)
        cfg["width"][i] = g.uniform(0.4, 1.2)
    for i in range(n):
        cfg["x"][i] = g.uniform(0, W)
        cfg["y"][i] = g.uniform(0, H)
        cfg["rotY"][i] = g.uniform(0, 2 * L.PI)

    vertices = np.zeros(4 * C + 4 * n, L.vertex)
    offl = np.zeros(n, L.rectangle)
    for i in range(n):
        l, w = cfg["length"][i], cfg["width"][i]
        b = 4 * C + 4 * i
        vertices["x"][b:b + 4] = [l / 2, l / 2, -l / 2, -l / 2]
        vertices["y"][b:b + 4] = [w / 2, -w / 2, -w / 2, w / 2]
        offl[i] = (b, b + 1, b + 2, b + 3, i)
    clr = np.zeros(C, L.rectangle)
    for c in range(C):
        s = (2 * c + 1) % n
        l, w = cfg["length"][s], cfg["width"][s]
        b = 4 * c
        vertices["x"][b:b + 4] = [l / 2 + 0.5, l / 2 + 0.5, -l / 2, -l / 2]
        vertices["y"][b:b + 4] = [w / 2, -w / 2, -w / 2, w / 2]
        clr[c] = (b, b + 1, b + 2, b + 3, s)

    rss = np.zeros(R, L.relationshipStruct)
    rsa = np.zeros(R, L.relationshipAngleStruct)
    for r in range(R):
        s = g.randint(n)
        t = g.randint(n)
        while t == s:
            t = g.randint(n)
        start = g.uniform(0.5, 1.5)
        end = start + g.uniform(0.5, 1.5)
        amin = g.uniform(0, 2 * L.PI)
        amax = math.fmod(amin + g.uniform(math.pi / 8, math.pi / 2), 2 * L.PI)
        rss[r] = (start, end, s, t, 1.0)
        rsa[r] = (amin, amax, s, t)
    return Room(srf, rss, rsa, cfg, clr, offl, vertices, sr, name)


def make_config(cfg_id, seed=None):
    n, C, R, W, H = CONFIGS[cfg_id]
    return make_room(n, C, R, W, H, (0x5EED0000 + cfg_id) if seed is None else seed, CONFIG_NAMES[cfg_id])


def reference_main_fixture():
    """The only input the reference's authors wrote down: main(), Kernel.cu:1007-1166.
    WeightOffLimits is left uninitialised there; it is 0 here."""
    N, C, R = 32, 2, 1
    srf = np.zeros(1, L.Surface)
    srf["nObjs"], srf["nRelationships"], srf["nClearances"] = N, R, C
    srf["WeightFocalPoint"] = -2.0
    srf["WeightPairWise"] = -2.0
    srf["WeightVisualBalance"] = 1.5
    srf["WeightSymmetry"] = -2.0
    srf["WeightClearance"] = -2.0
    srf["WeightSurfaceArea"] = -2.0
    srf["WeightOffLimits"] = 0.0
    srf["focalX"], srf["focalY"] = 5.0, 5.0
    sr = np.zeros(4, L.vertex)
    sr["x"] = [10, 10, 0, 0]
    sr["y"] = [10, 0, 0, 10]
    vtx = np.zeros(16, L.vertex)
    vtx["x"] = [2, 2, 0, 0, 3, 3, 1, 1, 2, 2, 0, 0, 3, 3, 1, 1]
    vtx["y"] = [2, 0, 0, 2, 2, 0, 0, 2, 2, 0, 0, 2, 2, 0, 0, 2]
    clr = np.zeros(C, L.rectangle)
    clr[0] = (0, 1, 2, 3, 0)
    clr[1] = (4, 5, 6, 7, 1)
    offl = np.zeros(N, L.rectangle)
    for i in range(N):
        offl[i] = (8, 9, 10, 11, 0) if i % 2 == 0 else (12, 13, 14, 15, 1)
    cfg = np.zeros(N, L.positionAndRotation)
    cfg["x"] = 2.0 * np.arange(N)
    cfg["y"] = 2.0 * np.arange(N)
    cfg["length"] = 1.0
    cfg["width"] = 1.0
    rss = np.zeros(R, L.relationshipStruct)
    rss[0] = (2.0, 4.0, 0, 1, 2.0)
    rsa = np.zeros(R, L.relationshipAngleStruct)
    rsa[0] = (L.PI / 4, 5 * L.PI / 8, 0, 1)
    return Room(srf, rss, rsa, cfg, clr, offl, vtx, sr, "reference main() fixture")


def random_layouts(room, count, seed, spread=1.25, f32=True):
    """`count` layouts of room.n objects: positions uniform over the room grown by `spread`
    (so that the surface-area term is exercised), rotations uniform over [0, 2*PI].
    f32=True rounds x, y, rotY to float32-representable doubles, which is what a chain's
    state looks like inside the float32 kernel."""
    g = np.random.default_rng(seed)
    n = room.n
    W = float(room.surfaceRectangle["x"].max() - room.surfaceRectangle["x"].min())
    H = float(room.surfaceRectangle["y"].max() - room.surfaceRectangle["y"].min())
    x0 = float(room.surfaceRectangle["x"].min())
    y0 = float(room.surfaceRectangle["y"].min())
    lay = np.tile(room.cfg, count)
    lay["x"] = x0 + W * (0.5 + spread * (g.random(count * n) - 0.5))
    lay["y"] = y0 + H * (0.5 + spread * (g.random(count * n) - 0.5))
    lay["rotY"] = 2 * L.PI * g.random(count * n)
    if f32:
        for f in ("x", "y", "rotY"):
            lay[f] = lay[f].astype(np.float32).astype(np.float64)
    return lay


def make_wild_room(n, C, R, seed):
    """A room that breaks every convenience of make_room: the surface is an arbitrary quadrilateral
    away from the origin, rectangles are arbitrary 4-vertex sets in arbitrary order (so the first
    vertex is not special except through quirk Q6), several clearances may share a source, a
    relationship may name the same object twice or use different pairs for distance and angle, the
    focal axis has any rotation (exercising both one-sided wraps of quirk Q18), weights have any sign
    (some zero), some objects are frozen.  Used by the parity tests only."""
    g = np.random.default_rng(seed)
    srf = np.zeros(1, L.Surface)
    srf["nObjs"], srf["nRelationships"], srf["nClearances"] = n, R, C
    for f in ("WeightFocalPoint", "WeightPairWise", "WeightVisualBalance", "WeightSymmetry", "WeightOffLimits", "WeightClearance",
              "WeightSurfaceArea"):
        srf[f] = 0.0 if g.random() < 0.15 else g.uniform(-3, 3)
    ox, oy = g.uniform(-20, 20, 2)
    W, H = g.uniform(3, 15, 2)
    srf["centroidX"], srf["centroidY"] = ox + g.uniform(0, W), oy + g.uniform(0, H)
    srf["focalX"], srf["focalY"] = ox + g.uniform(-1, W + 1), oy + g.uniform(-1, H + 1)
    srf["focalRot"] = g.uniform(-2 * math.pi, 2 * math.pi)
    sr = np.zeros(4, L.vertex)
    corners = np.array([[ox + W, oy + H], [ox + W, oy], [ox, oy], [ox, oy + H]]) + g.uniform(-0.3, 0.3, (4, 2))
    corners = corners[g.permutation(4)]
    sr["x"], sr["y"] = corners[:, 0], corners[:, 1]
    cfg = np.zeros(n, L.positionAndRotation)
    cfg["length"] = g.uniform(0.2, 2.5, n)
    cfg["width"] = g.uniform(0.2, 2.5, n)
    cfg["x"] = ox + g.uniform(-1, W + 1, n)
    cfg["y"] = oy + g.uniform(-1, H + 1, n)
    cfg["rotY"] = g.uniform(0, 2 * L.PI, n)
    cfg["z"] = g.uniform(-1, 1, n)
    cfg["rotX"] = g.uniform(-1, 1, n)
    cfg["rotZ"] = g.uniform(-1, 1, n)
    cfg["frozen"] = (g.random(n) < 0.2).astype(np.uint8)
    if n > 1:
        cfg["frozen"][g.integers(n)] = 0
    vertices = np.zeros(4 * C + 4 * n, L.vertex)
    vertices["x"] = g.uniform(-1.5, 1.5, len(vertices))
    vertices["y"] = g.uniform(-1.5, 1.5, len(vertices))
    vertices["z"] = g.uniform(-1, 1, len(vertices))
    # rectangles may start anywhere in the pool as long as 4 consecutive vertices exist
    offl = np.zeros(n, L.rectangle)
    for i in range(n):
        b = int(g.integers(0, len(vertices) - 3))
        offl[i] = (b, int(g.integers(0, 99)), int(g.integers(0, 99)), int(g.integers(0, 99)), int(g.integers(0, 99)))
    clr = np.zeros(C, L.rectangle)
    for c in range(C):
        b = int(g.integers(0, len(vertices) - 3))
        clr[c] = (b, 0, 0, 0, int(g.integers(0, n)))
    rss = np.zeros(R, L.relationshipStruct)
    rsa = np.zeros(R, L.relationshipAngleStruct)
    for r in range(R):
        s, t = int(g.integers(0, n)), int(g.integers(0, n))
        s2, t2 = (s, t) if g.random() < 0.7 else (int(g.integers(0, n)), int(g.integers(0, n)))
        start = g.uniform(0.2, 3.0)
        rss[r] = (start, start + g.uniform(0.0, 3.0), s, t, g.uniform(0, 5))
        rsa[r] = (g.uniform(0, 2 * L.PI), g.uniform(0, 2 * L.PI), s2, t2)
    return Room(srf, rss, rsa, cfg, clr, offl, vertices, sr, f"wild room seed {seed}")
