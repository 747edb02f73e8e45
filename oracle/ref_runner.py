"""Runs the reference's own KernelWrapper (oracle/_ref/*.so) in THIS process and prints one JSON
line.  Test/baseline infrastructure.  Always launched as a subprocess with a timeout by bench.py
and the probes: the unmodified reference deadlocks on sm_70+ for blockxDim > 1 (divergent
__syncthreads, Kernel.cu:747 under Kernel.cu:819), and a hang must not take the caller down.

usage: ref_runner.py <variant: plain|nb> <config id> <chains> <iterations> <blockxDim> <warmup> <steps> [finals.npy]

Reports the whole-call rate (what a caller of the reference sees: its H2D/D2H, its initRNG launch and its
kernel), the device-event time of the whole call, and the device time of the initRNG launch alone
(Kernel.cu:939-943, measured by RefInitRngMs at the same launch shape) so that the kernel-only rate can be
quoted without the RNG set-up.  With a last argument, the oracle's totalCosts of the layouts the reference
returned are saved there (the reference's own `costs` are uninitialised memory, quirk Q3): the sample of the
informational KS test against the reference kernel.
"""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    variant, cid, chains, iters, block, warmup, steps = sys.argv[1], *map(int, sys.argv[2:8])
    finals_path = sys.argv[8] if len(sys.argv) > 8 else None
    pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
    import oracle_lib
    path = os.path.join(ROOT, "oracle", "_ref", "libKernel_ref_nb.so" if variant == "nb" else "libKernel_ref.so")
    oracle_lib.ref_gpu_path = lambda: path
    ref = oracle_lib.RefGPU()
    room = pkg.synth.make_config(cid)
    heap = max(64 << 20, int(chains * (2 * 72 * room.n + 4 * 96) * 1.5))
    rc = ref.set_heap(heap)
    for _ in range(warmup):
        ref.run(room, chains, iters, block)
    wall, dev = [], []
    pts = None
    for _ in range(steps):
        t0 = time.perf_counter()
        pts, ms = ref.run(room, chains, iters, block)
        wall.append(time.perf_counter() - t0)
        dev.append(ms * 1e-3)
    o = oracle_lib.Oracle()
    lay = np.tile(room.cfg, chains)
    for f in ("x", "y", "z", "rotX", "rotY", "rotZ"):
        lay[f] = pts[f].reshape(-1)
    finite = bool(np.isfinite(pts["x"]).all() and np.isfinite(pts["rotY"]).all())
    tot = o.costs_batch(room, lay)["totalCosts"] if finite else np.array([np.nan])
    if finals_path:
        np.save(finals_path, tot)
    init_ms = float(np.median([ref.init_rng_ms(chains, block) for _ in range(3)]))
    print(json.dumps({"init_rng_ms": init_ms, "proposals_per_s_dev_without_init_rng": chains * iters / max(float(np.mean(dev)) - init_ms * 1e-3, 1e-9),
                      "variant": variant, "config": cid, "chains": chains, "iterations": iters, "block": block, "heap_rc": rc,
                      "wall_s": float(np.mean(wall)), "dev_s": float(np.mean(dev)), "proposals_per_s": chains * iters / float(np.mean(wall)),
                      "finite": finite, "mean_total": float(np.mean(tot)), "initial_total": float(o.costs(room)["totalCosts"])}), flush=True)


if __name__ == "__main__":
    main()
