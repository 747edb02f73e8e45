"""Development probe: where does the time of a KernelWrapperEx call go."""
import importlib, os, sys, time
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits one level below)
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
from importlib import import_module
B = import_module("metropolis-hastings-gpgpu_b200.binding")
k = pkg.Kernel(); room = pkg.synth.make_config(3); L = pkg.layout
chains, iters = 65536, 1000
k.wrapper_ex(room, 1024, 10, seed=1)
for rep in range(5):
    g = np.zeros(1, L.gpuConfig); g["gridxDim"], g["blockxDim"], g["iterations"] = chains, 64, iters
    o = B.make_options(seed=3)
    t0 = time.perf_counter()
    res = k.lib.KernelWrapperEx(*k._room_args(room), B._ptr(g), B._ptr(o))
    t1 = time.perf_counter()
    pts, costs = k._unpack(res, chains, room.n)
    t2 = time.perf_counter()
    print(f"KernelWrapperEx {1e3*(t1-t0):.1f} ms, python unpack+free {1e3*(t2-t1):.1f} ms")
    t0 = time.perf_counter()
    ctx = k.create(room, chains, seed=3); t1 = time.perf_counter()
    ctx.run(iters); ctx.synchronize(); t2 = time.perf_counter()
    p, c = ctx.results(); t3 = time.perf_counter()
    ms, _ = ctx.stats()
    ctx.close(); t4 = time.perf_counter()
    print(f"  create {1e3*(t1-t0):.1f} run {1e3*(t2-t1):.1f} (kernel {ms:.1f}) results {1e3*(t3-t2):.1f} destroy {1e3*(t4-t3):.1f} ms")
