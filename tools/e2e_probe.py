"""Development probe: where does the time of a KernelWrapperEx call go (MH_TIMING=1 prints the library's own phases).
usage: e2e_probe.py [config id] [chains] [iterations]   (env MH_PIN_RESULT=1: page-lock the result block for the copy)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits one level below)
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
k = pkg.Kernel(); room = pkg.synth.make_config(cid)
k.wrapper_ex(room, 1024, 10, seed=1)
print(f"config {cid}, {chains} chains x {iters} iterations, result block {chains * room.n * 24 / 1e6:.0f} MB, MH_PIN_RESULT={os.environ.get('MH_PIN_RESULT', '')}")
for rep in range(4):
    t0 = time.perf_counter()
    res, pts, costs = k.wrapper_ex_raw(room, chains, iters, seed=3)
    t1 = time.perf_counter()
    best = float(costs["totalCosts"].max())
    k.free(res)
    t2 = time.perf_counter()
    print(f"KernelWrapperEx {1e3*(t1-t0):.1f} ms, read + free {1e3*(t2-t1):.1f} ms")
