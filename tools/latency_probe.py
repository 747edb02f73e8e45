"""development probe: wall time of small one-shot KernelWrapperEx calls (the interactive use of the reference's caller)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
k = pkg.Kernel()
for cid in (1, 2, 3):
    room = pkg.synth.make_config(cid)
    k.wrapper_ex(room, 64, 10, seed=1)
    for chains, iters in ((1, 1000), (64, 200), (1024, 200), (1024, 1000), (4096, 1000)):
        ts = []
        for rep in range(5):
            t0 = time.perf_counter()
            k.wrapper_ex(room, chains, iters, seed=rep)
            ts.append(time.perf_counter() - t0)
        print(f"cfg{cid} n={room.n} chains={chains} iters={iters}: min {1e3*min(ts):.2f} ms  median {1e3*sorted(ts)[2]:.2f} ms", flush=True)
