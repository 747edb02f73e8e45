"""Throughput of the chain kernel per lane width (development probe, not a test)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits one level below)
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
k = pkg.Kernel()
print(k.device_info())
for cid, chains, iters in ((3, 65536, 200), (2, 65536, 500), (1, 65536, 1000), (4, 16384, 20), (3, 1024, 500)):
    room = pkg.synth.make_config(cid)
    for lanes in (1, 2, 4, 8, 16, 32):
        try:
            with k.create(room, chains, seed=1, lanes_per_chain=lanes) as ctx:
                ctx.run(2); ctx.synchronize(); ctx.stats()
                t0 = time.time(); ctx.run(iters); ctx.synchronize(); dt = time.time() - t0
                ms, _ = ctx.stats()
            print(f"cfg{cid} n={room.n} chains={chains} lanes={lanes}: {chains*iters/dt:.3e} proposals/s (wall {dt*1e3:.1f} ms)", flush=True)
        except Exception as e:
            print(f"cfg{cid} lanes={lanes}: {e}", flush=True)
