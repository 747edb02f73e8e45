"""Profiling target: one warm-up launch and one measured launch of the chain kernel.
usage: prof_target.py <config id> <chains> <iterations> [lanes]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (tools/ sits one level below)
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
spec = sys.argv[1]
cid = int(spec) if spec.isdigit() else spec
chains, iters = int(sys.argv[2]), int(sys.argv[3])
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
k = pkg.Kernel()
room = pkg.synth.make_config(cid) if isinstance(cid, int) else pkg.synth.make_room(*[int(v) for v in cid.split(',')], 12.0, 9.0, 4242)  # custom 'n,C,R'
with k.create(room, chains, seed=1, lanes_per_chain=lanes, eval_mode=mode) as ctx:
    ctx.run(iters); ctx.synchronize(); ms_warm, _ = ctx.stats()
    ctx.reset()
    t0 = time.time(); ctx.run(iters); ctx.synchronize(); dt = time.time() - t0
    ms0 = ms_warm if 'ms_warm' in dir() else 0.0
    ms, n = ctx.stats(); ms -= ms0
    print(f"cfg{cid} chains={chains} iters={iters} lanes={lanes} mode={mode}: kernel {ms:.2f} ms, {chains*iters/(ms*1e-3):.4e} proposals/s")
