"""development probe: kernel-only rate of one libKernel build (env MH_LIB) over a matrix of runs.
usage: MH_LIB=... ab_matrix.py <tag> cfg:chains:iters:lanes:mode ..."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
tag = sys.argv[1]
k = pkg.Kernel()
for spec in sys.argv[2:]:
    cid, chains, iters, lanes, mode = spec.split(":")
    chains, iters, lanes, mode = int(chains), int(iters), int(lanes), int(mode)
    if "x" in cid:                                              # custom room "NxCxR" on a 12 x 9 floor
        room = pkg.synth.make_room(*[int(v) for v in cid.split("x")], 12.0, 9.0, 4242)
    else:
        cid = int(cid)
        room = pkg.synth.make_config(cid)
    try:
        with k.create(room, chains, seed=1, lanes_per_chain=lanes, eval_mode=mode) as ctx:
            ctx.run(iters); ctx.synchronize(); ms0, _ = ctx.stats()
            best = 1e30
            for _ in range(3):
                ctx.reset(); ctx.run(iters); ctx.synchronize()
                ms1, _ = ctx.stats(); best = min(best, ms1 - ms0); ms0 = ms1
        print(f"{tag:8s} cfg{cid} chains={chains} iters={iters} G={lanes:2d} mode={mode}: {best:8.2f} ms  {chains*iters/(best*1e-3):.4e} /s", flush=True)
    except Exception as e:
        print(f"{tag:8s} cfg{cid} G={lanes} mode={mode}: FAILED {str(e)[:80]}", flush=True)
