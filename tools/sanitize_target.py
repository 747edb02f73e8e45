"""compute-sanitizer target: every kernel of the library once, at sizes a sanitizer run finishes in seconds --
the scan kernel, the memo and delta kernels (several lane widths), tempering + exchange, scoring, the ranking
kernels, a multi-shard context.  usage: compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
k = pkg.Kernel()
S = pkg.synth
for cid, chains, iters, lanes, mode in ((1, 40, 30, 0, 0), (2, 40, 30, 4, 3), (2, 40, 30, 2, 2), (3, 24, 20, 8, 0), (3, 24, 20, 8, 1),
                                        (3, 8, 20, 32, 2), (4, 4, 6, 32, 0), (3, 24, 10, 4, 3)):
    room = S.make_config(cid)
    p, c = k.wrapper_ex(room, chains, iters, seed=5, lanes_per_chain=lanes, eval_mode=mode, result_mode=cid % 2)
    print("ok", cid, chains, iters, lanes, mode, float(c["totalCosts"].max()), flush=True)
room = S.make_config(2)
with k.create(room, 64, seed=2, beta_start=0.5, beta_end=8.0, tempering_rungs=4, exchange_interval=5) as ctx:
    ctx.run(22)
    print("tempering", ctx.best(), ctx.top_k(5)[0], ctx.top_k_distinct(4, 0.5)[0], ctx.tempering_stats(4)[0], flush=True)
    ctx.set_ladder(k.propose_ladder(ctx.ladder(4), *ctx.tempering_stats(4)))
    ctx.run(10)
with k.create(room, 50, seed=2, devices=[0, 0, 0]) as ctx:
    ctx.run(15)
    print("multi", ctx.best(), ctx.top_k(3)[0], flush=True)
print("eval", k.eval_costs(room, S.random_layouts(room, 33, 1))["totalCosts"][:3])
