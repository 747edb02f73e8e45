"""Randomised cross-checks on the GPU (development probe; the deterministic versions of these checks are in tests/):
for random rooms (make_wild_room: arbitrary quadrilateral surface, rectangles anywhere in the vertex pool, shared
clearance sources, relationship hubs, frozen objects, any-sign weights), random lane widths, chain counts, iteration
splits, result modes, schedules and tempering ladders:
  * the memo form (and the library default) return the plain scan's bytes,
  * a run split over launches, over shards (chain_offset + total_chains), over devices (a repeated ordinal) and over
    one-shot chunks (MH_CHUNKS) returns the unsplit run's bytes,
  * the reported costs are the oracle's cost function of the returned layouts (1e-5 relative, assert_costs_close),
  * KernelBest / KernelTopK agree with numpy.
usage: python tools/fuzz_gpu.py [seconds, default 120] [seed]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from oracle_lib import Oracle
    from test_gpu_parity import assert_costs_close, layouts_from_points, near_jump
    pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
    S = pkg.synth
    k, o = pkg.Kernel(), Oracle()
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    g = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
    t0, rounds, mismatches = time.time(), 0, 0
    while time.time() - t0 < budget:
        n = int(g.integers(1, 90))
        room = S.make_wild_room(n, int(g.integers(0, n + 1)), int(g.integers(0, 70)), int(g.integers(0, 1 << 30)))
        lanes = int(g.choice([1, 2, 4, 8, 16, 32]))
        if (n + lanes - 1) // lanes > 32:
            lanes = 32
        chains = int(g.integers(1, 300))
        iters = int(g.integers(0, 160))
        kw = dict(seed=int(g.integers(0, 1 << 40)), lanes_per_chain=lanes, result_mode=int(g.integers(0, 2)))
        if g.random() < 0.3:
            kw.update(beta_start=float(g.uniform(0.2, 2)), beta_end=float(g.uniform(2, 12)), schedule=int(g.integers(1, 3)), schedule_length=max(iters, 1))
        elif g.random() < 0.3:
            rungs = int(g.choice([2, 3, 4]))
            chains = max(rungs, chains - chains % rungs)
            kw.update(beta_start=0.5, beta_end=8.0, tempering_rungs=rungs, exchange_interval=int(g.integers(5, 40)))
        tag = (n, room.C, room.R, lanes, chains, iters, {a: b for a, b in kw.items() if a != "seed"})
        try:
            base = k.wrapper_ex(room, chains, iters, eval_mode=3, **kw)
        except pkg.KernelError as e:
            assert "shared memory" in str(e), (tag, e)
            continue
        for mode in (0, 2):
            try:
                got = k.wrapper_ex(room, chains, iters, eval_mode=mode, **kw)
            except pkg.KernelError as e:
                assert "shared memory" in str(e), (tag, e)
                continue
            assert got[0].tobytes() == base[0].tobytes() and got[1].tobytes() == base[1].tobytes(), ("memo != scan", mode, tag)
        # split over launches
        with k.create(room, chains, eval_mode=3, **kw) as ctx:
            left = iters
            while left > 0:
                step = int(g.integers(1, left + 1))
                ctx.run(step)
                left -= step
            got = ctx.results()
            if iters > 0:
                i, t = ctx.best()
                assert i == int(np.argmax(got[1]["totalCosts"])) and t == got[1]["totalCosts"][i], ("best", tag)
                kk = int(g.integers(1, chains + 1))
                idx, tot = ctx.top_k(kk)
                order = np.lexsort((np.arange(chains), -got[1]["totalCosts"].astype(np.float64)))
                assert np.array_equal(idx, order[:kk]), ("topk", tag)
        assert got[0].tobytes() == base[0].tobytes() and got[1].tobytes() == base[1].tobytes(), ("split launches", tag)
        # devices (repeated ordinal) and chunks
        unit = kw.get("tempering_rungs", 1)
        os.environ["MH_CHUNKS"] = str(int(g.integers(1, 6)))
        got = k.wrapper_ex(room, chains, iters, eval_mode=3, devices=[0] * int(g.integers(2, 5)), **kw)
        del os.environ["MH_CHUNKS"]
        assert got[0].tobytes() == base[0].tobytes() and got[1].tobytes() == base[1].tobytes(), ("devices/chunks", tag)
        # a shard of the job
        if chains >= 2 * unit and "tempering_rungs" not in kw:
            first = int(g.integers(1, chains))
            a = k.wrapper_ex(room, first, iters, eval_mode=3, total_chains=chains, **kw)
            b = k.wrapper_ex(room, chains - first, iters, eval_mode=3, chain_offset=first, total_chains=chains, **kw)
            assert a[0].tobytes() + b[0].tobytes() == base[0].tobytes(), ("shards", tag)
        # costs against the oracle
        lay = layouts_from_points(room, base[0][:24])
        try:
            assert_costs_close(room, base[1][:24], o.costs_batch(room, lay), skip_pair=near_jump(o, room, lay), rtol=2e-5)
        except AssertionError as e:
            ref = o.costs_batch(room, lay)
            print("COST MISMATCH", tag, e, "weights", [float(room.srf[f][0]) for f in ("WeightSurfaceArea", "WeightClearance", "WeightSymmetry")],
                  "room x", float(room.surfaceRectangle["x"].min()), float(room.surfaceRectangle["x"].max()), flush=True)
            for f in pkg.layout.COST_FIELDS:
                d = np.abs(base[1][:24][f].astype(np.float64) - ref[f]); j = int(np.argmax(d))
                print("   ", f, "worst abs", d[j], "ref", ref[f][j], "got", base[1][f][j])
            mismatches += 1
        rounds += 1
    print(f"fuzz done: {rounds} random rooms in {time.time() - t0:.0f} s, {mismatches} cost mismatches (every byte-identity check passed)")


if __name__ == "__main__":
    main()
