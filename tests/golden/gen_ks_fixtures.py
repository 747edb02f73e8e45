"""Generates tests/golden/ks_oracle_finals.npz: the final resultCosts of the test oracle's chains
(oracle/mh_oracle.c, the C restatement of Kernel.cu:566-713 + 777-827) at the long horizons the KS parity
tests use -- sizes at which the oracle needs minutes of CPU, so its samples are committed instead of being
recomputed on the GPU box:

    config 3 (50 objects)   8 seeds x 4096 chains x 2000 iterations, beta = 2
    config 4 (200 objects)  8 seeds x 4096 chains x 100 iterations,  beta = 2

Seeds were fixed BEFORE any p-value was looked at: the oracle uses seeds
9001..9008, the kernel 101..108; pair i compares kernel seed 101+i with oracle seed 9001+i.  A chain depends
only on (seed, global chain id), so tests/test_oracle.py re-runs a few chains of every sample and checks them
against this file bit for bit: the fixture cannot drift away from the oracle's code.

    python tests/golden/gen_ks_fixtures.py [threads]        (~1 h on 8 cores)
"""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

ORACLE_SEEDS = list(range(9001, 9009))
KERNEL_SEEDS = list(range(101, 109))
PLAN = {3: dict(chains=4096, iterations=2000), 4: dict(chains=4096, iterations=100)}
OUT = os.path.join(HERE, "ks_oracle_finals.npz")


def main():
    from oracle_lib import Oracle
    pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    o = Oracle()
    out = {"oracle_seeds": np.array(ORACLE_SEEDS), "kernel_seeds": np.array(KERNEL_SEEDS)}
    if os.path.exists(OUT):                                     # resume: keep what is already there
        with np.load(OUT) as z:
            out.update({k: z[k] for k in z.files})
    for cid, plan in PLAN.items():
        room = pkg.synth.make_config(cid)
        for seed in ORACLE_SEEDS:
            key = f"cfg{cid}_seed{seed}"
            if key in out:
                continue
            t0 = time.time()
            _, costs = o.run(room, plan["chains"], plan["iterations"], seed=seed, threads=threads)
            out[key] = np.stack([costs[f] for f in pkg.layout.COST_FIELDS], 1).astype(np.float32)
            out[f"cfg{cid}_plan"] = np.array([plan["chains"], plan["iterations"]])
            np.savez_compressed(OUT, **out)
            print(f"{key}: {time.time() - t0:.0f} s, mean total {out[key][:, 0].mean():.3f}", flush=True)


if __name__ == "__main__":
    main()
