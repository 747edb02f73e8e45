"""Development probe: the p-values behind the KS parity tests, to check their margins."""
import importlib, os, sys
import numpy as np
from scipy import stats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200"); S = pkg.synth
from oracle_lib import Oracle
k, o = pkg.Kernel(), Oracle()
for cid, iters in ((1, 400), (2, 300)):
    room = S.make_config(cid)
    _, ck = k.wrapper_ex(room, 4096, iters, seed=2024); _, co = o.run(room, 4096, iters, seed=4048)
    print("full cfg", cid, stats.ks_2samp(ck["totalCosts"], co["totalCosts"]).pvalue)
    _, ck = k.wrapper_ex(room, 4096, iters, seed=2025, eval_mode=1); _, co = o.run(room, 4096, iters, seed=5050)
    print("delta cfg", cid, stats.ks_2samp(ck["totalCosts"], co["totalCosts"]).pvalue)
room = S.make_config(3)
_, ck = k.wrapper_ex(room, 4096, 250, seed=777); _, co = o.run(room, 4096, 250, seed=1555)
for f in ("totalCosts", "SymmetryCosts", "ClearanceCosts", "PairWiseCosts", "FocalPointCosts", "SurfaceAreaCosts"):
    print("cfg3", f, stats.ks_2samp(ck[f], co[f]).pvalue)
room = S.make_config(2)
_, ca = k.wrapper_ex(room, 1024, 600, seed=9, beta_start=0.5, beta_end=16.0, schedule=1); _, oa = o.run(room, 1024, 600, seed=10, beta_start=0.5, beta_end=16.0, schedule=1)
print("anneal", stats.ks_2samp(ca["totalCosts"], oa["totalCosts"]).pvalue)
_, ca = k.wrapper_ex(room, 1024, 600, seed=9, eval_mode=1, beta_start=0.5, beta_end=16.0, schedule=1)
print("anneal delta", stats.ks_2samp(ca["totalCosts"], oa["totalCosts"]).pvalue)
room = S.make_config(1)
opts = dict(beta_start=0.5, beta_end=8.0, tempering_rungs=4, exchange_interval=20)
_, ck = k.wrapper_ex(room, 2048, 400, seed=13, **opts); _, co = o.run(room, 2048, 400, seed=14, **opts)
print("tempering", stats.ks_2samp(ck["totalCosts"], co["totalCosts"]).pvalue)
room = S.make_config(3)
_, ck = k.wrapper_ex(room, 4096, 260, seed=77, eval_mode=1); _, co = o.run(room, 4096, 260, seed=7070)
print("delta cfg3", stats.ks_2samp(ck["totalCosts"], co["totalCosts"]).pvalue)
room = S.make_config(2)
_, ca = k.wrapper_ex(room, 4096, 600, seed=3, eval_mode=1, beta_start=0.5, beta_end=16.0, schedule=1)
_, oa = o.run(room, 4096, 600, seed=4, beta_start=0.5, beta_end=16.0, schedule=1)
print("delta anneal (test seeds)", stats.ks_2samp(ca["totalCosts"], oa["totalCosts"]).pvalue)
room = S.make_config(4)
_, co = o.run(room, 1024, 100, seed=31337)
_, ck = k.wrapper_ex(room, 1024, 100, seed=4242); _, cd = k.wrapper_ex(room, 1024, 100, seed=4343, eval_mode=1)
print("cfg4 default", stats.ks_2samp(ck["totalCosts"], co["totalCosts"]).pvalue, "delta", stats.ks_2samp(cd["totalCosts"], co["totalCosts"]).pvalue)
