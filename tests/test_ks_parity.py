"""Statistical parity of the chains: final-cost distributions of the CUDA path against the test oracle
(SURVEY.md section 8d "statistical parity"; BASELINE section 5: two-sample KS, p > 0.01, >= 4096 chains).

Protocol (fixed before any p-value was looked at; VERDICT round 1 asked for exactly this instead of single seed
pairs picked after the fact):
  * 8 seed pairs per case: kernel seeds 101..108 against oracle seeds 9001..9008, disjoint streams;
  * one two-sample KS per pair on the final totalCosts of 4096 chains (2048 for the tempering case);
  * the case passes when the FISHER-COMBINED p of the 8 pairs is > 0.01; for the per-term checks of the
    headline room (7 weighted terms) the threshold is Bonferroni-split, 0.01 / 7;
  * no pair is dropped, no seed is changed: a failure is a finding.
Long horizons (config 3: 2000 iterations, config 4: 100 iterations, 4096 chains each) use the oracle samples
committed in tests/golden/ks_oracle_finals.npz (gen_ks_fixtures.py; tests/test_oracle.py pins that file to the
oracle's code); the short cases run the oracle live.

The kernel is float32 and the reference mixes float and double; the same seed therefore gives the same chain
until an accept decision lands within rounding of its threshold (the trajectory tests), and DIFFERENT seeds give
independent samples of -- if the port is right -- the same distribution (this file)."""
import importlib
import os

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu

pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
L, S = pkg.layout, pkg.synth
HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_SEEDS = list(range(101, 109))
ORACLE_SEEDS = list(range(9001, 9009))
P_MIN = 0.01
TERMS = ("PairWiseCosts", "VisualBalanceCosts", "FocalPointCosts", "SymmetryCosts", "ClearanceCosts", "OffLimitsCosts", "SurfaceAreaCosts")


def fisher(ps):
    ps = np.maximum(np.asarray(ps, np.float64), 1e-300)
    return float(stats.chi2.sf(-2.0 * np.log(ps).sum(), 2 * len(ps)))


def ks_pairs(kernel_fn, oracle_samples, field="totalCosts"):
    """p-values of the 8 seed pairs; kernel_fn(seed) -> resultCosts array, oracle_samples[i] = the oracle's."""
    return [stats.ks_2samp(kernel_fn(sk)[field], co[field]).pvalue for sk, co in zip(KERNEL_SEEDS, oracle_samples)]


def fixture():
    path = os.path.join(HERE, "golden", "ks_oracle_finals.npz")
    z = np.load(path)
    assert list(z["kernel_seeds"]) == KERNEL_SEEDS and list(z["oracle_seeds"]) == ORACLE_SEEDS
    return z


@pytest.mark.parametrize("cid,iters", [(1, 400), (2, 300)])
def test_small_rooms_live_oracle(kernel, oracle, cid, iters):
    room = S.make_config(cid)
    ref = [oracle.run(room, 4096, iters, seed=s)[1] for s in ORACLE_SEEDS]
    ps = ks_pairs(lambda s: kernel.wrapper_ex(room, 4096, iters, seed=s)[1], ref)
    assert fisher(ps) > P_MIN, ps
    pd = ks_pairs(lambda s: kernel.wrapper_ex(room, 4096, iters, seed=s, eval_mode=1)[1], ref)
    assert fisher(pd) > P_MIN, pd                               # delta evaluation: statistically equivalent
    # and the SAME seed gives near-identical populations (most chains never meet a tie)
    _, ck = kernel.wrapper_ex(room, 512, iters, seed=KERNEL_SEEDS[0])
    _, co = oracle.run(room, 512, iters, seed=KERNEL_SEEDS[0])
    assert np.isclose(ck["totalCosts"], co["totalCosts"], rtol=1e-4, atol=1e-3).mean() > 0.8


def test_config3_long_horizon_against_committed_oracle_samples(kernel):
    """The headline room, 50 objects, all terms: 4096 chains x 2000 iterations, default evaluation (the memo
    form) and delta evaluation, total and every weighted term."""
    z = fixture()
    chains, iters = (int(v) for v in z["cfg3_plan"])
    assert chains >= 4096 and iters >= 2000
    room = S.make_config(3)
    runs = {mode: [kernel.wrapper_ex(room, chains, iters, seed=s, eval_mode=mode)[1] for s in KERNEL_SEEDS] for mode in (0, 1)}
    for mode, cs in runs.items():
        ps = [stats.ks_2samp(c["totalCosts"], z[f"cfg3_seed{so}"][:, 0]).pvalue for c, so in zip(cs, ORACLE_SEEDS)]
        assert fisher(ps) > P_MIN, (mode, ps)
    for f in TERMS:
        j = L.COST_FIELDS.index(f)
        ps = [stats.ks_2samp(c[f], z[f"cfg3_seed{so}"][:, j]).pvalue for c, so in zip(runs[0], ORACLE_SEEDS)]
        assert fisher(ps) > P_MIN / len(TERMS), (f, ps)
    # the sampler climbs (it maximises totalCosts, quirk Q10), by as much as the oracle's chains do
    mk = np.mean([c["totalCosts"].mean() for c in runs[0]])
    mo = np.mean([z[f"cfg3_seed{so}"][:, 0].mean() for so in ORACLE_SEEDS])
    sd = np.mean([z[f"cfg3_seed{so}"][:, 0].std() for so in ORACLE_SEEDS])
    assert abs(mk - mo) < 0.05 * sd, (mk, mo, sd)


def test_config4_against_committed_oracle_samples(kernel):
    """200 objects (the default runs the memo form, 32 lanes per chain): 4096 chains x 100
    iterations, default and delta evaluation."""
    z = fixture()
    chains, iters = (int(v) for v in z["cfg4_plan"])
    assert chains >= 4096
    room = S.make_config(4)
    for mode in (0, 1):
        ps = [stats.ks_2samp(kernel.wrapper_ex(room, chains, iters, seed=sk, eval_mode=mode)[1]["totalCosts"], z[f"cfg4_seed{so}"][:, 0]).pvalue
              for sk, so in zip(KERNEL_SEEDS, ORACLE_SEEDS)]
        assert fisher(ps) > P_MIN, (mode, ps)


def test_annealing_schedule(kernel, oracle):
    """Extension without a reference counterpart (geometric beta 0.5 -> 16 over the run): kernel and oracle
    implement the same spec independently."""
    room = S.make_config(2)
    kw = dict(beta_start=0.5, beta_end=16.0, schedule=1)
    ref = [oracle.run(room, 4096, 600, seed=s, **kw)[1] for s in ORACLE_SEEDS]
    for mode in (0, 1):
        ps = ks_pairs(lambda s: kernel.wrapper_ex(room, 4096, 600, seed=s, eval_mode=mode, **kw)[1], ref)
        assert fisher(ps) > P_MIN, (mode, ps)


def test_parallel_tempering(kernel, oracle):
    room = S.make_config(1)
    kw = dict(beta_start=0.5, beta_end=8.0, tempering_rungs=4, exchange_interval=20)
    ref = [oracle.run(room, 2048, 400, seed=s, **kw)[1] for s in ORACLE_SEEDS]
    ps = ks_pairs(lambda s: kernel.wrapper_ex(room, 2048, 400, seed=s, **kw)[1], ref)
    assert fisher(ps) > P_MIN, ps
