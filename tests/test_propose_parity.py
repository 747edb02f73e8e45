"""propose() / Accept() pinned to the REFERENCE's own device code.

oracle/ref_gpu_harness.cu #includes the unmodified Kernel.cu and exposes its `propose` (Kernel.cu:576-704)
and `Accept` (Kernel.cu:706-713) one thread per item (RefProposeGPU / RefAcceptGPU, XORWOW seeded as
initRNG does, Kernel.cu:159).  The reference's XORWOW bit stream is not reproduced by the new build
(counter-based Philox), so the comparison is distributional:

  three sources of ONE proposal applied to the same layout --
    ref     the reference's propose() on the device,
    oracle  the C restatement (oracle/mh_oracle.c), one iteration at beta = 1e-30 (every proposal accepted),
    kernel  libKernel.so through KernelCreate / KernelRunTraced, same setting,
  compared on: move-type frequencies, object choice (frozen objects never picked), the swap pair law,
  dx / sigma_x, dy / sigma_y, dRot / sigma_t against N(0, 1) and against each other (KS), dx-dy
  independence, the clamp to the room (exact wall values, equal clamp rates), the one-sided rotation
  wrap (range and distribution), what a swap carries (x, y, z, rotX, rotY, rotZ; not length / width /
  frozen), and the acceptance rule u < min(1, exp(2 dE)) (Q10: it maximises).

Seeds and thresholds were fixed before the first run; every statistical check uses p > 1e-3 (about 40
checks in the file).  kernel and oracle share the Philox stream, so they are also compared element-wise.
"""
import importlib
import os

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu

pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
L, S = pkg.layout, pkg.synth
P_MIN = 1e-3
N_PROPOSALS = 240_000
TINY_BETA = 1e-30          # exp(beta dE) == 1.0f: every proposal is accepted (u < 1 except u == 1.0f, p ~ 3e-8)
TWO_PI = 2 * L.PI


def _ref():
    from oracle_lib import RefGPU, ref_gpu_path
    if not os.path.exists(ref_gpu_path()):
        pytest.skip("oracle/_ref/libKernel_ref.so not built")
    return RefGPU()


def _f32(room):
    """Round the layout to float32-representable doubles: the state a chain of the float32 kernel holds,
    and it makes the reference's float temporaries in a swap (quirk Q12) lossless."""
    for f in ("x", "y", "z", "rotX", "rotY", "rotZ"):
        room.cfg[f] = room.cfg[f].astype(np.float32).astype(np.float64)
    return room


def classify(before, after):
    """before: struct array [n]; after: dict of [N, n] float64 arrays x, y, rotY.
    -> move (0 translate, 1 rotate, 2 swap), obj1, obj2 (-1 if none; swap with itself: obj1 = obj2 = -2)."""
    cx = after["x"] != before["x"][None, :]
    cy = after["y"] != before["y"][None, :]
    cr = after["rotY"] != before["rotY"][None, :]
    ch = cx | cy | cr
    cnt = ch.sum(1)
    n = ch.shape[1]
    first = ch.argmax(1)
    last = n - 1 - ch[:, ::-1].argmax(1)
    rows = np.arange(len(cnt))
    move = np.full(len(cnt), -1)
    obj1 = np.full(len(cnt), -1)
    obj2 = np.full(len(cnt), -1)
    one = cnt == 1
    tr = one & (cx | cy)[rows, first] & ~cr[rows, first]
    ro = one & cr[rows, first] & ~(cx | cy)[rows, first]
    sw = cnt == 2
    self_sw = cnt == 0
    move[tr], move[ro], move[sw | self_sw] = 0, 1, 2
    obj1[tr | ro | sw] = first[tr | ro | sw]
    obj2[sw] = last[sw]
    obj1[self_sw] = obj2[self_sw] = -2
    assert (move >= 0).all(), "a proposal changed something no single move changes"
    return move, obj1, obj2


def three_sources(kernel, oracle, room, count, seed):
    """One proposal per layout from the reference, the oracle and the kernel.  Returns
    {name: dict(x, y, rotY, z, rotX, rotZ as [count, n] float64)} and the kernel's trace."""
    ref = _ref()
    n = room.n
    out = {}
    lay = ref.propose_gpu(room, np.tile(room.cfg, count), seed)
    out["ref"] = {f: lay[f].reshape(count, n).astype(np.float64) for f in ("x", "y", "rotY", "z", "rotX", "rotZ")}
    out["ref"]["_extra"] = {f: lay[f].reshape(count, n) for f in ("length", "width", "frozen")}
    po, _ = oracle.run(room, count, 1, seed=seed, beta_start=TINY_BETA)
    out["oracle"] = {f: po[f].astype(np.float64) for f in ("x", "y", "rotY", "z", "rotX", "rotZ")}
    with kernel.create(room, count, seed=seed, beta_start=TINY_BETA) as ctx:
        tr = ctx.run_traced(1)[0]
        pk, _ = ctx.results()
    out["kernel"] = {f: pk[f].astype(np.float64) for f in ("x", "y", "rotY", "z", "rotX", "rotZ")}
    assert tr["accepted"].mean() > 0.99999
    return out, tr


def chi2_p(counts, probs):
    counts = np.asarray(counts, np.float64)
    return stats.chisquare(counts, counts.sum() * np.asarray(probs, np.float64)).pvalue


def test_move_type_object_choice_and_swap_pairs(kernel, oracle):
    """chi^2 on the move type, on the object a translate / rotate picks, and on the unordered pair a
    swap picks, for all three sources; frozen objects are never touched (Kernel.cu:601, 637, 662, 666)."""
    room = _f32(S.make_config(2))
    frozen = [3, 8, 15]
    room.cfg["frozen"][frozen] = 1
    free = np.array([i for i in range(room.n) if i not in frozen])
    f = len(free)
    src, tr = three_sources(kernel, oracle, room, N_PROPOSALS, seed=1001)
    tables = {}
    for name, a in src.items():
        move, o1, o2 = classify(room.cfg, a)
        # the first draw: trunc(u * 2.999999) (Kernel.cu:566-574, 583)
        mc = np.bincount(move, minlength=3)
        assert chi2_p(mc, [1 / 2.999999, 1 / 2.999999, 0.999999 / 2.999999]) > P_MIN, (name, mc)
        for m in (0, 1):
            oc = np.bincount(o1[move == m], minlength=room.n)
            assert oc[frozen].sum() == 0, name
            assert chi2_p(oc[free], np.full(f, 1 / f)) > P_MIN, (name, m, oc)
        sw = move == 2
        self_count = int((o1[sw] == -2).sum())
        pair = np.zeros((room.n, room.n), np.int64)
        np.add.at(pair, (o1[sw & (o1 >= 0)], o2[sw & (o1 >= 0)]), 1)
        assert pair[frozen, :].sum() == 0 and pair[:, frozen].sum() == 0, name
        iu = np.triu_indices(f, 1)
        pc = pair[np.ix_(free, free)][iu]
        # obj1, obj2 independent and uniform over the free objects: P(self) = 1/f, P({a, b}) = 2/f^2
        assert chi2_p(np.concatenate([[self_count], pc]), np.concatenate([[1 / f], np.full(len(pc), 2 / f ** 2)])) > P_MIN, name
        tables[name] = np.concatenate([mc, np.bincount(o1[move == 0], minlength=room.n)[free],
                                       np.bincount(o1[move == 1], minlength=room.n)[free], [self_count], pc])
        if name == "kernel":                                   # the classification agrees with the kernel's own trace
            assert np.array_equal(move, tr["move"])
            one = move < 2
            assert np.array_equal(o1[one], tr["obj1"][one])
            two = (move == 2) & (o1 >= 0)
            assert np.array_equal(np.minimum(tr["obj1"][two], tr["obj2"][two]), o1[two])
            assert np.array_equal(np.maximum(tr["obj1"][two], tr["obj2"][two]), o2[two])
            assert np.array_equal(tr["obj1"][o1 == -2], tr["obj2"][o1 == -2])
    # two-sample: the reference's table against the kernel's and the oracle's
    for other in ("kernel", "oracle"):
        p = stats.chi2_contingency(np.stack([tables["ref"], tables[other]]))[1]
        assert p > P_MIN, (other, p)
    # same Philox stream: kernel and oracle make the same proposals
    for fld in ("x", "y", "rotY"):
        assert np.allclose(src["kernel"][fld], src["oracle"][fld], rtol=0, atol=2e-6)


def _wrapped(d):
    return (d + L.PI) % TWO_PI - L.PI


def test_translate_and_rotate_deltas_are_the_references(kernel, oracle):
    """dx / (W/16), dy / (H/16) (quirk Q19) and dRot / (15/90 * 3.1416) are N(0, 1) draws in the reference;
    the new build's must be too, and indistinguishable from the reference's (two-sample KS).  A wide room
    with every object near its centre keeps the clamp out of this test (walls at >= 6 sigma)."""
    room = S.make_room(12, 4, 6, 48.0, 32.0, 4242)
    g = np.random.default_rng(5)
    room.cfg["x"] = 24.0 + g.uniform(-4, 4, room.n)
    room.cfg["y"] = 16.0 + g.uniform(-3, 3, room.n)
    _f32(room)
    sx, sy, st = np.float32(48.0) / 16, np.float32(32.0) / 16, np.float32(15.0 / 90.0 * L.PI)
    src, _ = three_sources(kernel, oracle, room, N_PROPOSALS, seed=2002)
    z = {}
    for name, a in src.items():
        move, o1, _ = classify(room.cfg, a)
        rows = np.nonzero(move == 0)[0]
        dx = (a["x"][rows, o1[rows]] - room.cfg["x"][o1[rows]]) / sx
        dy = (a["y"][rows, o1[rows]] - room.cfg["y"][o1[rows]]) / sy
        rows = np.nonzero(move == 1)[0]
        dr = _wrapped(a["rotY"][rows, o1[rows]] - room.cfg["rotY"][o1[rows]]) / st
        z[name] = (dx, dy, dr)
        for what, v in zip(("dx", "dy", "dRot"), (dx, dy, dr)):
            assert len(v) > N_PROPOSALS / 4
            assert stats.kstest(v, "norm").pvalue > P_MIN, (name, what, v.mean(), v.std())
        r = np.corrcoef(dx, dy)[0, 1]                           # curand_normal twice: independent
        assert abs(r) < 4.5 / np.sqrt(len(dx)), (name, r)
    for other in ("kernel", "oracle"):
        for i, what in enumerate(("dx", "dy", "dRot")):
            assert stats.ks_2samp(z["ref"][i], z[other][i]).pvalue > P_MIN, (other, what)


def test_clamp_to_the_room_and_rotation_wrap(kernel, oracle):
    """Kernel.cu:616-633: a translate that leaves the room snaps to the wall (exactly), each axis on its own;
    Kernel.cu:649-652: rotY is wrapped ONCE into [0, 2 * 3.1416].  Objects are parked next to the walls and
    next to both ends of the rotation range so that a good share of the moves triggers them."""
    room = S.make_config(1)                                     # 4 x 4 room: sigma = 0.25
    room.cfg["x"] = [0.05, 3.95, 0.30, 3.70, 2.0, 0.10, 3.90, 2.0]
    room.cfg["y"] = [2.0, 0.08, 3.92, 0.25, 3.75, 3.97, 0.03, 2.0]
    room.cfg["rotY"] = [0.05, 0.3, 6.25, 6.0, 3.0, 0.6, 5.7, 6.2831]
    _f32(room)
    W = H = 4.0
    src, _ = three_sources(kernel, oracle, room, N_PROPOSALS, seed=3003)
    rates, rot_after = {}, {}
    for name, a in src.items():
        move, o1, _ = classify(room.cfg, a)
        assert a["x"].min() >= 0.0 and a["x"].max() <= W and a["y"].min() >= 0.0 and a["y"].max() <= H, name
        assert a["rotY"].min() >= 0.0 and a["rotY"].max() <= TWO_PI + 1e-6, name
        rows = np.nonzero(move == 0)[0]
        nx, ny = a["x"][rows, o1[rows]], a["y"][rows, o1[rows]]
        # per object: how often each wall was hit (exact wall value), x and y separately
        r = []
        for obj in range(room.n):
            m = o1[rows] == obj
            r += [np.sum(nx[m] == 0.0), np.sum(nx[m] == W), np.sum(ny[m] == 0.0), np.sum(ny[m] == H), int(m.sum())]
        rates[name] = np.array(r, np.float64).reshape(room.n, 5)
        rows = np.nonzero(move == 1)[0]
        rot_after[name] = [a["rotY"][rows[o1[rows] == obj], obj] for obj in range(room.n)]
    # the clamp rate of the reference is Phi(-d / sigma) per wall; compare all three with it and with each other
    sig = 0.25
    for obj in range(room.n):
        x, y = room.cfg["x"][obj], room.cfg["y"][obj]
        expect = [stats.norm.cdf(-x / sig), stats.norm.cdf(-(W - x) / sig), stats.norm.cdf(-y / sig), stats.norm.cdf(-(H - y) / sig)]
        for name in src:
            tot = rates[name][obj, 4]
            for w in range(4):
                k, p = rates[name][obj, w], expect[w]
                if p * tot < 5:
                    assert k <= max(12, 10 * p * tot), (name, obj, w, k)
                else:
                    assert abs(k - p * tot) < 4.5 * np.sqrt(p * (1 - p) * tot), (name, obj, w, k, p * tot)
    assert rates["ref"][:, :4].sum() > 0.1 * rates["ref"][:, 4].sum()         # the test does exercise the clamp
    wrapped = 0
    for obj in range(room.n):
        for other in ("kernel", "oracle"):
            assert stats.ks_2samp(rot_after["ref"][obj], rot_after[other][obj]).pvalue > P_MIN, (other, obj)
        wrapped += np.sum(np.abs(rot_after["ref"][obj] - room.cfg["rotY"][obj]) > 3.0)
    assert wrapped > 0.05 * sum(len(v) for v in rot_after["ref"])               # ... and the wrap


def test_swap_carries_position_and_all_three_rotations(kernel, oracle):
    """Kernel.cu:675-700: a swap exchanges x, y, z, rotX, rotY, rotZ and leaves length, width, frozen."""
    room = S.make_config(1)
    room.cfg["z"] = 0.5 + np.arange(8)
    room.cfg["rotX"] = 0.25 * (1 + np.arange(8))
    room.cfg["rotZ"] = -0.125 * (1 + np.arange(8))
    _f32(room)
    src, tr = three_sources(kernel, oracle, room, 20000, seed=4004)
    for name, a in src.items():
        move, o1, o2 = classify(room.cfg, a)
        rows = np.nonzero((move == 2) & (o1 >= 0))[0]
        assert len(rows) > 4000
        for fld in ("x", "y", "z", "rotX", "rotY", "rotZ"):
            b = room.cfg[fld]
            assert np.array_equal(a[fld][rows, o1[rows]], b[o2[rows]]), (name, fld)
            assert np.array_equal(a[fld][rows, o2[rows]], b[o1[rows]]), (name, fld)
        other = np.nonzero(move != 2)[0]
        for fld in ("z", "rotX", "rotZ"):                       # translate / rotate leave them alone
            assert np.array_equal(a[fld][other], np.tile(room.cfg[fld], (len(other), 1))), (name, fld)
    ex = src["ref"]["_extra"]
    for fld in ("length", "width", "frozen"):
        assert np.array_equal(ex[fld], np.tile(room.cfg[fld], (20000, 1))), fld


def test_accept_rule_of_the_reference_and_of_the_kernel(kernel, oracle):
    """Kernel.cu:706-713: accept iff u < min(1, (float)exp(BETA (costStar - costCur))), BETA = 2 -- a rule that
    MAXIMISES totalCosts (quirk Q10).  (i) the reference's own Accept() on a grid of dE: acceptance frequency
    against min(1, exp(2 dE)); (ii) every decision in a kernel trace and in an oracle trace reproduces the rule
    from the trace's own (u, star, previous cur); (iii) kernel acceptance frequencies per dE bin agree with
    the rule's expectation."""
    ref = _ref()
    per = 60000
    grid = np.array([-4.0, -2.0, -1.0, -0.5, -0.2, -0.05, -0.005, 0.0, 0.01, 0.7, 5.0])
    star = np.repeat(1234.5 + grid, per)
    cur = np.full(len(star), 1234.5)
    acc = ref.accept_gpu(star, cur, seed=5005).reshape(len(grid), per)
    for d, a in zip(grid, acc):
        p = min(1.0, float(np.exp(2.0 * d)))
        k = a.sum()
        if p >= 1.0:
            assert k >= per - 1, (d, k)                          # u == 1.0f is possible (curand_uniform is (0, 1])
        else:
            assert abs(k - p * per) < 4.5 * np.sqrt(p * (1 - p) * per), (d, k, p * per)

    room = S.make_config(2)
    c0 = float(oracle.costs(room)["totalCosts"])
    with kernel.create(room, 512, seed=6006) as ctx:
        tk = ctx.run_traced(400)
    _, _, to = oracle.run(room, 512, 400, seed=6006, trace=True)
    for name, t in (("kernel", tk), ("oracle", to)):
        # from iteration 1 on, the total the chain held before the decision is the previous entry's cur_total
        prev = t["cur_total"][:-1].astype(np.float64)
        star, u, beta = t["star_total"][1:].astype(np.float64), t["u"][1:], t["beta"][1:].astype(np.float64)
        acc = t["accepted"][1:].astype(bool)
        thr = np.minimum(np.float32(1.0), np.exp(np.minimum(beta * (star - prev), 80.0)).astype(np.float32))
        want = u < thr
        edge = np.abs(u.astype(np.float64) - thr) <= 1e-6 * np.maximum(thr, 1e-30)   # exp() of two libms may round apart
        bad = (want != acc) & ~edge
        assert not bad.any(), (name, int(bad.sum()))
        assert edge.mean() < 1e-4
        assert np.all(t["cur_total"][t["accepted"] == 1] == t["star_total"][t["accepted"] == 1])
        assert np.all(t["cur_total"][1:][~acc] == t["cur_total"][:-1][~acc])
        assert np.all(t["beta"] == np.float32(2.0))
    # frequencies per dE bin (kernel): observed acceptance against the mean of min(1, exp(2 dE)) in the bin
    prev = np.vstack([np.full((1, 512), np.float32(c0)), tk["cur_total"][:-1]])
    dE = (tk["star_total"].astype(np.float64) - prev)[1:].ravel()
    a = tk["accepted"][1:].ravel()
    for lo, hi in ((-3.0, -1.0), (-1.0, -0.3), (-0.3, -0.05), (-0.05, 0.0)):
        m = (dE >= lo) & (dE < hi)
        if m.sum() < 500:
            continue
        p = np.minimum(1.0, np.exp(2.0 * dE[m]))
        assert abs(a[m].sum() - p.sum()) < 4.5 * np.sqrt((p * (1 - p)).sum() + 1.0), (lo, hi, a[m].sum(), p.sum())
    assert np.all(a[dE > 1e-3] == 1)                             # uphill is always taken: the rule maximises
