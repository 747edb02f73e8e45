"""Worst and 99.9th-percentile RELATIVE error of every cost term, KernelEvalCosts (the float32 CUDA path) against
the test oracle (pinned bit-for-bit to the reference's own cost code), per BASELINE room, over >= 1e5 layouts:
half of them uniform over the room grown by 25 %, half of them layouts the sampler itself visits (final states
of chains after 0..400 iterations, i.e. the piled-up rooms quirk Q10 produces).

Writes profiles/parity_errors_r2.json.  north_star's bar is 1e-5 relative per term; this table is the evidence
for (and the limits of) tests/test_gpu_parity.py::assert_costs_close.

    python tools/parity_errors.py [layouts per room, default 100000]
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from oracle_lib import Oracle
    from test_gpu_parity import layouts_from_points, near_jump, term_scales
    pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
    L, S = pkg.layout, pkg.synth
    k, o = pkg.Kernel(), Oracle()
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    out = {"what": "relative error |kernel - oracle| / |oracle| of each weighted cost term where |oracle| > 0.1 x the term's natural scale "
                   "(tests/test_gpu_parity.py: term_scales); absolute error / scale below that",
           "layouts_per_room": count, "device": k.device_info()["name"], "rooms": {}}
    for cid in (1, 2, 3, 4):
        room = S.make_config(cid)
        half = count // 2
        lay = [S.random_layouts(room, half, 4000 + cid)]
        per = max(1, (count - half) // 5)
        for j, iters in enumerate((0, 25, 100, 200, 400)):
            if cid == 4:
                iters = iters // 4
            pts, _ = k.wrapper_ex(room, per, iters, seed=50 + j)
            lay.append(layouts_from_points(room, pts))
        lay = _cat(lay, L)
        got = k.eval_costs(room, lay)
        ref = o.costs_batch(room, lay)
        skip = near_jump(o, room, lay)
        sc = term_scales(room, ref)
        terms = {}
        for f in L.COST_FIELDS:
            g, r = got[f].astype(np.float64), ref[f].astype(np.float64)
            if f in ("PairWiseCosts", "totalCosts"):
                g, r = g[~skip], r[~skip]
            big = np.abs(r) > 0.1 * sc[f]
            rel = np.abs(g[big] - r[big]) / np.abs(r[big]) if big.any() else np.zeros(0)
            small = np.abs(g[~big] - r[~big]) / sc[f] if (~big).any() else np.zeros(0)
            terms[f] = {"scale": float(sc[f]), "n_rel": int(big.sum()), "rel_worst": float(rel.max()) if len(rel) else None,
                        "rel_p999": float(np.quantile(rel, 0.999)) if len(rel) else None,
                        "rel_median": float(np.median(rel)) if len(rel) else None,
                        "n_small": int((~big).sum()), "abs_over_scale_worst": float(small.max()) if len(small) else None,
                        "abs_over_scale_p999": float(np.quantile(small, 0.999)) if len(small) else None}
        out["rooms"][f"config{cid}"] = {"n": room.n, "C": room.C, "R": room.R, "layouts": int(len(lay) // room.n),
                                       "skipped_near_an_angle_jump": int(skip.sum()), "terms": terms}
        print(cid, {f: (t["rel_worst"], t["abs_over_scale_worst"]) for f, t in terms.items()}, flush=True)
    path = os.path.join(ROOT, "profiles", "parity_errors_r2.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


def _cat(parts, L):
    total = sum(len(p) for p in parts)
    lay = np.empty(total, L.positionAndRotation)                # (np.concatenate would repack the padded struct)
    at = 0
    for p in parts:
        lay[at:at + len(p)] = p
        at += len(p)
    return lay


if __name__ == "__main__":
    main()
