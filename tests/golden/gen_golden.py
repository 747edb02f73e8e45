"""Generates tests/golden/costs_golden.json from the REFERENCE's own cost functions.

Run in the build container (needs /root/reference): `python tests/golden/gen_golden.py`.
It builds oracle/_ref/libref_costs_host.so (the reference's Kernel.cu:162-550 compiled as host
C++ from the reference tree, see oracle/Makefile) and records, bit for bit (uint32 hex of each
float), the eight resultCosts fields and the raw terms it returns for:
  * the reference's smoke fixture main() (Kernel.cu:1007-1166) at its initial layout,
  * the initial layout of BASELINE.json configs 1-4,
  * 24 seeded random layouts of configs 1-3 (synth.random_layouts, f32=False and f32=True).
The inputs are regenerated from seeds by synth.py; a sha256 of every input array is stored so
that a drift of the generator is detected rather than silently re-baselined.
"""
import hashlib
import importlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
S = importlib.import_module("metropolis-hastings-gpgpu_b200.synth")
from oracle_lib import RefHost  # noqa: E402


def room_hash(room, cfg=None):
    h = hashlib.sha256()
    for a in (room.srf, room.rss, room.rsa, room.cfg if cfg is None else cfg, room.clearances, room.offlimits, room.vertices,
              room.surfaceRectangle):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def bits(c):
    return [f"{int(x):08x}" for x in np.frombuffer(c.tobytes(), np.uint32)]


def main():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"], check=True)
    ref = RefHost()
    cases = []

    def add(name, room, cfg, gen):
        c, raw = ref.costs(room, cfg, raw=True)
        cases.append({"name": name, "gen": gen, "sha256": room_hash(room, cfg), "costs_bits": bits(c),
                      "costs": [float(x) for x in np.frombuffer(c.tobytes(), np.float32)], "raw": [float.hex(float(x)) for x in raw]})

    fx = S.reference_main_fixture()
    add("reference_main_fixture", fx, fx.cfg, {"kind": "main"})
    for cid in (1, 2, 3, 4):
        room = S.make_config(cid)
        add(f"config{cid}_initial", room, room.cfg, {"kind": "config", "config": cid})
    for cid in (1, 2, 3):
        room = S.make_config(cid)
        for f32 in (False, True):
            lays = S.random_layouts(room, 12, 1000 + cid, f32=f32)
            for l in range(12):
                add(f"config{cid}_random{l}_{'f32' if f32 else 'f64'}", room, lays[l * room.n:(l + 1) * room.n],
                    {"kind": "random", "config": cid, "seed": 1000 + cid, "count": 12, "index": l, "f32": f32})
    out = {"generator": "tests/golden/gen_golden.py", "source": "reference Kernel.cu:162-550 compiled as host C++ (g++ -O2 -ffp-contract=off, glibc libm)",
           "cases": cases}
    with open(os.path.join(HERE, "costs_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(cases)} cases")


if __name__ == "__main__":
    main()
