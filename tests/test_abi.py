"""CPU tests of the drop-in boundary: the library loads, exports every symbol that
include/mh_kernel.h declares, the headers compile as C and C++, the numpy mirrors agree with
the C layout, and -- without a GPU -- the compute entry points fail loudly instead of falling
back to anything."""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
L, S = pkg.layout, pkg.synth


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_numpy_mirrors_match_header_sizes():
    assert L.check_layout()
    assert L.mhOptions.itemsize == 136 and L.mhTraceEntry.itemsize == 32


def test_headers_compile_as_c_and_cpp(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mh_kernel.h"\nint main(void){return (int)sizeof(mhOptions) - 136;}\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        exe = tmp_path / ("t_" + cc)
        subprocess.run([cc, std, "-x", "c" if cc == "gcc" else "c++", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       check=True)
        assert subprocess.run([str(exe)]).returncode == 0


def test_library_exports_every_declared_symbol():
    path = pkg.lib_path()
    assert os.path.exists(path), "libKernel.so not built: run __graft_entry__.build()"
    hdr = open(os.path.join(ROOT, "include", "mh_kernel.h")).read()
    declared = set(re.findall(r"MH_API\s+[\w\s\*]+?\b(Kernel\w+)\s*\(", hdr))
    assert "KernelWrapper" in declared and len(declared) >= 16
    assert declared == set(pkg.binding.EXPORTS)
    lib = C.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == declared, exported ^ declared        # nothing else leaks out of the library


def test_every_export_has_a_ctypes_signature():
    """A context handle is a 64-bit pointer; without argtypes ctypes passes a bare Python int as a 32-bit C int
    and a handle above 4 GiB is truncated.  Every export must declare its signature."""
    k = pkg.Kernel()
    for name in pkg.binding.EXPORTS:
        fn = getattr(k.lib, name)
        assert fn.argtypes is not None, name
    for name in ("KernelWrapper", "KernelWrapperEx", "KernelCreate"):
        assert getattr(k.lib, name).restype is C.c_void_p, name
    # a handle with the high bits set must arrive whole: KernelShape(NULL-ish bogus) is never called, but the
    # marshalling of a 64-bit value through argtypes can be checked without a device
    big = 0x7F12_3456_789A
    assert C.c_void_p(big).value == big


def test_options_struct_grows_compatibly():
    """mhOptions carries struct_size: the round-1 layout (88 bytes) is a prefix of the current one, field for field."""
    r1 = ["struct_size", "flags", "seed", "chain_offset", "iteration_offset", "beta_start", "beta_end", "schedule", "schedule_length",
          "result_mode", "eval_mode", "lanes_per_chain", "device", "tempering_rungs", "exchange_interval", "chain_stride"]
    assert list(L.mhOptions.names[:len(r1)]) == r1
    assert L.mhOptions.fields["chain_stride"][1] == 80 and L.mhOptions.fields["total_chains"][1] == 88
    assert L.mhOptions.fields["n_devices"][1] == 96 and L.mhOptions.fields["devices"][1] == 100
    o = pkg.binding.make_options(devices=[2, 0, 1], device=0)
    assert int(o["n_devices"][0]) == 3 and list(o["devices"][0][:3]) == [2, 0, 1] and int(o["flags"][0]) & L.MH_OPT_EXPLICIT_DEVICE
    assert int(pkg.binding.make_options()["device"][0]) == -1


def test_ladder_policy_is_a_pure_host_function():
    """KernelTemperingProposeLadder needs no device: end points stay, the ladder stays monotone, equal exchange rates
    leave it alone, rungs move towards the gaps that exchange least, and on a synthetic model of the exchange rate
    (acceptance = exp(-c(beta) * gap in log beta), harder towards the cold end) iterating it equalises the rates."""
    k = pkg.Kernel()
    cur = 0.25 * 32.0 ** (np.arange(8) / 7)
    att = np.full(7, 1000)
    same = k.propose_ladder(cur, att, np.full(7, 400))
    assert np.allclose(same, cur, rtol=1e-12)
    new = k.propose_ladder(cur, att, np.array([900, 800, 600, 300, 100, 20, 5]))
    assert new[0] == cur[0] and new[-1] == cur[-1] and np.all(np.diff(new) > 0)
    assert np.all(new[1:-1] > cur[1:-1])                        # the cold end exchanges least: rungs move there
    half = k.propose_ladder(cur, att, np.array([900, 800, 600, 300, 100, 20, 5]), damping=0.5)
    assert np.allclose(np.log(half), 0.5 * (np.log(cur) + np.log(new)))
    few = k.propose_ladder(cur, np.full(7, 3), np.zeros(7, np.int64))    # too few attempts to say anything
    assert np.allclose(few, cur, rtol=1e-12)

    def rates(lad):
        mid = np.sqrt(lad[1:] * lad[:-1])
        return np.exp(-(0.3 + 0.5 * mid) * np.diff(np.log(lad)))
    lad = cur.copy()
    for _ in range(30):
        r = rates(lad)
        lad = k.propose_ladder(lad, np.full(7, 100000), np.round(r * 100000).astype(np.int64), damping=0.7)
    r = rates(lad)
    assert r.max() / r.min() < 1.15 and rates(cur).max() / rates(cur).min() > 3
    with pytest.raises(pkg.KernelError):
        k.propose_ladder([1.0, -2.0], [1], [1])


def test_oracle_is_not_linked_into_the_product():
    """Nothing under oracle/ may be included, linked or loaded by the product (comments may say the word)."""
    out = subprocess.run(["nm", "-D", pkg.lib_path()], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in out
    out = subprocess.run(["ldd", pkg.lib_path()], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in out
    pdir = os.path.join(ROOT, "metropolis-hastings-gpgpu_b200")
    for d, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".c", ".cu", ".cuh", ".h", ".py")) or f == "Makefile":
                text = open(os.path.join(d, f)).read()
                for needle in ("oracle/", "mh_oracle", "liboracle", "oracle_lib", "ref_runner", "_ref/"):
                    assert needle not in text, (f, needle)


def test_bad_arguments_are_rejected_before_any_device_work():
    k = pkg.Kernel()
    room = S.make_config(1)
    room.srf["nClearances"] = room.n + 1                       # quirk Q7 needs C <= n
    with pytest.raises(pkg.KernelError, match="nClearances"):
        k.wrapper_ex(room, 1, 1, seed=1)
    room = S.make_config(1)
    room.offlimits["point1Index"][0] = 10 ** 6
    with pytest.raises(pkg.KernelError, match="point1Index"):
        k.wrapper_ex(room, 1, 1, seed=1)
    room = S.make_config(1)
    room.rss["TargetIndex"][0] = -1
    with pytest.raises(pkg.KernelError, match="relationship"):
        k.wrapper_ex(room, 1, 1, seed=1)
    with pytest.raises(pkg.KernelError, match="gridxDim"):
        k.wrapper_ex(S.make_config(1), 0, 1, seed=1)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    k = pkg.Kernel()
    room = S.make_config(1)
    with pytest.raises(pkg.KernelError) as e:
        k.wrapper(room, 2, 10)
    assert "failed" in str(e.value)
    with pytest.raises(pkg.KernelError):
        k.eval_costs(room, room.cfg)


def test_missing_library_is_a_loud_error(tmp_path):
    """No silent fallback: without libKernel.so the binding raises, it does not compute elsewhere."""
    code = ("import importlib, sys; sys.path.insert(0, %r); p = importlib.import_module('metropolis-hastings-gpgpu_b200');\n"
            "try:\n    p.Kernel()\nexcept p.KernelError as e:\n    print('LOUD', e)\n" % ROOT)
    env = dict(os.environ, MH_LIB=str(tmp_path / "nope.so"))
    out = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert "LOUD" in out.stdout and "missing" in out.stdout, out.stdout + out.stderr


@pytest.mark.skipif(_has_gpu(), reason="exercises the no-GPU path of the reference arm")
@pytest.mark.timeout(240)
def test_bench_reference_arm_never_crashes_without_a_gpu():
    """`bench.py --impl reference` must always print one JSON line: with no device the reference's
    CUDA kernel cannot run, and the arm falls back to the transcribed C loop on the host cores."""
    import json
    out = subprocess.run([os.sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=230)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0
