"""Multi-GPU behind the C ABI (mhOptions.n_devices / devices[], env MH_DEVICES) and the device-side ranking
calls.  The reference's caller is ONE process calling KernelWrapper (Kernel.cu:873); what it can reach is what
the C ABI offers, so the fan-out over GPUs lives inside libKernel.so: one context and one stream per device,
every device copying its slice straight into the caller's one result block.

On a one-GPU box the same code path runs with a repeated ordinal (devices = [0, 0, 0]: three shards sharing
the GPU); with more GPUs visible the tests use all of them as well."""
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
L, S = pkg.layout, pkg.synth
HERE = os.path.dirname(os.path.abspath(__file__))


def device_lists(kernel):
    out = [[0, 0, 0], [0, 0]]
    n = kernel.device_count()
    if n > 1:
        out += [list(range(n)), list(range(n - 1, -1, -1))]
    return out


@pytest.mark.parametrize("cid,chains,iters", [(2, 1000, 150), (3, 333, 120), (1, 7, 300)])
def test_multi_device_run_equals_the_one_device_run(kernel, cid, chains, iters):
    """Per-chain results do not depend on how many devices the chains are spread over -- with the DEFAULT lane
    width: it is chosen from the job's chain count (mhOptions.total_chains), not from a shard's."""
    room = S.make_config(cid)
    want_p, want_c = kernel.wrapper_ex(room, chains, iters, seed=41)
    for devs in device_lists(kernel):
        p, c = kernel.wrapper_ex(room, chains, iters, seed=41, devices=devs)
        assert p.tobytes() == want_p.tobytes() and c.tobytes() == want_c.tobytes(), (cid, devs)


def test_multi_device_context_calls(kernel):
    room = S.make_config(2)
    chains = 2500
    with kernel.create(room, chains, seed=8) as one, kernel.create(room, chains, seed=8, devices=[0, 0, 0]) as many:
        sh = many.shape()
        assert len(sh["devices"]) == 3 and sum(sh["chains_per_device"]) == chains and sh["lanes_per_chain"] == one.shape()["lanes_per_chain"]
        for ctx in (one, many):
            ctx.run(60)
            ctx.run(40)
        p1, c1 = one.results()
        p2, c2 = many.results()
        assert p1.tobytes() == p2.tobytes() and c1.tobytes() == c2.tobytes()
        assert one.best() == many.best()
        i = int(np.argmax(c1["totalCosts"]))
        assert many.best() == (i, c1["totalCosts"][i])
        for k in (1, 10, 300, 2000):
            a, b = one.top_k(k), many.top_k(k)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), k
        ms, launches = many.stats()
        assert ms > 0 and launches >= 6
        many.reset()
        one.reset()
        many.run(30)
        one.run(30)
        assert one.results()[0].tobytes() == many.results()[0].tobytes()
        for min_d, rw in ((0.75, 0.0), (1.0, 0.5), (1e9, 0.0)):  # distinct top-k: the picks' layouts cross devices through the host
            a, b = one.top_k_distinct(9, min_d, rw), many.top_k_distinct(9, min_d, rw)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (min_d, rw)
        for call in (many.device_results, lambda: many.set_stream(0), lambda: many.best_key(8)):
            with pytest.raises(pkg.KernelError, match="multi-device"):
                call()
    # traces
    with kernel.create(room, 90, seed=3) as one, kernel.create(room, 90, seed=3, devices=[0, 0, 0, 0]) as many:
        assert one.run_traced(50).tobytes() == many.run_traced(50).tobytes()
    # more devices than chains: the empty shards are left out
    p, c = kernel.wrapper_ex(room, 2, 20, seed=1, devices=[0] * 8)
    q, d = kernel.wrapper_ex(room, 2, 20, seed=1)
    assert p.tobytes() == q.tobytes()
    with pytest.raises(pkg.KernelError, match="chain_stride"):
        kernel.wrapper_ex(room, 8, 5, seed=1, devices=[0, 0], chain_stride=2)


def test_multi_device_tempering_keeps_whole_ladders_per_device(kernel):
    room = S.make_config(1)
    opts = dict(seed=5, beta_start=0.5, beta_end=8.0, tempering_rungs=4, exchange_interval=25)
    pa, ca = kernel.wrapper_ex(room, 44, 230, **opts)
    pb, cb = kernel.wrapper_ex(room, 44, 230, devices=[0, 0, 0], **opts)
    assert pa.tobytes() == pb.tobytes() and ca.tobytes() == cb.tobytes()
    with kernel.create(room, 44, **opts) as one, kernel.create(room, 44, devices=[0, 0, 0], **opts) as many:
        one.run(100)
        many.run(100)
        assert all(c % 4 == 0 for c in many.shape()["chains_per_device"])
        a1, b1 = one.tempering_stats(4)
        a2, b2 = many.tempering_stats(4)
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2)


def test_a_shard_with_total_chains_equals_its_slice_of_the_whole_job(kernel):
    """The default lane width follows mhOptions.total_chains: a 96-chain shard of a 65536-chain job returns the
    bytes the whole job returns for those chains, nobody pinning lanes_per_chain (BASELINE section 5 gate 4)."""
    room = S.make_config(3)
    pa, ca = kernel.wrapper_ex(room, 65536, 40, seed=77)
    pb, cb = kernel.wrapper_ex(room, 96, 40, seed=77, chain_offset=4000, total_chains=65536)
    assert pa[4000:4096].tobytes() == pb.tobytes() and ca[4000:4096].tobytes() == cb.tobytes()
    with kernel.create(room, 96, seed=77, chain_offset=4000, total_chains=65536) as s, kernel.create(room, 96, seed=77, chain_offset=4000) as lone:
        assert s.shape()["lanes_per_chain"] == 8 and lone.shape()["lanes_per_chain"] == 32


def test_plain_c_caller_over_several_devices(kernel, tmp_path):
    """tests/c/call_kernel_wrapper.c calls the reference's own entry point, KernelWrapper, which has no options:
    env MH_DEVICES spreads it over GPUs.  Its output with 1 device, with a repeated ordinal and with every
    visible device must be the same bytes."""
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "metropolis-hastings-gpgpu_b200")
    exe = tmp_path / "call_kernel_wrapper"
    subprocess.run(["gcc", "-std=c11", "-O1", "-I", os.path.join(root, "include"), os.path.join(HERE, "c", "call_kernel_wrapper.c"),
                    "-o", str(exe), "-L", libdir, "-lKernel", "-Wl,-rpath," + libdir], check=True)
    outs = {}
    for name, devs in (("one", None), ("dup", "0,0,0"), ("all", "all")):
        env = dict(os.environ, MH_SEED="77")
        env.pop("MH_DEVICES", None)
        if devs:
            env["MH_DEVICES"] = devs
        r = subprocess.run([str(exe), "37", "120"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        outs[name] = r.stdout
    assert outs["one"] == outs["dup"] == outs["all"] and outs["one"].count("costs") == 37
    bad = subprocess.run([str(exe), "4", "10"], capture_output=True, text=True, env=dict(os.environ, MH_DEVICES="0;1"))
    assert bad.returncode != 0 and "MH_DEVICES" in bad.stderr


def test_device_side_ranking(kernel):
    """KernelBest (multi-block arg-max), KernelTopK (tile sort + merge on the device for k <= 512, host sort above)
    and KernelTopKDistinct (all rounds enqueued at once) against numpy on the returned costs."""
    room = S.make_config(2)
    for chains in (1, 31, 1024, 5000, 70001):
        with kernel.create(room, chains, seed=12) as ctx:
            ctx.run(40)
            _, c = ctx.results()
            order = np.lexsort((np.arange(chains), -c["totalCosts"].astype(np.float64)))
            assert ctx.best() == (int(order[0]), c["totalCosts"][order[0]])
            for k in (1, 7, 300, 512, 513, 3000):
                idx, tot = ctx.top_k(k)
                m = min(k, chains)
                assert np.array_equal(idx, order[:m]) and np.array_equal(tot, c["totalCosts"][order[:m]]), (chains, k)
    # ties go to the lower chain: zero iterations leave every chain on the same layout
    with kernel.create(room, 3000, seed=1) as ctx:
        assert ctx.best()[0] == 0
        assert np.array_equal(ctx.top_k(20)[0], np.arange(20))
        assert list(ctx.top_k_distinct(5, 0.1)[0]) == [0]


def test_best_key_on_contiguous_and_strided_contexts(kernel):
    """KernelBestKey -> KernelDecodeBestKey against argmax of KernelResults: contiguous shard, strided shard (the
    cross-GPU tempering layout), and dist.global_best end to end on one rank."""
    import torch
    room = S.make_config(2)
    dev = torch.device("cuda", 0)
    for kw in (dict(chain_offset=5000), dict(chain_offset=3, chain_stride=8), dict()):
        with kernel.create(room, 700, seed=6, **kw) as ctx:
            ctx.run(80)
            key = torch.zeros(1, dtype=torch.int64, device=dev)
            ctx.best_key(key.data_ptr())
            ctx.synchronize()                                   # the key was written on the context's own stream
            g, t = kernel.decode_best_key(int(key.item()))
            _, c = ctx.results()
            i = int(np.argmax(c["totalCosts"]))
            assert g == kw.get("chain_offset", 0) + i * kw.get("chain_stride", 1) and t == c["totalCosts"][i]
    with kernel.create(room, 700, seed=6, chain_offset=5000, total_chains=9000) as ctx:
        ctx.run(80)
        g, t, lay = pkg.dist.global_best(kernel, ctx, room.n, 5000, 9000, 0, 1, dev)
        pts, c = ctx.results()
        i = int(np.argmax(c["totalCosts"]))
        assert g == 5000 + i and t == c["totalCosts"][i]
        assert lay.cpu().numpy().tobytes() == pts[i].tobytes()
    with kernel.create(room, 4, seed=6, chain_offset=2 ** 32 - 2) as ctx:
        key = torch.zeros(1, dtype=torch.int64, device=dev)
        with pytest.raises(pkg.KernelError, match="2\\^32"):
            ctx.best_key(key.data_ptr())


def test_a_failed_launch_leaves_the_context_usable(kernel, monkeypatch):
    """Fault injection (env MH_FAULT=launch fails the next chain launch inside launch_segment): the call reports the
    failure, the event pair taken for it is given back (KernelStats must not trip over an unrecorded event), and the
    context keeps working afterwards."""
    room = S.make_config(1)
    with kernel.create(room, 64, seed=2) as ctx, kernel.create(room, 64, seed=2) as ref:
        ctx.run(50)
        ref.run(50)
        monkeypatch.setenv("MH_FAULT", "launch")
        with pytest.raises(pkg.KernelError, match="failed"):
            ctx.run(50)
        ms, launches = ctx.stats()
        assert ms > 0 and launches == 1
        monkeypatch.delenv("MH_FAULT")
        ctx.run(50)
        ref.run(50)
        assert ctx.results()[0].tobytes() == ref.results()[0].tobytes()
        assert ctx.stats()[1] == ref.stats()[1]


def test_ladder_tuning_retargets_the_chains(kernel):
    """KernelTemperingLadder / ProposeLadder / SetLadder: after a re-target every ladder still holds a permutation of
    its rungs -- the NEW rungs; the statistics restart; multi-device contexts follow; KernelReset restarts the chains
    on the tuned ladder."""
    room = S.make_config(2)
    rungs, ex = 6, 10
    opts = dict(seed=3, beta_start=0.25, beta_end=8.0, tempering_rungs=rungs, exchange_interval=ex)
    for devs in (None, [0, 0, 0]):
        with kernel.create(room, rungs * 50, devices=devs, **opts) as ctx:
            geo = 0.25 * 32.0 ** (np.arange(rungs) / (rungs - 1))
            np.testing.assert_allclose(ctx.ladder(rungs), geo, rtol=1e-6)
            ctx.run(20 * ex)
            att, acc = ctx.tempering_stats(rungs)
            assert att.sum() > 0
            new = kernel.propose_ladder(ctx.ladder(rungs), att, acc, damping=1.0)
            assert new[0] == pytest.approx(0.25) and new[-1] == pytest.approx(8.0) and np.all(np.diff(new) > 0)
            ctx.set_ladder(new)
            np.testing.assert_allclose(ctx.ladder(rungs), new, rtol=1e-6)
            att2, acc2 = ctx.tempering_stats(rungs)
            assert att2.sum() == 0 and acc2.sum() == 0
            tr = ctx.run_traced(3 * ex)
            got = np.sort(tr["beta"][-1].reshape(50, rungs), axis=1)
            np.testing.assert_allclose(got, np.tile(np.float32(new), (50, 1)), rtol=1e-6)
            ctx.reset()
            tr = ctx.run_traced(1)
            np.testing.assert_allclose(tr["beta"][0].reshape(50, rungs), np.tile(np.float32(new), (50, 1)), rtol=1e-6)
    with kernel.create(room, 8, seed=1) as plain:
        with pytest.raises(pkg.KernelError, match="ladder"):
            plain.ladder(4)


def test_chunked_one_shot_call_returns_the_same_bytes(kernel, monkeypatch):
    """KernelWrapperEx runs a call with a large result block as several launches (chunks of chains) so that scoring
    and the D2H of one chunk overlap the kernels of the next; a chain's result must not depend on the chunking.
    MH_CHUNKS forces the chunk count (1 = one launch)."""
    for cid, chains, iters, kw in ((2, 5000, 120, dict()), (3, 700, 60, dict(result_mode=1, beta_start=0.5, beta_end=6.0, schedule=1)),
                                   (2, 3001, 80, dict(devices=[0, 0, 0])), (1, 5, 40, dict())):
        room = S.make_config(cid)
        monkeypatch.setenv("MH_CHUNKS", "1")
        want = kernel.wrapper_ex(room, chains, iters, seed=9, **kw)
        for chunks in ("2", "3", "8"):
            monkeypatch.setenv("MH_CHUNKS", chunks)
            got = kernel.wrapper_ex(room, chains, iters, seed=9, **kw)
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes(), (cid, chunks)
    monkeypatch.delenv("MH_CHUNKS")
    room = S.make_config(4)                                      # the default rule: one chunk per 48 MB of results
    a = kernel.wrapper_ex(room, 40000, 8, seed=2)                # 192 MB: 4 chunks
    monkeypatch.setenv("MH_CHUNKS", "1")
    b = kernel.wrapper_ex(room, 40000, 8, seed=2)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


def test_a_round_one_caller_with_the_short_options_struct_still_works(kernel):
    """mhOptions grew from 88 to 136 bytes; struct_size says how much the caller filled in.  A caller compiled against
    the round-1 header passes 88: whatever lies behind those 88 bytes must be ignored."""
    import ctypes as C
    room = S.make_config(1)
    want = kernel.wrapper_ex(room, 50, 60, seed=4)
    o = pkg.binding.make_options(seed=4)
    o["n_devices"] = 9                                          # a tail that would be refused if it were read ("at most 8 devices")
    o["struct_size"] = 88
    g = np.zeros(1, L.gpuConfig)
    g["gridxDim"], g["blockxDim"], g["iterations"] = 50, 64, 60
    res = kernel.lib.KernelWrapperEx(*kernel._room_args(room), g.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    assert res, kernel.last_error()
    got = kernel._unpack(res, 50, room.n)
    assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes()


def test_interactive_session_from_plain_c(kernel, tmp_path):
    """tests/c/interactive_session.c: the persistent context, stepping, distinct top-k and result fetch driven from
    plain C through include/mh_kernel.h (one device and spread over three shards); its picks must be the Python
    binding's for the same room, seed and options."""
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "metropolis-hastings-gpgpu_b200")
    exe = tmp_path / "interactive_session"
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(HERE, "c", "interactive_session.c"), "-o", str(exe), "-L", libdir, "-lKernel", "-Wl,-rpath," + libdir], check=True)
    room = S.reference_main_fixture()
    chains, steps = 300, 3
    want = []
    with kernel.create(room, chains, seed=2026, result_mode=1) as ctx:
        for s in range(steps):
            ctx.run(150)
            idx, tot = ctx.top_k_distinct(4, 0.5, 0.25)
            pts, _ = ctx.results()
            for j, (i, t) in enumerate(zip(idx, tot)):
                want.append(f"step {s} pick {j} chain {i} total {float(t).hex()} first {float(pts[i, 0]['x']).hex()} "
                            f"{float(pts[i, 0]['y']).hex()} {float(pts[i, 0]['rotY']).hex()}")
    for ndev in ("0", "3"):
        out = subprocess.run([str(exe), str(chains), str(steps), ndev], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = []
        for line in out.stdout.strip().splitlines():          # C prints %a, Python float.hex(): compare the values
            f = line.split()
            got.append(" ".join(f[:7] + [float.fromhex(f[7]).hex(), f[8]] + [float.fromhex(v).hex() for v in f[9:]]))
        assert got == want, ndev
