"""CPU tests of the multi-GPU host logic (metropolis-hastings-gpgpu_b200/dist.py) with
world_size 2 over gloo: chain sharding by global chain id, the packed (totalCosts, chain) key
whose MAX all-reduce is the arg-best (NCCL has no MAXLOC), and the owner's layout broadcast.
The per-chain work itself is stood in for by the oracle (a checker), sharded exactly the way
bench.py shards the GPU contexts."""
import importlib
import os
import socket

import numpy as np
import pytest

pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
D, S = pkg.dist, pkg.synth


def test_shards_partition_the_chains():
    for total in (1, 7, 64, 65536, 262144, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o, c), (o2, _) in zip(spans, spans[1:]):
                assert o + c == o2
            for g in (0, total // 3, total - 1):
                r = D.owner_of(g, total, world)
                o, c = spans[r]
                assert o <= g < o + c


def test_key_orders_like_total_then_lower_chain():
    k = pkg.Kernel()
    rng = np.random.default_rng(0)
    totals = np.concatenate([rng.normal(0, 500, 400).astype(np.float32), np.float32([0.0, -0.0, 1e-30, -1e-30, 3.4e38, -3.4e38])])
    chains = rng.integers(0, 2 ** 31, len(totals))
    keys = [D.pack_best_key(t, c) for t, c in zip(totals, chains)]
    for i in range(len(keys)):
        g, t = k.decode_best_key(keys[i])
        assert g == chains[i] and (t == totals[i] or (t == 0 and totals[i] == 0))
    order = np.argsort(np.array(keys, dtype=np.int64), kind="stable")
    st = totals[order]
    assert np.all(np.diff(st) >= 0)
    # equal totals: the lower chain id wins a MAX
    assert D.pack_best_key(1.5, 10) > D.pack_best_key(1.5, 11)
    assert D.pack_best_key(-1.5, 10) > D.pack_best_key(-1.5, 11)
    assert D.pack_best_key(2.0, 99) > D.pack_best_key(1.5, 0) > D.pack_best_key(-0.5, 0) > D.pack_best_key(-2.0, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_chains, iters, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from oracle_lib import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk = importlib.import_module("metropolis-hastings-gpgpu_b200")
    k = pk.Kernel()
    room = pk.synth.make_config(1)
    offset, count = pk.dist.shard(total_chains, rank, world)
    pts, costs = Oracle().run(room, count, iters, seed=77, chain_offset=offset)     # this rank's shard
    li = int(np.argmax(costs["totalCosts"]))
    key = torch.tensor([pk.dist.pack_best_key(costs["totalCosts"][li], offset + li)], dtype=torch.int64)
    dist.all_reduce(key, op=dist.ReduceOp.MAX)
    g, total = k.decode_best_key(int(key.item()))
    owner = pk.dist.owner_of(g, total_chains, world)
    layout = torch.zeros(room.n * 24, dtype=torch.uint8)
    if rank == owner:
        layout.copy_(torch.from_numpy(np.frombuffer(pts[g - offset].tobytes(), np.uint8).copy()))
    dist.broadcast(layout, src=owner)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([[g, total], layout.numpy().astype(np.float64)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_global_best_equals_unsharded(tmp_path, oracle):
    import torch.multiprocessing as mp
    total_chains, iters, world = 37, 60, 2
    mp.spawn(_worker, args=(world, _free_port(), total_chains, iters, str(tmp_path)), nprocs=world, join=True)
    room = S.make_config(1)
    pts, costs = oracle.run(room, total_chains, iters, seed=77)
    best = int(np.argmax(costs["totalCosts"]))
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert int(got[0]) == best
        assert np.float32(got[1]) == costs["totalCosts"][best]
        assert got[2:].astype(np.uint8).tobytes() == pts[best].tobytes()
