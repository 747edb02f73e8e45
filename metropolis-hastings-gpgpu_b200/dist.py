"""Multi-GPU host logic: chain sharding and the global-best reduction.

Chains are independent, so ranks shard contiguous ranges of the GLOBAL
chain id and never exchange data on the sampling path.  The only collective is the arg-best at
the end of a run: NCCL has no MINLOC/MAXLOC, so each rank packs (order-preserving totalCosts,
complemented global chain id) into one signed 64-bit key on its device (KernelBestKey), the
keys are MAX-all-reduced (8 bytes), and the owner of the winning chain broadcasts its n x 24 B
layout.  torch.distributed is plumbing only; the packing and the arg-max run in libKernel.so.

Bit-identity of a sharded run with the unsharded one needs every shard to use the same lane width (it fixes
the float reduction order): create the contexts with `total_chains=` the size of the whole job (shard_options
below does), and the library picks the width from that instead of from the shard's own count.

Stream ordering: a context runs on a non-blocking stream of its own unless KernelSetStream says otherwise, and
torch / NCCL work on torch's current stream.  The helpers below put the context on torch's current stream
(`ctx.set_stream`) before they touch its device buffers, so that kernel -> collective -> kernel is ordered
without a host synchronisation.
"""
import numpy as np


def shard(total_chains, rank, world):
    """Contiguous range [offset, offset+count) of global chain ids owned by `rank`."""
    base, rem = divmod(total_chains, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def shard_options(total_chains, rank, world):
    """(chain count, mhOptions keywords) of rank's shard: contiguous global chain ids, and the job's size so
    that every shard picks the lane width the unsharded job would."""
    offset, count = shard(total_chains, rank, world)
    return count, {"chain_offset": offset, "total_chains": total_chains}


def owner_of(global_chain, total_chains, world):
    base, rem = divmod(total_chains, world)
    edge = rem * (base + 1)
    if global_chain < edge:
        return global_chain // (base + 1)
    return rem + (global_chain - edge) // base


def pack_best_key(total, global_chain):
    """Host restatement of mh_bestkey_kernel (csrc/mh_kernels.cu) for CPU tests."""
    u = int(np.float32(total).view(np.uint32))
    u = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
    k = (u << 32) | (0xFFFFFFFF - (int(global_chain) & 0xFFFFFFFF))
    k ^= 0x8000000000000000
    return k - (1 << 64) if k >= (1 << 63) else k


class DeviceBytes:
    """Zero-copy torch view of device memory owned by libKernel.so."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_view(ptr, nbytes, device):
    import torch
    return torch.as_tensor(DeviceBytes(ptr, nbytes), device=device)


def _on_torch_stream(ctx):
    """Run the context on torch's current stream (idempotent): everything torch enqueues afterwards is ordered
    after the context's kernels and vice versa."""
    import torch
    h = torch.cuda.current_stream().cuda_stream
    if getattr(ctx, "_stream_handle", None) != h:
        ctx.set_stream(h)
        ctx._stream_handle = h


def global_best(kernel, ctx, n, offset, total_chains, rank, world, device, dist=None):
    """Returns (global chain id, totalCosts, layout bytes tensor of n*24 B) on every rank."""
    import torch
    _on_torch_stream(ctx)
    key = torch.zeros(1, dtype=torch.int64, device=device)
    ctx.best_key(key.data_ptr())
    if dist is not None and world > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX)
    g, total = kernel.decode_best_key(int(key.item()))
    owner = owner_of(g, total_chains, world)
    layout = torch.empty(n * 24, dtype=torch.uint8, device=device)
    if rank == owner:
        d_points, _ = ctx.device_results()
        view = device_view(d_points, ctx.n_chains * n * 24, device)
        layout.copy_(view[(g - offset) * n * 24:(g - offset + 1) * n * 24])
    if dist is not None and world > 1:
        dist.broadcast(layout, src=owner)
    return g, total, layout


def tempering_epoch(ctxs_or_ctx, iterations, device, dist=None, world=1):
    """One exchange epoch of a ladder spread over ranks (chain_stride = world): run `iterations`
    (= exchange_interval) MH steps, all-gather the per-chain totals and betas (8 bytes per chain),
    let every rank decide its swaps.  `ctxs_or_ctx` is this rank's context -- or, to emulate the
    ranks of a multi-GPU run on ONE device (tests), the list of all ranks' contexts."""
    import torch
    ctxs = ctxs_or_ctx if isinstance(ctxs_or_ctx, (list, tuple)) else [ctxs_or_ctx]
    for ctx in ctxs:
        _on_torch_stream(ctx)
        ctx.run(iterations)
    tots, bets = [], []
    for ctx in ctxs:
        dt, db = ctx.tempering_state()
        tots.append(device_view(dt, 4 * ctx.n_chains, device).view(torch.float32))
        bets.append(device_view(db, 4 * ctx.n_chains, device).view(torch.float32))
    if dist is not None and world > 1:
        assert len(ctxs) == 1
        all_t = torch.empty(world * ctxs[0].n_chains, dtype=torch.float32, device=device)
        all_b = torch.empty_like(all_t)
        dist.all_gather_into_tensor(all_t, tots[0])
        dist.all_gather_into_tensor(all_b, bets[0])
    else:
        all_t, all_b = torch.cat(tots), torch.cat(bets)        # rank-major, like an all-gather
    for ctx in ctxs:
        ctx.tempering_exchange(all_t.data_ptr(), all_b.data_ptr())
    return all_t, all_b
