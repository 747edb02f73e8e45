#!/usr/bin/env python
"""bench.py -- MH proposals evaluated per second (chains x iterations / s), BASELINE.json's
headline metric, on the 50-object living room (config 3: n=50, C=25, R=50, 65536 chains x
10000 iterations, all cost terms, beta = 2).

  python bench.py --gpus N --steps K --warmup W            this repo's sm_100a path
  python bench.py --eval-mode 3 ...                         the same with the plain scan (every term from scratch)
  python bench.py --scaling strong --gpus N ...             BASELINE config 4 as named: 262144 chains of 200 objects
                                                            x 1000 iterations SHARDED over the N GPUs (strong scaling)
  python bench.py --impl reference ...                      the reference's own kernel, rebuilt for
                                                            sm_100 from /root/reference (oracle/_ref); one GPU always
                                                            (the reference has no multi-GPU path, Kernel.cu:951)

One step = one pass of the hot path over the whole batch: every chain runs `iterations` MH steps
from the caller's layout.  The library's default evaluation (MH_EVAL_FULL) runs, from 28 objects up (18 for big jobs), in
its memo form: every proposal's costs, every accept decision and every returned bit equal the plain
full re-evaluation's (tested), at a fraction of the work; the plain scan's rate is reported beside it.  `value` is timed with the problem and chain state already resident in
HBM (KernelCreate once, then KernelReset + KernelRun per step, CUDA events around the kernel);
`e2e` is the same job through the reference-facing call KernelWrapperEx with host buffers (H2D of
the room, the kernels, D2H of every layout and its costs, result assembly).  Prints ONE JSON line.

Sub-records of the default line: `config4_strong` (config 4 as named, total chains fixed, sharded over the N
GPUs: kernel-only and e2e rates, arg-best time, a hash of global chains 0..1023 that must be equal for every N),
`c_abi_multi_gpu` (N > 1: the same jobs through ONE process and the C ABI's device list, mhOptions.devices),
`config2_as_named` (N = 1: 16 objects, 1024 chains x 10000 iterations, beside the reference kernel on the same
workload).
"""
import argparse
import hashlib
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "MH proposals evaluated/sec (chains x iters/s) at 50 objects"
PROFILE_FILE = "r2g_memo_n50_g8_ncu_full.txt"   # ncu --set full capture of the default kernel at 65536 chains
UNIT = "proposals/s"


def profiled_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the chain kernel at 65536 chains,
    from the committed `ncu --set full` summary (profiles/).  The chain state lives in shared memory, so
    the traffic is the result block written at the end of a launch and does not depend on the iteration
    count: it is reported to show that HBM is not the bound, not as the roofline denominator."""
    path = os.path.join(ROOT, "profiles", PROFILE_FILE)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    try:
        for line in open(path):
            for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
                if line.startswith(key):
                    unit = line[len(key):line.index("]")]
                    total += float(line.split("=")[1]) * scale[unit]
        return total or None
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            busy = [s for s, w in zip(sm, power) if w >= 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def cpu_baseline(room, target_seconds=12.0):
    """The oracle's C transcription of the per-chain loop (oracle/mh_oracle.c), one chain per
    OpenMP thread on all host cores, on a bounded sample of the same workload."""
    from oracle_lib import Oracle, build_native_oracle
    native = build_native_oracle()
    o = Oracle(native) if native else Oracle()
    flags = "gcc -O2 -march=native" if native else "gcc -O2"
    try:                                       # torchrun pins OMP_NUM_THREADS=1: ask for the real core count
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    chains = threads * 4
    o.run(room, chains, 20, seed=1, timed=True, threads=threads)                           # wake the thread pool
    _, _, secs, th = o.run(room, chains, 400, seed=1, timed=True, threads=threads)         # calibrate
    rate = chains * 400 / max(secs, 1e-6)
    iters = max(20, int(rate * target_seconds / chains))
    _, costs, secs, th = o.run(room, chains, iters, seed=1, timed=True, threads=threads)
    one_iters = max(20, int(rate / max(threads, 1) * 2.0))     # ~2 s of one core (single-core and all-core)
    _, _, one_secs, _ = o.run(room, 1, one_iters, seed=1, timed=True, threads=1)
    return {"value": chains * iters / secs, "unit": UNIT, "cores": th, "kind": "port", "single_core": one_iters / max(one_secs, 1e-9),
            "sample": f"{chains} chains x {iters} iterations of the same room ({secs:.1f} s), {flags} -ffp-contract=off, OpenMP",
            "seconds": secs, "best_totalCosts": float(costs["totalCosts"].max()),
            "median_final_totalCosts": float(np.median(costs["totalCosts"]))}


def time_to_best_cost(k, room, chains, target, epoch=50, max_iterations=60000):
    """The second half of BASELINE.json's metric: wall time until the global best totalCosts (the sampler
    maximises it, quirk Q10) first reaches `target`, checked every `epoch` iterations with the device
    arg-max (KernelBest: one 8-byte read per check).  Starts from the caller's layout, context creation
    included."""
    t0 = time.perf_counter()
    done, best = 0, -1e30
    with k.create(room, chains, seed=424242) as ctx:
        while done < max_iterations:
            ctx.run(epoch)
            done += epoch
            best = ctx.best()[1]
            if best >= target:
                return {"seconds": time.perf_counter() - t0, "iterations_per_chain": done, "chains": chains, "reached": True,
                        "best_totalCosts": best}
    return {"seconds": time.perf_counter() - t0, "iterations_per_chain": done, "chains": chains, "reached": False, "best_totalCosts": best}


def reference_gpu(config_id, chains, iters, steps, warmup, timeout_s=300):
    """The reference's own KernelWrapper rebuilt for sm_100 (oracle/_ref/libKernel_ref_nb.so: the
    unmodified Kernel.cu with the one divergent barrier that deadlocks on sm_70+ neutralised, see
    oracle/ref_gpu_harness.cu), at the launch shape its main() uses (blockxDim = 64).  Runs in a
    subprocess under a timeout: a hang of the reference must not take the bench down.
    Returns (proposals/s over the whole call, proposals/s by device events, seconds per step)."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "nb", str(config_id), str(chains), str(iters), "64",
           str(warmup), str(steps)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    if out.returncode != 0:
        raise RuntimeError("reference kernel failed: " + (out.stderr.strip().splitlines() or ["?"])[-1])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    reference_gpu.last = r
    return r["proposals_per_s"], chains * iters / r["dev_s"], r["wall_s"]


def quick_rate(k, room, chains, iters, **opts):
    """proposals/s of the chain kernel alone (CUDA events), one warm-up launch then one measured."""
    with k.create(room, chains, seed=1, **opts) as ctx:
        ctx.run(max(1, iters // 8))
        ctx.synchronize()
        ms0, _ = ctx.stats()
        ctx.reset()
        ctx.run(iters)
        ctx.synchronize()
        ms1, _ = ctx.stats()
    return chains * iters / ((ms1 - ms0) * 1e-3)


def other_configs(k, pkg):
    """Kernel-only throughput on the other rooms of BASELINE.json (parity-test cases, not the
    headline): full evaluation, and delta evaluation where it is the faster mode."""
    out = {}
    for cid, chains, iters in ((1, 65536, 2000), (2, 65536, 1000), (4, 16384, 60)):
        room = pkg.synth.make_config(cid)
        e = {"n": room.n, "chains": chains, "iterations": iters, "full_eval": quick_rate(k, room, chains, iters),
             "flops_per_proposal_contract": room.flops_per_proposal()}
        if cid == 4:
            e["full_eval_plain_scan"] = quick_rate(k, room, chains, iters, eval_mode=3)
            e["delta_eval"] = quick_rate(k, room, chains, 256, eval_mode=1)
        out[f"config{cid}"] = e
    room = pkg.synth.make_config(3)
    out["config3_delta_eval"] = quick_rate(k, room, 65536, 512, eval_mode=1)
    out["note"] = ("full_eval = the library default (bit-identical memo form from 18-28 objects up); full_eval_plain_scan = every term "
                   "from scratch; delta_eval = incremental running sums, statistically equivalent (MH_EVAL_DELTA)")
    return out


def config3_annealing(k, room, chains, iterations):
    """"run once with fixed beta = 2 and once with the annealing extension".  Same room, chain
    count, iteration budget and seed; the sampler maximises totalCosts (quirk Q10), so higher is better.  The schedule is
    evaluated inside the chain kernel (one beta per iteration), so the rate is the fixed-beta kernel's."""
    import numpy as np
    out = {"chains": chains, "iterations": iterations,
           "note": "kernel-only rate (CUDA events) and the totalCosts the chains end on: the global best, the median and the 99th "
                   "percentile of the per-chain finals; result_mode = final layout in both runs"}
    for name, opts in (("fixed_beta2", {}),
                       ("geometric_0.5_to_8", dict(beta_start=0.5, beta_end=8.0, schedule=1, schedule_length=iterations)),
                       ("linear_0.5_to_8", dict(beta_start=0.5, beta_end=8.0, schedule=2, schedule_length=iterations))):
        with k.create(room, chains, seed=7, **opts) as ctx:
            ctx.run(iterations)
            ctx.synchronize()
            ms, _ = ctx.stats()
            _, costs = ctx.results()
        t = costs["totalCosts"].astype(np.float64)
        out[name] = {"proposals_per_s": chains * iterations / (ms * 1e-3), "best_totalCosts": float(t.max()),
                     "median_final_totalCosts": float(np.median(t)), "p99_final_totalCosts": float(np.percentile(t, 99))}
    return out


def run_tempering(args, pkg, k, room, pl):
    """BASELINE config 5: the config-3 room under parallel tempering, TIME TO TARGET COST.  A ladder of 8 rungs
    (geometric beta 0.25 .. 8); with N > 1 GPUs the rungs of every ladder are spread over the ranks
    (chain_stride = N) and neighbours exchange betas across NVLink: per epoch one all-gather of 8 bytes per chain
    (--rungs / --exchange-interval change the ladder).
    Five samplers get the same budget (chains x iterations per GPU) on each of >= 3 seeds, the global best
    totalCosts (the sampler maximises it, quirk Q10) is read after every epoch:
      plain_beta2        the reference's sampler, BETA = 2 (Kernel.cu:33)
      plain_beta8        plain MH at the ladder's coldest beta
      anneal_geometric   plain chains, beta annealed geometrically from the ladder's hottest to its coldest beta over the budget
      tempering_fixed    the geometric ladder, exchange every 100 iterations
      tempering_adapted  the same, the ladder re-tuned every 5 epochs during the first 30 % of the run from the
                         exchange statistics (KernelTemperingStats -> KernelTemperingProposeLadder ->
                         KernelTemperingSetLadder)
    target = median over the seeds of plain_beta2's final global best; reported per sampler: in how many seeds and
    after how many seconds the target is first reached, and the best at the end of the budget."""
    torch, dist, rank, world, device = pl.torch, pl.dist, pl.rank, pl.world, pl.device
    rungs, ex = args.rungs, args.exchange_interval
    assert rungs % world == 0 or world == 1, "the rungs of a ladder are spread evenly over the ranks"
    epochs = max(1, args.iterations // ex)
    iters = epochs * ex
    chains = max(rungs, args.chains - args.chains % rungs)
    total = chains * world
    seeds = [99, 1234, 777][:max(1, args.seeds)] if args.seeds <= 3 else [99, 1234, 777] + list(range(5000, 5000 + args.seeds - 3))
    stream = torch.cuda.current_stream().cuda_stream
    adapt_every, adapt_until = 5, int(0.3 * epochs)

    def make(kind, seed):
        if kind.startswith("plain"):
            beta = 2.0 if kind == "plain_beta2" else 8.0
            return k.create(room, chains, seed=seed, chain_offset=rank * chains, total_chains=total, beta_start=beta)
        if kind == "anneal_geometric":
            return k.create(room, chains, seed=seed, chain_offset=rank * chains, total_chains=total, beta_start=0.25, beta_end=8.0,
                            schedule=1, schedule_length=iters)
        opts = dict(seed=seed, beta_start=0.25, beta_end=8.0, tempering_rungs=rungs, exchange_interval=ex, total_chains=total)
        if world > 1:
            return k.create(room, chains, chain_offset=rank, chain_stride=world, **opts)
        return k.create(room, chains, **opts)

    def reduced_stats(ctx):
        att, acc = ctx.tempering_stats(rungs)
        t = torch.tensor(np.concatenate([att, acc]), dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(t)
        v = t.cpu().numpy()
        return v[:rungs - 1], v[rungs - 1:]

    def one_run(kind, seed):
        ctx = make(kind, seed)
        ctx.set_stream(stream)
        tempering = kind.startswith("tempering")
        pl.barrier()
        ms0, l0 = ctx.stats()
        t0 = time.perf_counter()
        curve, best = [], -1e30
        for e in range(epochs):
            if tempering and world > 1:
                pkg.dist.tempering_epoch(ctx, ex, device, dist, world)
            else:
                ctx.run(ex)
            best = max(best, pl.gmax(ctx.best()[1])[0])
            curve.append((time.perf_counter() - t0, (e + 1) * ex, best))
            if kind == "tempering_adapted" and (e + 1) % adapt_every == 0 and e + 1 <= adapt_until:
                att, acc = reduced_stats(ctx)
                ctx.set_ladder(k.propose_ladder(ctx.ladder(rungs), att, acc, damping=0.7))
        pl.barrier()
        wall = time.perf_counter() - t0
        ms1, l1 = ctx.stats()
        out = {"kind": kind, "seed": seed, "wall_s": wall, "final_best": best, "curve": curve, "kernel_ms": ms1 - ms0, "launches": int(l1 - l0)}
        if tempering:
            att, acc = reduced_stats(ctx)
            out["exchange_rates"] = [float(a) / max(1, int(t)) for a, t in zip(acc, att)]
            out["ladder"] = [float(b) for b in ctx.ladder(rungs)]
        ctx.close()
        return out

    kinds = ("plain_beta2", "plain_beta8", "anneal_geometric", "tempering_fixed", "tempering_adapted")
    one_run("tempering_fixed", 1)                               # warm-up: every kernel compiled and loaded, pools filled
    sampler = ClockSampler(pl.local_rank) if rank == 0 else None
    runs = {kind: [one_run(kind, s) for s in seeds] for kind in kinds}
    clocks = sampler.stop() if sampler else None
    target = float(np.median([r["final_best"] for r in runs["plain_beta2"]]))
    summary = {}
    for kind in kinds:
        hits = []
        for r in runs[kind]:
            hit = next(((t, it) for t, it, b in r["curve"] if b >= target), None)
            r["time_to_target_s"], r["iterations_to_target"] = (hit if hit else (None, None))
            r["curve"] = r["curve"][::max(1, len(r["curve"]) // 20)]           # keep the line short
            hits.append(hit)
        reached = [h for h in hits if h]
        summary[kind] = {"reached": f"{len(reached)}/{len(hits)}",
                         "median_time_to_target_s": float(np.median([h[0] for h in reached])) if reached else None,
                         "median_iterations_to_target": float(np.median([h[1] for h in reached])) if reached else None,
                         "median_final_best": float(np.median([r["final_best"] for r in runs[kind]])),
                         "median_wall_s": float(np.median([r["wall_s"] for r in runs[kind]]))}
    if rank == 0:
        fx = runs["tempering_fixed"]
        wall = float(np.median([r["wall_s"] for r in fx]))
        line = {"metric": METRIC + "; time-to-target-cost under parallel tempering", "value": total * iters / wall, "unit": UNIT, "n_gpus": world,
                "steps": len(seeds), "warmup": 1, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"config 5: config-3 room (n=50) under parallel tempering, {rungs} rungs beta 0.25..8, "
                                       f"{chains} chains/GPU x {iters} iterations, exchange every {ex}, {len(seeds)} seeds",
                           "parallelism": f"rungs of every ladder spread over {world} GPU(s), NCCL all-gather of 8 B/chain per epoch" if world > 1 else "whole ladders in one context"},
                "gpu_launches": int(sum(r["launches"] for r in fx)), "clocks": clocks,
                "tempering": {"rungs": rungs, "exchange_interval": ex, "epochs": epochs, "seeds": seeds,
                              "exchange_bytes_per_epoch_per_rank": 8 * chains if world > 1 else 0,
                              "target_totalCosts": target, "target_definition": "median over the seeds of plain_beta2's final global best at the same budget",
                              "summary": summary, "runs": runs,
                              "adaptation": {"every_epochs": adapt_every, "until_epoch": adapt_until, "damping": 0.7,
                                             "policy": "KernelTemperingProposeLadder: interior rungs at equal steps of the cumulative -log(exchange rate)"}}}
        print(json.dumps(line), flush=True)


def run_reference_arm(args, room, rank):
    """The reference's own implementation, on ONE GPU whatever --gpus says: Kernel.cu:951 launches on the current
    device and nothing in the reference spreads a job over devices, so n_gpus is reported as 1 and the driver's
    per-N ratio at N > 1 reads "N GPUs of this repo against the reference's single GPU"."""
    if rank != 0:
        return
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "gpus_requested": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 mixed",
            "data": "synthetic"}
    chains, iters = args.ref_chains, args.ref_iterations
    try:
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libKernel_ref_nb.so")):
            raise RuntimeError("oracle/_ref/libKernel_ref_nb.so not built")
        v_wall, v_dev, step_s = reference_gpu(args.config, chains, iters, args.steps, args.warmup)
        r = reference_gpu.last
        line.update(value=v_wall, ms_per_step=step_s * 1e3,
                    config={"workload": f"config {args.config} ({room.name}) n={room.n} C={room.C} R={room.R}, bounded sample {chains} chains x {iters} iterations per step "
                                        f"({chains} blocks of 64 threads = {chains / (148 * 32):.2f} waves of 32 resident blocks per SM on 148 SMs)",
                            "implementation": "reference Kernel.cu:873 KernelWrapper rebuilt for sm_100 (blockxDim=64; its one divergent __syncthreads, which deadlocks on sm_70+, neutralised), whole call incl. its H2D/D2H and curand init",
                            "multi_gpu": "none in the reference: one GPU at every --gpus"},
                    cpu_baseline={"value": v_wall, "unit": UNIT, "cores": 0, "kind": "reference",
                                  "sample": f"{chains} chains x {iters} iterations; the reference's path is a CUDA kernel, timed on the same B200 (device-event rate {v_dev:.4g}/s)"},
                    device_rates={"whole_call_device_events": v_dev, "without_init_rng": r.get("proposals_per_s_dev_without_init_rng"),
                                  "init_rng_ms": r.get("init_rng_ms"), "note": "device-event time of KernelWrapper, and the same minus the initRNG launch (Kernel.cu:939-943) measured alone at the same launch shape"},
                    e2e={"value": v_wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
        try:   # the other baseline of BASELINE.md section 2, beside it: the transcribed C loop on the host cores
            line["cpu_port_baseline"] = cpu_baseline(room, target_seconds=5.0)
        except Exception as ex2:
            line["cpu_port_baseline"] = {"unavailable": str(ex2)}
    except Exception as ex:  # no GPU build of the reference: its algorithm as transcribed C on the host cores
        cb = cpu_baseline(room, target_seconds=max(5.0, 4.0 * args.steps))
        cb["sample"] += f" (reference kernel unavailable: {ex})"
        line.update(value=cb["value"], ms_per_step=None, config={"workload": f"config {args.config} ({room.name}), bounded sample", "implementation": "oracle port on host cores"},
                    cpu_baseline=cb, e2e={"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
    print(json.dumps(line), flush=True)


class Plumbing:
    """rank / world / device / collectives of one bench process."""

    def __init__(self, torch, dist, rank, local_rank, world, device):
        self.torch, self.dist, self.rank, self.local_rank, self.world, self.device = torch, dist, rank, local_rank, world, device
        self.cpu_group = None
        if world > 1:
            try:   # a CPU-side group: ranks wait on it while rank 0 drives every GPU in-process (an NCCL barrier would spin on those GPUs)
                self.cpu_group = dist.new_group(backend="gloo")
            except Exception:
                self.cpu_group = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            if self.cpu_group is not None:
                self.dist.barrier(group=self.cpu_group)
            else:
                self.dist.barrier()

    def gmax(self, *values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]


def measure_job(pl, k, pkg, room, per_rank_chains, total_chains, offset, iterations, steps, warmup, flush, lanes=0, eval_mode=0, seed=20261018,
                hash_chains=0, clock_gpu=None):
    """One workload, sharded by global chain id: `value` leg (resident, KernelReset + KernelRun + NCCL arg-best per
    step, CUDA events around the chain kernels) and `e2e` leg (KernelWrapperEx with host buffers per step).
    Returns a dict; times are the max over ranks."""
    torch = pl.torch
    n = room.n
    stream = torch.cuda.current_stream().cuda_stream
    ctx = k.create(room, per_rank_chains, seed=seed, chain_offset=offset, total_chains=total_chains, lanes_per_chain=lanes, eval_mode=eval_mode)
    ctx.set_stream(stream)
    shape = ctx.shape()

    def step():
        flush.zero_()
        ctx.reset()
        ctx.run(iterations)
        return pkg.dist.global_best(k, ctx, n, offset, total_chains, pl.rank, pl.world, pl.device, pl.dist if pl.world > 1 else None)

    for _ in range(warmup):
        step()
    pl.barrier()
    ms0, l0 = ctx.stats()
    sampler = ClockSampler(clock_gpu) if clock_gpu is not None else None
    t0 = time.perf_counter()
    best = None
    for _ in range(steps):
        best = step()
    pl.barrier()
    wall = time.perf_counter() - t0
    ms1, l1 = ctx.stats()
    clocks = sampler.stop() if sampler else None
    # the arg-best alone: packed-key arg-max on the device, 8-byte MAX all-reduce, owner's layout broadcast
    # (one untimed call first: NCCL sets a broadcast from a new root up lazily)
    pkg.dist.global_best(k, ctx, n, offset, total_chains, pl.rank, pl.world, pl.device, pl.dist if pl.world > 1 else None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pkg.dist.global_best(k, ctx, n, offset, total_chains, pl.rank, pl.world, pl.device, pl.dist if pl.world > 1 else None)
    torch.cuda.synchronize()
    argbest_s = (time.perf_counter() - t0) / 5
    ctx.close()

    # ---- e2e: the reference-facing call with HOST buffers ------------------------------------------------------
    opts = dict(chain_offset=offset, total_chains=total_chains, lanes_per_chain=lanes, eval_mode=eval_mode)
    k.wrapper_ex(room, per_rank_chains, max(1, iterations // 100), seed=7, **opts)            # warm
    digest = None
    pl.barrier()
    t0 = time.perf_counter()
    for s in range(steps):
        res, pts, costs = k.wrapper_ex_raw(room, per_rank_chains, iterations, seed=7 + s, **opts)
        e2e_best = float(costs["totalCosts"].max())            # the caller reads the result in place ...
        if s == 0 and hash_chains and pl.rank == 0:
            m = min(hash_chains, per_rank_chains)
            digest = hashlib.sha256(pts[:m].tobytes() + costs[:m].tobytes()).hexdigest()[:16]
        k.free(res)                                            # ... and hands it back
    pl.barrier()
    e2e_wall = time.perf_counter() - t0
    wall_max, kernel_max, e2e_max, argbest_max = pl.gmax(wall, (ms1 - ms0) * 1e-3, e2e_wall, argbest_s)
    proposals = float(total_chains) * iterations * steps
    in_bytes = sum(a.nbytes for a in (room.rss, room.rsa, room.cfg, room.clearances, room.offlimits, room.vertices,
                                      room.surfaceRectangle, room.srf)) + 24
    return {"value": proposals / wall_max, "kernel_only": proposals / kernel_max, "e2e": proposals / e2e_max, "ms_per_step": wall_max * 1e3 / steps,
            "kernel_ms_per_launch": kernel_max * 1e3 / steps, "e2e_ms_per_step": e2e_max * 1e3 / steps, "argbest_ms": argbest_max * 1e3,
            "launches": int(l1 - l0), "clocks": clocks, "best": best, "shape": shape, "h2d_bytes_per_step": in_bytes,
            "d2h_bytes_per_step": per_rank_chains * n * 24 + per_rank_chains * 32, "hash": digest, "e2e_best": e2e_best}


def config4_strong(pl, k, pkg, flush, steps, warmup, clock_gpu=None):
    """BASELINE config 4 exactly as named: 262144 chains of the 200-object hall x 1000 iterations, the chains SHARDED
    over the N GPUs (strong scaling), NCCL arg-best.  Rank 0 holds global chains 0..1023 at every N <= 8, so the hash
    of their layouts and costs (e2e leg, fixed seed) is the 1-GPU vs N-GPU per-chain identity check (BASELINE section 5
    gate 4) on real hardware -- with the DEFAULT lane width (mhOptions.total_chains)."""
    total, iters = 262144, 1000
    room = pkg.synth.make_config(4)
    offset, count = pkg.dist.shard(total, pl.rank, pl.world)
    m = measure_job(pl, k, pkg, room, count, total, offset, iters, steps, warmup, flush, hash_chains=1024, clock_gpu=clock_gpu)
    out_total = total * room.n * 24 + total * 32
    return {"workload": f"config 4 ({room.name}): n={room.n} C={room.C} R={room.R}, {total} chains x {iters} iterations sharded over {pl.world} GPU(s)",
            "scaling": "strong", "chains_per_gpu": count, "steps": steps, "warmup": warmup, "value": m["value"], "kernel_only": m["kernel_only"],
            "e2e": m["e2e"], "unit": UNIT, "ms_per_step": m["ms_per_step"], "kernel_ms_per_launch": m["kernel_ms_per_launch"],
            "e2e_ms_per_step": m["e2e_ms_per_step"], "e2e_overhead_ms": m["e2e_ms_per_step"] - m["kernel_ms_per_launch"],
            "d2h_bytes_per_step_per_gpu": m["d2h_bytes_per_step"], "d2h_bytes_per_step_total": out_total, "argbest_ms": m["argbest_ms"],
            "hash_chains_0_1023": m["hash"], "lanes_per_chain": m["shape"]["lanes_per_chain"], "eval_form": m["shape"]["eval_form"],
            "best": {"global_chain": int(m["best"][0]), "totalCosts": float(m["best"][1])}, "clocks": m["clocks"],
            "e2e_limit": "what e2e adds to the kernel is e2e_overhead_ms: the D2H of this GPU's slice of the 1.26 GB result block into the caller's "
                         "malloc'd (pageable) memory, pre-faulted while the kernel runs; every rank copies its own slice concurrently",
            "_measure": m}


def c_abi_multi_gpu(pl, k, pkg, steps):
    """The jobs through ONE process and the C ABI's device list (mhOptions.devices): what the reference's caller, a
    single C# process calling KernelWrapper, can actually use.  Run by rank 0 over all N GPUs while the other ranks
    wait on a CPU-side barrier."""
    out = {}
    if pl.rank == 0:
        devs = list(range(pl.world))
        for name, cid, total, iters in (("config3_weak", 3, 65536 * pl.world, 10000), ("config4_strong", 4, 262144, 1000)):
            room = pkg.synth.make_config(cid)
            k.wrapper_ex(room, total, max(1, iters // 100), seed=7, devices=devs)                  # warm
            t0 = time.perf_counter()
            digest = None
            for s in range(steps):
                res, pts, costs = k.wrapper_ex_raw(room, total, iters, seed=7 + s, devices=devs)
                best = float(costs["totalCosts"].max())
                if s == 0:
                    digest = hashlib.sha256(pts[:1024].tobytes() + costs[:1024].tobytes()).hexdigest()[:16]
                k.free(res)
            dt = time.perf_counter() - t0
            out[name] = {"workload": f"config {cid}: {total} chains x {iters} iterations over devices {devs} in one process",
                         "e2e": total * iters * steps / dt, "unit": UNIT, "e2e_ms_per_step": dt * 1e3 / steps, "steps": steps,
                         "hash_chains_0_1023": digest, "best_totalCosts": best,
                         "call": "KernelWrapperEx(..., mhOptions{n_devices, devices[]}): one context + stream per device, every device "
                                 "copies its slice into the one malloc'd result block on a thread of its own"}
    pl.cpu_barrier()
    return out


def config2_as_named(k, pkg, with_reference=True):
    """BASELINE config 2 exactly as named: the 16-object bedroom, 1024 chains x 10000 iterations on one B200 (an
    under-filled machine: 1024 chains are 7 warps per SM at the default lane width), beside the reference kernel
    rebuilt for sm_100 on the SAME workload."""
    room = pkg.synth.make_config(2)
    chains, iters = 1024, 10000
    with k.create(room, chains, seed=1) as ctx:
        ctx.run(iters // 10)
        ctx.synchronize()
        ms0, _ = ctx.stats()
        ctx.reset()
        ctx.run(iters)
        ctx.synchronize()
        ms1, _ = ctx.stats()
        shape = ctx.shape()
    k.wrapper_ex(room, chains, 10, seed=3)
    t0 = time.perf_counter()
    reps = 3
    for s in range(reps):
        res, pts, costs = k.wrapper_ex_raw(room, chains, iters, seed=3 + s)
        k.free(res)
    e2e = chains * iters * reps / (time.perf_counter() - t0)
    out = {"workload": f"config 2 ({room.name}): n={room.n} C={room.C} R={room.R}, {chains} chains x {iters} iterations", "unit": UNIT,
           "kernel_only": chains * iters / ((ms1 - ms0) * 1e-3), "e2e": e2e, "lanes_per_chain": shape["lanes_per_chain"]}
    try:
        if with_reference and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libKernel_ref_nb.so")):
            v_wall, v_dev, _ = reference_gpu(2, chains, iters, 1, 1, timeout_s=120)
            r = reference_gpu.last
            out["ref_gpu_baseline"] = {"value": v_wall, "unit": UNIT, "device_events": v_dev, "without_init_rng": r.get("proposals_per_s_dev_without_init_rng"),
                                       "kind": "reference kernel rebuilt for sm_100 (divergent barrier neutralised), blockxDim=64, same GPU, same workload"}
            out["e2e_over_reference"] = e2e / v_wall
    except Exception as ex:
        out["ref_gpu_baseline"] = {"unavailable": str(ex)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --chains chains per GPU of --config (the headline); strong: BASELINE config 4 as named, 262144 chains sharded over the GPUs")
    ap.add_argument("--chains", type=int, default=65536, help="chains per GPU (weak scaling)")
    ap.add_argument("--iterations", type=int, default=10000)
    ap.add_argument("--ref-chains", type=int, default=8192, help="reference arm: 8192 blocks of 64 threads = 1.7 waves of 32 resident blocks per SM")
    ap.add_argument("--ref-iterations", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--eval-mode", type=int, default=0, choices=[0, 1, 2, 3],
                    help="mhOptions.eval_mode of the timed runs: 0 library default (memo form from 28 objects, bit-identical to 3), "
                         "3 plain scan (every term from scratch), 1 delta evaluation, 2 memo form")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (other rooms, config 4 strong, config 2 as named, in-process multi-GPU)")
    ap.add_argument("--seeds", type=int, default=3, help="config 5: seeds per sampler")
    ap.add_argument("--rungs", type=int, default=8, help="config 5: rungs per ladder (a multiple of --gpus)")
    ap.add_argument("--exchange-interval", type=int, default=100, help="config 5: iterations between exchange epochs")
    ap.add_argument("--sub-steps", type=int, default=2, help="timed steps of the sub-records (config 4 strong, in-process multi-GPU)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = importlib.import_module("metropolis-hastings-gpgpu_b200")
    tempering = args.config == 5
    if tempering:
        args.config = 3
    room = pkg.synth.make_config(args.config)

    if args.impl == "reference":
        run_reference_arm(args, room, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pl = Plumbing(torch, dist, rank, local_rank, world, device)
    k = pkg.Kernel()
    info = k.device_info()
    if tempering:
        run_tempering(args, pkg, k, room, pl)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2

    if args.scaling == "strong":
        c4 = config4_strong(pl, k, pkg, flush, args.steps, args.warmup, clock_gpu=local_rank if rank == 0 else None)
        if rank == 0:
            m = c4.pop("_measure")
            room4 = pkg.synth.make_config(4)
            peaks = measured_peaks()
            sm_max_mhz = float(peaks.get("sm_max_mhz") or info["sm_clock_khz"] / 1e3)
            peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
            per_gpu = m["kernel_only"] / world
            f_live = room4.flops_per_proposal(live=True)
            line = {"metric": "MH proposals evaluated/sec (chains x iters/s) at 200 objects, 262144 chains sharded", "value": m["value"], "unit": UNIT,
                    "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": c4["workload"], "parallelism": f"262144 chains sharded over {world} GPU(s) by global chain id, NCCL arg-best",
                               "l2": "256 MiB memset between steps; the chain state lives in shared memory"},
                    "e2e": {"value": m["e2e"], "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes_per_step"], "d2h_bytes_per_step": m["d2h_bytes_per_step"],
                            "call": "KernelWrapperEx per rank (host buffers in, malloc'd result block out)"},
                    "gpu_launches": m["launches"], "clocks": m["clocks"], "device": info["name"],
                    "roofline": {"bound": "fp32", "achieved": per_gpu * f_live / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                                 "frac": per_gpu * f_live / 1e12 / peak_tflops, "effective": True, "traffic": None,
                                 "kernel": "mh_delta_kernel<32, exact, 8>", "kernel_ms_per_launch": m["kernel_ms_per_launch"]},
                    "config4_strong": c4}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    n = room.n
    total_chains = args.chains * world
    offset = rank * args.chains
    m = measure_job(pl, k, pkg, room, args.chains, total_chains, offset, args.iterations, args.steps, args.warmup, flush, lanes=args.lanes,
                    eval_mode=args.eval_mode, clock_gpu=local_rank if rank == 0 else None)
    value, e2e_value, clocks, best = m["value"], m["e2e"], m["clocks"], m["best"]
    kernel_max = m["kernel_ms_per_launch"] * 1e-3 * args.steps

    extras = {}
    if not args.no_extras:
        sub = max(1, min(args.sub_steps, args.steps))
        c4 = config4_strong(pl, k, pkg, flush, sub, 1)
        c4.pop("_measure", None)
        extras["config4_strong"] = c4
        if world > 1:
            extras["c_abi_multi_gpu"] = c_abi_multi_gpu(pl, k, pkg, sub)

    if rank == 0:
        peaks = measured_peaks()
        sm_max_mhz = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or info["sm_clock_khz"] / 1e3)
        peak_tflops = info["sm_count"] * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        f_contract, f_live = room.flops_per_proposal(), room.flops_per_proposal(live=True)
        per_gpu_rate = float(args.chains) * args.iterations * args.steps / kernel_max
        achieved = per_gpu_rate * f_live / 1e12
        scan_rate = quick_rate(k, room, args.chains, max(200, min(2000, args.iterations)), eval_mode=3, lanes_per_chain=args.lanes)
        roofline = {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                    "effective": args.eval_mode != 3 and n >= 28,
                    "effective_note": "algorithmic flops of a full evaluation x proposals/s: the default kernel returns the full evaluation's "
                                      "bits but executes a fraction of its arithmetic (exact memos); the plain scan, which executes all of it, "
                                      "is in full_scan",
                    "full_scan": {"value": scan_rate, "unit": UNIT, "achieved": scan_rate * f_live / 1e12,
                                  "frac": scan_rate * f_live / 1e12 / peak_tflops, "kernel": "mh_chain_kernel (MH_EVAL_FULL_SCAN)"},
                    "traffic": profiled_dram_traffic(),
                    "traffic_note": f"bytes per launch at 65536 chains from profiles/{PROFILE_FILE} (result block; independent of the iteration count)",
                    "flops_per_proposal": {"live": f_live, "contract": f_contract},
                    "achieved_contract": per_gpu_rate * f_contract / 1e12, "frac_contract": per_gpu_rate * f_contract / 1e12 / peak_tflops,
                    "kernel": {0: f"mh_delta_kernel<{m['shape']['lanes_per_chain']}, exact> (MH_EVAL_FULL in its memo form)" if n >= 28 else "mh_chain_kernel",
                               1: "mh_delta_kernel<., delta> (MH_EVAL_DELTA)", 2: "mh_delta_kernel<., exact> (MH_EVAL_MEMO)",
                               3: "mh_chain_kernel (MH_EVAL_FULL_SCAN)"}[args.eval_mode],
                    "kernel_ms_per_launch": kernel_max * 1e3 / args.steps,
                    "peak_source": f"FP32 pipe = {info['sm_count']} SMs x 128 lanes x 2 x {sm_max_mhz:.0f} MHz (sm_max_mhz of "
                                   + ("MEASURED_PEAKS.json" if peaks.get("sm_max_mhz") else "nvidia-smi") + "); HBM is not the bound: "
                                   "chain state lives in shared memory",
                    "frac_at_observed_clock": (achieved / (info["sm_count"] * 256 * clocks["sm_mhz"] * 1e6 / 1e12)) if clocks and clocks.get("sm_mhz") else None}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"config {args.config} ({room.name}): n={n} C={room.C} R={room.R}, {args.chains} chains/GPU x {args.iterations} iterations, all cost terms, beta=2",
                           "evaluation": {0: "MH_EVAL_FULL (library default): every proposal's full cost, bit-identical to the plain re-evaluation, computed through exact memos",
                                          1: "MH_EVAL_DELTA: incremental running sums, statistically equivalent to the full evaluation",
                                          2: "MH_EVAL_MEMO: the full evaluation's bits through exact memos",
                                          3: "MH_EVAL_FULL_SCAN: every live cost term of every proposal from scratch"}[args.eval_mode],
                           "chains_total": total_chains, "lanes_per_chain": m["shape"]["lanes_per_chain"], "parallelism": f"chains sharded over {world} GPU(s), NCCL arg-best",
                           "l2": "256 MiB memset between steps; the chain state lives in shared memory"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m["h2d_bytes_per_step"], "d2h_bytes_per_step": m["d2h_bytes_per_step"],
                        "call": "KernelWrapperEx (host buffers in, malloc'd result block out)"},
                "gpu_launches": m["launches"], "roofline": roofline, "clocks": clocks, "argbest_ms": m["argbest_ms"],
                "best": {"global_chain": int(best[0]), "totalCosts": float(best[1])}, "device": info["name"]}
        line.update(extras)
        if not args.no_extras and world == 1:                    # the side measurements run at N=1 only
            line["config2_as_named"] = config2_as_named(k, pkg, with_reference=not args.no_ref_gpu)
            line["other_configs_kernel_only"] = other_configs(k, pkg)
            line["config3_annealing"] = config3_annealing(k, room, args.chains, min(args.iterations, 10000))
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(room)
            line["cpu_baseline"] = cb
            # time-to-best-cost: how long until the GPU run's global best reaches what the CPU port's chains
            # reached (their best, and the median of their finals) in its cb["seconds"] of host time
            ttb = {"cpu_seconds": cb["seconds"],
                   "note": "targets = totalCosts reached by the CPU port's sample (best of its chains / median of its chains' finals) in "
                           "cpu_seconds of host time; GPU: wall time from KernelCreate until the device arg-max over the chains reaches "
                           "it, checked every 200 iterations.  Reaching a given cost is a matter of iterations per chain, so fewer "
                           "chains (wider lane groups, an under-filled but low-latency machine) get there first"}
            for name in ("best_totalCosts", "median_final_totalCosts"):
                ttb["to_cpu_" + name] = {"target": cb[name],
                                         "runs": [time_to_best_cost(k, room, ch, cb[name], epoch=200) for ch in (1024, 4096, args.chains)]}
                ttb["to_cpu_" + name]["seconds"] = min(r["seconds"] for r in ttb["to_cpu_" + name]["runs"] if r["reached"]) \
                    if any(r["reached"] for r in ttb["to_cpu_" + name]["runs"]) else None
            line["time_to_best_cost"] = ttb
        if not args.no_ref_gpu and world == 1:
            try:
                if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libKernel_ref_nb.so")):
                    v_wall, v_dev, _ = reference_gpu(args.config, args.ref_chains, args.ref_iterations, 1, 1, timeout_s=120)
                    r = reference_gpu.last
                    line["ref_gpu_baseline"] = {"value": v_wall, "unit": UNIT, "kind": "reference kernel rebuilt for sm_100 (divergent barrier neutralised), blockxDim=64, same GPU",
                                                "device_events": v_dev, "without_init_rng": r.get("proposals_per_s_dev_without_init_rng"),
                                                "sample": f"{args.ref_chains} chains x {args.ref_iterations} iterations, whole KernelWrapper call; device-event rate {v_dev:.4g}/s"}
            except Exception as ex:
                line["ref_gpu_baseline"] = {"unavailable": str(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
