#!/usr/bin/env python
"""Benchmark of the hot path: parallel-tempering MCMC steps per second with every temperature of
the ladder advanced (BASELINE.json's metric), on synthetic `simulate_data` of the named shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" is everything MCMC::burnin(step) does (src/mcmc.cpp:157-188): the eight Metropolis
updates on every chain of the ladder, the per-chain log-likelihood reduction, the swap pass and
(when due) the ladder adaptation.  With N > 1 (launched under torchrun, one rank per GPU) the
ladder is cut into N contiguous blocks of temperatures and the ranks exchange only the per-chain
scalar log-likelihoods (NCCL all-gather); total work is fixed, so scaling is "strong".

Our arm prints ONE JSON line: `value` = steps/s with the state resident in HBM (CUDA events on the
engine's stream around each step, L2 flushed between steps, max over ranks), `e2e` = the same
metric through the public `run_mcmc` call with host buffers (data packing, upload, chain
initialisation, the batched read-back of traces, ladder and stored draws, result assembly), `roofline` for the dominant kernel, `cpu_baseline` = the reference's own C++ sampler
(compiled unmodified under oracle/_ref) on this box's host cores.  The timed steps run as the
product runs them (the engine replays its captured step graph); the per-kernel numbers behind
`roofline` and `kernel_ms_per_step` come from K further steps issued as plain launches with an
event pair around every update kind and every evaluator launch.  With N > 1 the exchange is the
C++ driver's own ncclAllGather on the engine's stream inside the step graph (`exchange` says which
path ran), the swap pass runs on the device, every rank's ladder is checked against rank 0's
(`ladder_consistent`), and the line carries an `other_workloads` block with the multi-GPU shapes of
BASELINE.json (c4; c5 at 16/32/64 temperatures) timed the same way on the same ranks.

`--impl reference` times that CPU sampler alone and prints the same line with "impl": "reference".
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mcmc_steps_per_sec_all_temps"
UNIT = "steps/s"
B_EVAL = 24  # algorithmic bytes per cell evaluation (obs mask + latent mask + llik)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.reasons, self.max_mhz, self.err = [], set(), None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.stop_flag.wait(0.05)
        except Exception as exc:  # no NVML: report that, do not fail the bench
            self.err = repr(exc)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.err}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---- the reference's CPU sampler (oracle/_ref, built from /root/reference by oracle/Makefile) ----
def cpu_reference_rate(wl, steps, warmup, budget_s):
    """MCMC::burnin steps/s (all temperatures) of the unmodified reference C++ on this host, all
    OpenMP threads.  If the full ladder does not fit the time budget the step is run on a
    sub-ladder (every k-th temperature) and scaled by chains, and `sample` says so."""
    from oracle.pyoracle import Ref
    ref = Ref(32)
    pk, temps = wl["packed"], np.asarray(wl["temps"], dtype=np.float64)
    T = len(temps)
    cores = os.cpu_count() or 1
    threads = min(cores, T)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)

    evals = [0]

    def run(sub_temps, w, k):
        ref.set_params(pt_chains=sub_temps, pt_num_threads=min(cores, len(sub_temps)),
                       allow_relatedness=int(wl["model"]["allow_relatedness"]),
                       max_coi=wl["model"]["max_coi"], burnin=10000)
        h = ref.mcmc_new(seed=1)
        t0 = time.time()
        ref.eval_count_reset()
        one = ref.mcmc_time_burnin(h, 0, 0, 1)        # first step, also the estimate
        evals[0] = ref.eval_count()
        if w + k > 1 and one * (w + k) > budget_s:
            ref.mcmc_free(h)
            return None, one, time.time() - t0
        for i in range(max(w - 1, 0)):
            ref.mcmc_time_burnin(h, 1 + i, 0, 1)
        ref.eval_count_reset()
        secs = ref.mcmc_time_burnin(h, max(w, 1), 0, k)
        evals[0] = ref.eval_count()
        ref.mcmc_free(h)
        return secs, one, time.time() - t0

    secs, one, _ = run(temps, warmup, steps)
    used_T = T
    if secs is None:
        # sub-ladder sized to the budget: chains cost about the same, threads <= chains
        per_chain_round = one / math.ceil(T / threads)
        rounds = max(1, int(budget_s / ((warmup + steps) * per_chain_round)))
        used_T = max(2, min(T, rounds * threads))
        idx = np.unique(np.round(np.linspace(0, T - 1, used_T)).astype(int))
        used_T = len(idx)
        secs, one, _ = run(temps[idx], warmup, steps)
        if secs is None:   # still too slow: one timed step after the estimate step
            secs, steps = one, 1
    chain_steps_per_s = used_T * steps / secs
    value = chain_steps_per_s / T
    sample = (f"{steps} MCMC::burnin steps after {warmup} warm-up, "
              f"{used_T} of {T} temperatures" + ("" if used_T == T else " (every k-th; scaled by chains)") +
              f", {min(cores, used_T)} OpenMP threads, reference C++ (-O2 -fopenmp) via oracle/_ref/libmoire_ref32.so")
    # non-missing Chain::calculate_genotype_likelihood calls of the timed steps, counted inside the
    # reference (oracle/ref_harness.cpp) -- the second half of BASELINE.json's metric
    return dict(value=value, unit=UNIT, cores=min(cores, used_T), kind="reference", sample=sample,
                host_cores=cores, seconds=secs, likelihood_evals_per_sec=evals[0] / secs,
                evals_per_step_all_temps=evals[0] / steps * T / used_T)


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_rate(wl, args.steps, args.warmup, budget_s=240.0)
    line = dict(impl="reference", metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 / base["value"],
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                data="synthetic", config=config_of(wl, args),
                parallelism=f"{base['cores']} OpenMP threads over the chains (host CPU; no GPU)",
                cpu_baseline=dict(value=base["value"], unit=UNIT, cores=base["cores"],
                                  kind=base["kind"], sample=base["sample"],
                                  likelihood_evals_per_sec=base["likelihood_evals_per_sec"]),
                likelihood_evals_per_sec=base["likelihood_evals_per_sec"],
                evals_per_step=base["evals_per_step_all_temps"],
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def config_of(wl, args):
    pk = wl["packed"]
    # the workload only: both arms (ours, --impl reference) print the same dict
    return dict(workload=f"{wl['name']}: simulate_data {pk.num_samples} samples x {pk.num_loci} loci, "
                         f"A_j {int(pk.num_alleles.min())}..{int(pk.num_alleles.max())}, "
                         f"{len(wl['temps'])} temperatures, relatedness on, max_coi {wl['model']['max_coi']}",
                samples=pk.num_samples, loci=pk.num_loci, temperatures=len(wl["temps"]))


# ---- our arm ---------------------------------------------------------------------------------------
def ladder_digest(mc):
    """What every rank must agree on after the same steps: swap_store, the index map, the
    temperatures by chain, the adapted gradient, the barrier sums and the traces (bit patterns)."""
    import hashlib
    lad, tr = mc.ladder(), mc.traces()
    h = hashlib.sha256()
    for k in ("swap_store", "swap_indices", "chain_temps", "temp_gradient", "swap_barriers", "swap_acceptances"):
        h.update(np.ascontiguousarray(lad[k]).tobytes())
    for k in ("llik_burnin", "prior_burnin", "posterior_burnin"):
        h.update(np.ascontiguousarray(tr[k]).tobytes())
    return int.from_bytes(h.digest()[:6], "little")      # 48 bits: exact in a float64


def timed_workload(name, T, steps, warmup, world, rank, local, dev, seed):
    """steps/s of another BASELINE.json shape on the ranks of this run, timed like the main one
    (CUDA events on the engine's stream, L2 flush before every timed step, max over ranks)."""
    import torch
    import torch.distributed as dist
    from moire_b200.mcmc import MCMC
    from moire_b200.workloads import ladder, make_workload
    wl = make_workload(name)
    temps = ladder(T)
    mc = MCMC(wl["packed"], temps, burnin=warmup + steps + 4, samples=0, seed=seed, device=local, rank=rank,
              world_size=world, **wl["model"])
    try:
        if world > 1 and not mc.init_nccl(local):
            return dict(error="NCCL exchange unavailable")
        if not mc.init():
            return dict(error="chain initialisation failed")
        eng = mc.engine
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        step = 0
        for _ in range(warmup):
            mc.burnin(step)
            step += 1
        ev0, _ = eng.counters()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            eng.timer_start()
            mc.burnin(step)
            ms += eng.timer_stop()
            step += 1
        ev1, _ = eng.counters()
        digest = ladder_digest(mc)
        st = torch.tensor([ms, float(ev1 - ev0), float(digest)], dtype=torch.float64, device=dev)
        consistent = True
        if world > 1:
            allst = [torch.empty_like(st) for _ in range(world)]
            dist.all_gather(allst, st)
            ms_all = [float(x[0]) for x in allst]
            evals = sum(float(x[1]) for x in allst)
            consistent = all(float(x[2]) == float(allst[0][2]) for x in allst)
        else:
            ms_all, evals = [ms], float(ev1 - ev0)
        ms_step = max(ms_all) / steps
        pk = wl["packed"]
        return dict(workload=f"{name}: {pk.num_samples} x {pk.num_loci}, A_j {int(pk.num_alleles.min())}.."
                             f"{int(pk.num_alleles.max())}, {T} temperatures, max_coi {wl['model']['max_coi']}",
                    n_gpus=world, steps=steps, warmup=warmup, value=1e3 / ms_step, unit=UNIT,
                    ms_per_step=ms_step, per_rank_ms_per_step=[round(m / steps, 4) for m in ms_all],
                    likelihood_evals_per_sec=evals / (max(ms_all) * 1e-3), ladder=mc.mode(),
                    ladder_consistent=consistent)
    finally:
        mc.close()


def our_arm(args, wl):
    import torch
    import torch.distributed as dist
    from moire_b200.mcmc import MCMC, run_mcmc, torch_allgather

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL reports its communicators (nranks, transports, NVLS) at INFO; fd 1 was re-pointed at
        # stderr in main(), so none of it can reach the one JSON line on stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("MOIRE_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
    pk, temps, model = wl["packed"], np.asarray(wl["temps"], dtype=np.float64), wl["model"]
    T = len(temps)
    if world > T:
        raise SystemExit(f"{world} ranks for a ladder of {T} temperatures")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gather = torch_allgather(world, dev) if world > 1 else None
    burnin_total = args.warmup + 2 * args.steps + 8
    mc = MCMC(pk, temps, burnin=burnin_total, samples=0, seed=args.seed, device=local, rank=rank,
              world_size=world, allgather=gather, **model)
    exchange = "none (one rank)"
    if world > 1:
        native = not os.environ.get("MOIRE_B200_CALLBACK_EXCHANGE") and mc.init_nccl(local)
        exchange = ("ncclAllGather on the engine stream from device scalars (C++ driver)" if native
                    else "torch.distributed all_gather through the host callback")
    if not mc.init():
        raise SystemExit("chain initialisation failed")
    eng = mc.engine
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    step = 0
    for _ in range(args.warmup):
        mc.burnin(step)
        step += 1
    # ---- the timed region: K steps exactly as the product runs them (the engine replays its
    # captured step graph), device time from CUDA events on the engine's stream around each step
    ev0, l0 = eng.counters()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        eng.timer_start()
        mc.burnin(step)
        dev_ms += eng.timer_stop()
        step += 1
    barrier()
    wall = time.perf_counter() - wall0
    sampler.stop_flag.set()
    sampler.join(2.0)
    ev1, l1 = eng.counters()
    ladder_mode = mc.mode()
    digest = ladder_digest(mc)          # also the read-back of everything the steps left queued
    # ---- breakdown pass: the next K steps with an event pair around every update kind and every
    # evaluator launch (a graph cannot hold events between its kernels, so these steps go out as
    # plain launches); feeds `roofline` and `kernel_ms_per_step`, not `value`
    ke0 = eng.kind_evals()
    eng.set_timing(True)
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        mc.burnin(step)
        step += 1
    kinds = eng.timing()
    eval_ms, eval_n = eng.evaluator_timing()
    eng.set_timing(False)
    ke1 = eng.kind_evals()
    barrier()
    stats = torch.tensor([dev_ms, float(ev1 - ev0), float(l1 - l0)], dtype=torch.float64, device=dev)
    ladder_consistent = True
    per_rank_ms = [dev_ms / args.steps]
    if world > 1:
        dg = torch.tensor([float(digest), dev_ms / args.steps], dtype=torch.float64, device=dev)
        alld = [torch.empty_like(dg) for _ in range(world)]
        dist.all_gather(alld, dg)
        ladder_consistent = all(float(x[0]) == float(alld[0][0]) for x in alld)
        per_rank_ms = [float(x[1]) for x in alld]
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dev_ms_max = float(mx[0])
    else:
        dev_ms_max = dev_ms
    evals_total, launches_total = float(stats[1]), int(stats[2])
    ms_per_step = dev_ms_max / args.steps
    value = 1e3 / ms_per_step
    mc.close()

    # ---- end to end through the public call, host buffers in, result list out: packing, upload,
    # engine creation, chain initialisation, every step's scalar read-back and swap pass, the
    # stored draws and the assembly of the result list are all inside the timed region.  200 steps
    # at the default K: still 55x shorter than a default run (10000 burn-in + 1000 sampling steps),
    # so the one-off costs weigh more here than they do for a user
    e2e_burn, e2e_samp = min(max(15 * args.steps, 150), 1500), min(max(5 * args.steps, 50), 500)
    barrier()
    t0 = time.perf_counter()
    res = run_mcmc(wl["sim"], wl["sim"]["is_missing"], burnin=e2e_burn, samples_per_chain=e2e_samp,
                   pt_chains=temps, verbose=False, seed=args.seed, device=local,
                   distributed=world > 1, **{k: v for k, v in model.items()})
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    chain = res["chains"][0]
    n_steps = e2e_burn + e2e_samp
    N, L, AP = pk.num_samples, pk.num_loci, (int(pk.num_alleles.max()) + 3) // 4 * 4
    # host -> device: the data and the ladder once per rank, two flag words per step, the ladder
    # again after each adaptation; device -> host: per step its trace entry (3 doubles + the
    # swap_store word), the ladder state once per read-back batch (<= 32 steps), the stored draws
    ladder_bytes = T * (3 * 8 + 2 * 4 + 8 + 8) + 64
    h2d = (pk.obs_mask.nbytes + pk.is_missing.nbytes + pk.num_alleles.nbytes + pk.observed_coi.nbytes
           + ladder_bytes) * world / n_steps + 16 + ladder_bytes / 25
    draw_bytes = L * AP * 8 + N * (4 + 4 * 8) + 8 + 4
    d2h = 28 + ladder_bytes / 25 + draw_bytes * len(chain["lam_coi"]) / n_steps
    e2e = dict(value=n_steps / e2e_s, unit=UNIT, h2d_bytes_per_step=int(h2d),
               d2h_bytes_per_step=int(d2h), seconds=e2e_s, steps=n_steps,
               call=f"moire_b200.mcmc.run_mcmc(burnin={e2e_burn}, samples_per_chain={e2e_samp}, "
                    f"pt_chains={T}) incl. pack_data, upload, chain init, batched read-back of traces / "
                    f"ladder / {len(chain['lam_coi'])} stored draws, host-side ladder adaptation, result assembly")

    # ---- the multi-GPU shapes of BASELINE.json on the same ranks (configs[3], configs[4])
    other = {}
    if world > 1 and wl["name"] == "c3" and not args.no_other_workloads:
        # ... and the benchmark shape with the ladder grown with the box (40 temperatures per GPU): the
        # headline line above cuts a fixed 40-temperature ladder N ways (strong scaling); this one
        # keeps every GPU's share fixed (weak scaling); chain_steps_per_sec = steps/s x temperatures
        shapes = [("c4", "c4", 64, 5), ("c5_T16", "c5", 16, 5), ("c5_T32", "c5", 32, 5), ("c5_T64", "c5", 64, 5),
                  (f"c3_weak_T{40 * world}", "c3", 40 * world, 30)]
        for key, name, T, nsteps in shapes:
            if T < world:
                continue
            try:
                other[key] = timed_workload(name, T, steps=nsteps, warmup=3, world=world, rank=rank, local=local,
                                            dev=dev, seed=args.seed)
                other[key]["chain_steps_per_sec"] = other[key]["value"] * T
            except Exception as exc:  # a shape that does not fit is reported, not fatal
                other[key] = dict(error=repr(exc))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (this rank's launches, CUDA events on its stream)
    peak, peak_src = measured_peak()
    kind_ms = {k: v[0] for k, v in kinds.items()}
    kind_n = {k: v[1] for k, v in kinds.items()}
    step_kernel_ms = max(sum(kind_ms.values()), 1e-12)
    if eval_n > 0:
        # the flat pipelines: one persistent evaluator kernel does every transmission term of the
        # p, m, r, eff_coi and samples updates; it is timed alone, launch by launch
        top, top_ms, top_n = "k_flat_eval", eval_ms, eval_n
        top_evals = sum(ke1[k] - ke0[k] for k in ("p", "m", "r", "eff_coi", "samples"))
    else:
        # tiny problems run the fused per-kind kernels: the dominant update kind is the kernel
        kind = max((k for k in kind_ms if k != "reduce"), key=lambda k: kind_ms[k])
        top = {"p": "k_update_p"}.get(kind, f"k_update_sample<{kind}>")
        top_ms, top_n = kind_ms[kind], kind_n[kind]
        top_evals = ke1[kind] - ke0[kind] if kind in ke1 else 0
    achieved = top_evals * B_EVAL / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
    traffic, compute = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_inputs.json")) as f:
            inputs = json.load(f)
        traffic = inputs.get("traffic_bytes_per_launch", {}).get(top)
        compute = inputs.get("compute", {}).get(top)
    except Exception:
        pass
    if compute is not None:
        # what actually bounds the kernel: the fp64 pipe and the issue slots, from the
        # committed ncu capture; the live part is the instruction rate these launches sustained
        compute = dict(compute, source="profiles/roofline_inputs.json (ncu --set full capture of one full-ladder launch)",
                       measured_thread_instructions_per_sec=(compute["thread_instructions_per_evaluation"] * top_evals
                                                             / (top_ms * 1e-3)) if top_ms > 0 else None)
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=traffic, compute=compute, kernel=top, peak_source=peak_src,
                    launches=top_n, ms_per_launch=top_ms / max(top_n, 1),
                    algorithmic_bytes_per_launch=top_evals * B_EVAL / max(top_n, 1),
                    share_of_step=top_ms / step_kernel_ms,
                    note="algorithmic fraction: 24 B per cell evaluation; the kernel is fp64-issue/"
                         "latency bound (SURVEY F11), see DESIGN.md; traffic is the ncu capture of "
                         "one full-ladder launch of this kernel (profiles/roofline_inputs.json); "
                         "launch durations from the breakdown pass (K steps after the timed region, "
                         "plain launches with an event pair per evaluator launch)")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_rate(wl, steps=2, warmup=1, budget_s=30.0)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "likelihood_evals_per_sec")}
        except Exception as exc:
            cpu = dict(value=None, unit=UNIT, cores=0, kind="reference", sample=f"unavailable: {exc!r}")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=config_of(wl, args), exchange=exchange,
                parallelism=f"temperature ladder sharded over {world} GPU(s)",
                cache="L2 flushed (256 MiB memset) before every timed step",
                ladder=ladder_mode, ladder_consistent=ladder_consistent, per_rank_ms_per_step=[round(x, 4) for x in per_rank_ms],
                clocks=sampler.summary(), e2e=e2e,
                gpu_launches=launches_total, roofline=roofline, cpu_baseline=cpu,
                likelihood_evals_per_sec=evals_total / (dev_ms_max * 1e-3),
                evals_per_step=evals_total / args.steps,
                wall_ms_per_step_incl_flush=1e3 * wall / args.steps,
                kernel_ms_per_step={k: round(v / args.steps, 4) for k, v in kind_ms.items()})
    if other:
        line["other_workloads"] = other
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    # stdout carries exactly one JSON line: everything else that libraries write to fd 1 (NCCL's
    # version banner at communicator creation, for one) goes to stderr
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="N > 1: skip the c4 / c5 block timed after the main workload")
    ap.add_argument("--temperatures", type=int, default=0,
                    help="tuning runs: a shorter ladder of the workload's shape (what one GPU of a sharded "
                         "ladder holds); the config line says so")
    args = ap.parse_args()
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return
    from moire_b200.workloads import ladder, make_workload
    wl = make_workload(args.workload)
    if args.temperatures > 0:
        wl["temps"] = ladder(args.temperatures)
    if args.impl == "reference":
        reference_arm(args, wl)
    else:
        if args.warmup < 3:
            print(f"note: --warmup {args.warmup} < 3", file=sys.stderr)
        our_arm(args, wl)


if __name__ == "__main__":
    main()
