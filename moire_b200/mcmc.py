"""``run_mcmc`` and the parallel-tempering driver object, mirroring the reference's R entry point
(R/mcmc.R:76-182) and its C++ ``MCMC`` class (src/mcmc.h).

The driver itself is C++ (moire_b200/csrc/mcmc_driver.cpp, C ABI in include/moire_mcmc.h); this
module only marshals arguments and results, and -- when the temperature ladder is sharded over
several GPUs -- supplies the one exchange the path has, an all-gather of per-chain scalar
log-likelihoods, through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time

import numpy as np

from . import _lib
from .data import PackedData, pack_data
from .engine import DEFAULT_MODEL, Engine, MoireError, _ptr, pipeline_flags

PRIOR_TERMS = ("eps_neg", "eps_pos", "relatedness", "coi", "mean_coi_hyper")


class InitializationFailure(MoireError):
    """The reference's ``moire_initialization_failure`` condition (R/initialization_diagnostics.R)."""

    def __init__(self, message, diagnostics):
        super().__init__(message)
        self.diagnostics = diagnostics


def _ladder(pt_chains, pt_grad):
    # R/mcmc.R:139-141
    if np.ndim(pt_chains) == 0:
        T = int(pt_chains)
        return np.array([1.0]) if T == 1 else np.linspace(1.0, 0.0, T) ** pt_grad
    return np.asarray(pt_chains, dtype=np.float64)


def classify_initialization(records, prior_counts, max_tries, chains_attempted, per_chain,
                            majority_threshold=0.8):
    """InitializationDiagnostics::classify / to_list (src/initialization_diagnostics.h:62-196) on
    the engine's per-attempt failure records [(chain, sample, locus)], 1-based, 0 = unknown."""
    locus_counts, sample_counts, examples, unknown = {}, {}, [], 0
    for _, s, l in records:
        if s > 0 and l > 0:
            sample_counts[s] = sample_counts.get(s, 0) + 1
            locus_counts[l] = locus_counts.get(l, 0) + 1
            if len(examples) < 5:
                examples.append((s, l))
        else:
            unknown += 1
    total = len(records)
    d = dict(max_initialization_tries=max_tries, chains_attempted=chains_attempted,
             total_ill_conditioned=total, unknown_genotyping_count=unknown,
             locus_counts=dict(locus=sorted(locus_counts), n=[locus_counts[k] for k in sorted(locus_counts)]),
             sample_counts=dict(sample=sorted(sample_counts), n=[sample_counts[k] for k in sorted(sample_counts)]),
             prior_nonfinite_counts=dict(term=[t for t in sorted(prior_counts) if prior_counts[t]],
                                         n=[prior_counts[t] for t in sorted(prior_counts) if prior_counts[t]]),
             examples=dict(sample=[e[0] for e in examples], locus=[e[1] for e in examples]),
             ill_conditioned_per_chain=list(per_chain), classification="hard_starting_set",
             concentration="", dominant_locus=None, dominant_sample=None, dominant_share=0.0,
             secondary_sample=None)
    if total <= 0:
        return d

    def best(counts):  # first maximum in key order, like the std::map scan
        bk, bc = None, 0
        for k in sorted(counts):
            if counts[k] > bc:
                bk, bc = k, counts[k]
        return bk, bc

    bl, blc = best(locus_counts)
    bs, bsc = best(sample_counts)
    ls, ss = blc / total, bsc / total
    if blc > 0 and ls >= majority_threshold:
        d.update(classification="consistent_failure_cause", concentration="locus",
                 dominant_locus=bl, dominant_share=ls)
        if bsc > 0 and ss >= majority_threshold:
            d["secondary_sample"] = bs
    elif bsc > 0 and ss >= majority_threshold:
        d.update(classification="consistent_failure_cause", concentration="sample",
                 dominant_sample=bs, dominant_share=ss)
    else:
        d["dominant_share"] = max(ls, ss)
        if blc >= bsc:
            d["dominant_locus"] = bl
        else:
            d["dominant_sample"] = bs
    return d


def format_initialization_failure_message(diagnostics, sample_ids=None, loci=None):
    """R/initialization_diagnostics.R:34-112."""
    def loc(index, names):
        if index is None:
            return None
        if names is not None and 1 <= index <= len(names):
            return f"{index} ({names[index - 1]})"
        return str(index)

    d = diagnostics
    lines = ["Initialization failed: no finite genotyping log-likelihood after Initialization retries.",
             f"Attempted {d.get('chains_attempted')} chain(s); recorded {d.get('total_ill_conditioned', 0)} "
             f"Ill-conditioned start(s) (max_initialization_tries = {d.get('max_initialization_tries')})."]
    share = d.get("dominant_share", float("nan"))
    if d.get("classification") == "consistent_failure_cause":
        if d.get("concentration") == "locus":
            lines.append(f"Consistent failure cause: locus {loc(d['dominant_locus'], loci)} in "
                         f"{100 * share:.0f}% of Ill-conditioned starts.")
            lines.append("Guidance: inspect this locus; high allelic diversity or extremely rare alleles "
                         "often destabilize the genotyping likelihood. Consider dropping rare alleles or "
                         "this locus.")
            sec = loc(d.get("secondary_sample"), sample_ids)
            if sec is not None:
                lines.append(f"Secondary concentration on sample {sec} was also observed.")
        elif d.get("concentration") == "sample":
            lines.append(f"Consistent failure cause: sample {loc(d['dominant_sample'], sample_ids)} in "
                         f"{100 * share:.0f}% of Ill-conditioned starts.")
            lines.append("Guidance: inspect this sample's missingness or call patterns; consider "
                         "excluding it for a sanity run.")
    else:
        lines.append("Classification: hard starting set (no single locus or sample dominates failures).")
        lines.append("Guidance: increase max_initialization_tries; if it still fails, simplify the "
                     "dataset (fewer loci/alleles) and retry.")
    ex = d.get("examples", {})
    if ex.get("sample") and ex.get("locus"):
        bits = [f"sample {loc(s, sample_ids)}, locus {loc(l, loci)}"
                for s, l in zip(ex["sample"], ex["locus"])]
        lines.append("Example Failure loci: " + "; ".join(bits))
    pr = d.get("prior_nonfinite_counts", {})
    if pr.get("term"):
        lines.append("Secondary non-finite prior terms observed: " +
                     ", ".join(f"{t}={n}" for t, n in zip(pr["term"], pr["n"])))
    return "\n".join(lines)


class MCMC:
    """One parallel-tempering run: this process's block of the ladder on one GPU."""

    def __init__(self, data: PackedData, pt_chains, *, burnin, samples, thin=1, adapt_temp=True,
                 pre_adapt_steps=25, temp_adapt_steps=25, max_initialization_tries=10000,
                 record_latent_genotypes=False, max_runtime=math.inf, seed=1, device=0, rank=0,
                 world_size=1, allgather=None, pipeline="auto", ref32_emulation=False, **model):
        self._h = None
        self.W = data.mask_words   # two-word masks (libmoire_b200_w128.so) past 64 alleles per locus
        self.lib = _lib.load(self.W)
        prm = dict(DEFAULT_MODEL)
        unknown = set(model) - set(prm)
        if unknown:
            raise TypeError(f"unknown model parameters: {sorted(unknown)}")
        prm.update(model)
        prm["burnin"] = int(burnin)
        self.model, self.data = prm, data
        self.N, self.L = data.num_samples, data.num_loci
        self.temps0 = np.ascontiguousarray(np.asarray(pt_chains, dtype=np.float64))
        self.T = len(self.temps0)
        self.rank, self.world = rank, world_size
        self.max_tries = int(max_initialization_tries)
        self.record_latent = bool(record_latent_genotypes)
        self._keep = dict(
            na=np.ascontiguousarray(data.num_alleles, dtype=np.int32),
            obs=np.ascontiguousarray(data.obs_mask, dtype=np.uint64),
            mis=np.ascontiguousarray(data.is_missing, dtype=np.uint8),
            coi=np.ascontiguousarray(data.observed_coi, dtype=np.int32))
        d = _lib.MoireData(self.N, self.L, _ptr(self._keep["na"]), _ptr(self._keep["obs"]),
                           _ptr(self._keep["mis"]), _ptr(self._keep["coi"]))
        p = _lib.MoireParams(int(bool(prm["allow_relatedness"])), int(burnin), int(prm["max_coi"]),
                             pipeline_flags(pipeline, ref32_emulation),
                             prm["mean_coi_shape"], prm["mean_coi_scale"], prm["max_eps_pos"],
                             prm["eps_pos_alpha"], prm["eps_pos_beta"], prm["max_eps_neg"],
                             prm["eps_neg_alpha"], prm["eps_neg_beta"], prm["r_alpha"],
                             prm["r_beta"])
        mr = float(max_runtime)
        mp = _lib.MoireMcmcParams(int(thin), int(burnin), int(samples), int(bool(adapt_temp)),
                                  int(pre_adapt_steps), int(temp_adapt_steps), self.max_tries,
                                  int(self.record_latent), mr if math.isfinite(mr) else -1.0)
        h = C.c_void_p()
        rc = self.lib.moire_mcmc_create(C.byref(d), C.byref(p), C.byref(mp), _ptr(self.temps0),
                                        self.T, rank, world_size, device, seed, C.byref(h))
        if rc != 0:
            raise MoireError(f"moire_mcmc_create failed ({rc}): "
                             f"{self.lib.moire_mcmc_last_error(None).decode()}")
        self._h = h
        self._cb = None
        if allgather is not None:
            self.set_allgather(allgather)
        # a borrowed Engine view of this rank's chains (state dumps, diagnostics, counters);
        # None in ladder-only mode (device < 0: host-logic tests, no evaluation)
        eh = self.lib.moire_mcmc_engine(h)
        self.engine, self.a_pad = None, 4
        if eh:
            self.engine = Engine.__new__(Engine)
            self.engine.lib, self.engine._h, self.engine.W = self.lib, C.c_void_p(eh), self.W
            self.engine.N, self.engine.L, self.engine.data, self.engine.model = self.N, self.L, data, prm
            self.engine.a_pad = self.a_pad = self.lib.moire_engine_a_pad(self.engine._h)
            self.engine.T = self.lib.moire_engine_num_chains(self.engine._h)
            self.engine.close = lambda: None   # the driver owns it

    # ---- plumbing ----
    def _ck(self, rc, what):
        if rc != 0:
            raise MoireError(f"{what} failed ({rc}): "
                             f"{self.lib.moire_mcmc_last_error(self._h).decode()}")

    def close(self):
        if self._h is not None:
            if self.engine is not None:
                self.engine._h = None
            self.lib.moire_mcmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_nccl(self, device):
        """The native exchange (``moire_mcmc_init_nccl``): rank 0 makes the NCCL id, the process
        group broadcasts it, every rank joins the communicator.  Collective over the ranks of the
        initialised ``torch.distributed`` group; returns False (callback exchange stays) when the
        NCCL library cannot be bound."""
        import torch
        import torch.distributed as dist
        path = nccl_library_path()
        cpath = path.encode() if path else None
        if _NCCL_READY.get((self.W, self.world, self.rank)):     # the process keeps its communicator
            self._ck(self.lib.moire_mcmc_init_nccl(self._h, None, cpath), "init_nccl")
            return True
        idbuf = np.zeros(128, np.uint8)
        ok = 1
        if self.rank == 0:
            ok = int(self.lib.moire_mcmc_nccl_unique_id(_ptr(idbuf), cpath) == 0)
        t = torch.cat([torch.from_numpy(idbuf.copy()), torch.tensor([ok], dtype=torch.uint8)]).to(
            torch.device("cuda", device))
        dist.broadcast(t, 0)
        t = t.cpu().numpy()
        if not t[128]:
            return False
        idbuf = np.ascontiguousarray(t[:128])
        self._ck(self.lib.moire_mcmc_init_nccl(self._h, _ptr(idbuf), cpath), "init_nccl")
        _NCCL_READY[(self.W, self.world, self.rank)] = True   # per library: each keeps its own communicator
        return True

    def set_allgather(self, fn):
        """``fn(send: np.ndarray[count]) -> np.ndarray[world * count]``."""
        def tramp(_ctx, send, count, recv):
            try:
                out = fn(np.ctypeslib.as_array(send, shape=(count,)))
                np.ctypeslib.as_array(recv, shape=(count * self.world,))[:] = out
                return 0
            except Exception as exc:  # never unwind through C
                self._cb_error = exc
                return 1
        self._cb = _lib.ALLGATHER_FN(tramp)
        self._ck(self.lib.moire_mcmc_set_allgather(self._h, self._cb, None), "set_allgather")

    # ---- MCMC::MCMC ----
    def init(self):
        failed = C.c_int32()
        self._ck(self.lib.moire_mcmc_init(self._h, C.byref(failed)), "init")
        return not bool(failed.value)

    def init_counts(self):
        n = self.engine.T if self.engine is not None else 0
        ill, pri = np.zeros(max(n, 1), np.int32), np.zeros((max(n, 1), 5), np.int32)
        self._ck(self.lib.moire_mcmc_get_init_counts(self._h, _ptr(ill), _ptr(pri)), "init_counts")
        return ill[:n], pri[:n]

    def init_log(self):
        n = C.c_int32()
        eh = self.engine._h
        self.lib.moire_engine_read_init_log(eh, None, 0, C.byref(n))
        rec = np.zeros((max(n.value, 1), 3), np.int32)
        self.lib.moire_engine_read_init_log(eh, _ptr(rec), n.value, C.byref(n))
        return rec[:n.value]

    # ---- MCMC::burnin / sample / finalize ----
    def burnin(self, step):
        self._ck(self.lib.moire_mcmc_burnin(self._h, step), "burnin")

    def sample(self, step):
        self._ck(self.lib.moire_mcmc_sample(self._h, step), "sample")

    def finalize(self):
        self._ck(self.lib.moire_mcmc_finalize(self._h), "finalize")

    def run(self, progress=None):
        reached, minutes = C.c_int32(), C.c_double()
        cb = _lib.PROGRESS_FN(lambda _c, ph, st, post, hot: int(bool(progress(ph, st, post, hot)))) \
            if progress else C.cast(None, _lib.PROGRESS_FN)
        self._ck(self.lib.moire_mcmc_run(self._h, cb, None, C.byref(reached), C.byref(minutes)),
                 "run")
        return bool(reached.value), float(minutes.value)

    def mode(self):
        """'device ladder, step graph' | 'device ladder, plain launches' | 'host ladder'."""
        a, b = C.c_int32(), C.c_int32()
        self._ck(self.lib.moire_mcmc_get_mode(self._h, C.byref(a), C.byref(b)), "get_mode")
        if not a.value:
            return "host ladder"
        return "device ladder, " + ("step graph" if b.value else "plain launches")

    # ---- results ----
    def sizes(self):
        s = _lib.MoireMcmcSizes()
        self._ck(self.lib.moire_mcmc_get_sizes(self._h, C.byref(s)), "get_sizes")
        return {n: getattr(s, n) for n, _ in s._fields_}

    def traces(self):
        s = self.sizes()
        b, n = s["burnin_steps"], s["sample_steps"]
        out = dict(llik_burnin=np.zeros(b), prior_burnin=np.zeros(b), posterior_burnin=np.zeros(b),
                   llik_sample=np.zeros(n), prior_sample=np.zeros(n), posterior_sample=np.zeros(n))
        self._ck(self.lib.moire_mcmc_get_traces(
            self._h, _ptr(out["llik_burnin"]), _ptr(out["prior_burnin"]),
            _ptr(out["posterior_burnin"]), _ptr(out["llik_sample"]), _ptr(out["prior_sample"]),
            _ptr(out["posterior_sample"])), "get_traces")
        return out

    def ladder(self):
        s, T = self.sizes(), self.T
        out = dict(swap_store=np.zeros(s["num_swap_store"], np.int32),
                   swap_acceptances=np.zeros(max(T - 1, 0), np.int64),
                   swap_barriers=np.zeros(max(T - 1, 0)), temp_gradient=np.zeros(T),
                   swap_indices=np.zeros(T, np.int32), chain_temps=np.zeros(T))
        self._ck(self.lib.moire_mcmc_get_ladder(
            self._h, _ptr(out["swap_store"]), _ptr(out["swap_acceptances"]),
            _ptr(out["swap_barriers"]), _ptr(out["temp_gradient"]), _ptr(out["swap_indices"]),
            _ptr(out["chain_temps"])), "get_ladder")
        return out

    def draws(self):
        s = self.sizes()
        D, N, L, A = s["num_draws"], self.N, self.L, self.a_pad
        out = dict(draw_step=np.zeros(D, np.int32), allele_freqs=np.zeros((D, L, A)),
                   coi=np.zeros((D, N), np.int32), eps_neg=np.zeros((D, N)),
                   eps_pos=np.zeros((D, N)), relatedness=np.zeros((D, N)),
                   data_llik=np.zeros((D, N)), lam_coi=np.zeros(D),
                   latent=np.zeros((D, N, L) if self.W == 1 else (D, N, L, self.W), np.uint64)
                   if self.record_latent else None)
        self._ck(self.lib.moire_mcmc_get_draws(
            self._h, _ptr(out["draw_step"]), _ptr(out["allele_freqs"]), _ptr(out["coi"]),
            _ptr(out["eps_neg"]), _ptr(out["eps_pos"]), _ptr(out["relatedness"]),
            _ptr(out["data_llik"]), _ptr(out["lam_coi"]), _ptr(out["latent"])), "get_draws")
        return out

    # ---- test hooks ----
    def swap_pass(self, chain_llik, step, is_burnin, uniforms=None):
        ll = np.ascontiguousarray(chain_llik, dtype=np.float64)
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64)
            self._ck(self.lib.moire_mcmc_set_swap_uniforms(self._h, _ptr(u), len(u)), "set_u")
        self._ck(self.lib.moire_mcmc_swap_pass(self._h, _ptr(ll), step, int(is_burnin)), "swap_pass")

    def set_ladder(self, temp_gradient=None, swap_barriers=None, swap_indices=None,
                   chain_temps=None, num_swaps=-1, even_swap=-1):
        f = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
        keep = [f(temp_gradient, np.float64), f(swap_barriers, np.float64),
                f(swap_indices, np.int32), f(chain_temps, np.float64)]
        self._ck(self.lib.moire_mcmc_set_ladder(self._h, *[_ptr(k) for k in keep], int(num_swaps),
                                                int(even_swap)), "set_ladder")

    def adapt_temp(self):
        self._ck(self.lib.moire_mcmc_adapt_temp(self._h), "adapt_temp")

    def hot(self):
        out = np.zeros(self.T, np.int32)
        self._ck(self.lib.moire_mcmc_get_hot(self._h, _ptr(out)), "get_hot")
        return out

    def set_local_scalars(self, llik, prior):
        a = np.ascontiguousarray(llik, dtype=np.float64)
        b = np.ascontiguousarray(prior, dtype=np.float64)
        self._ck(self.lib.moire_mcmc_set_local_scalars(self._h, _ptr(a), _ptr(b)), "set_scalars")


def monotone_spline_solve(x, y, target):
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    return float(lib.moire_monotone_spline_solve(_ptr(x), _ptr(y), len(x), float(target)))


_NCCL_READY = {}


def nccl_library_path():
    """The libnccl.so.2 that torch itself uses (its bundled wheel), or None to let the loader find
    the one already in the process."""
    import importlib.util
    import os as _os
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations or []) if spec else []:
            cand = _os.path.join(root, "lib", "libnccl.so.2")
            if _os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def torch_allgather(world_size, device=None):
    """The ladder's one exchange as a ``torch.distributed`` all-gather (NCCL when ``device`` is a
    CUDA device, gloo on CPU).  Buffers are allocated once (pinned on the host for NCCL): a step
    costs one small H2D copy, the collective, one D2H copy and one stream synchronisation."""
    import torch
    import torch.distributed as dist
    state = {}

    def fn(send):
        n = send.shape[0]
        if device is None:
            out = torch.empty(world_size * n, dtype=torch.float64)
            dist.all_gather_into_tensor(out, torch.from_numpy(send))
            return out.numpy()
        if state.get("n") != n:
            state.update(n=n, send_h=torch.empty(n, dtype=torch.float64).pin_memory(),
                         recv_h=torch.empty(world_size * n, dtype=torch.float64).pin_memory(),
                         send_d=torch.empty(n, dtype=torch.float64, device=device),
                         recv_d=torch.empty(world_size * n, dtype=torch.float64, device=device))
        state["send_h"].numpy()[:] = send
        state["send_d"].copy_(state["send_h"], non_blocking=True)
        dist.all_gather_into_tensor(state["recv_d"], state["send_d"])
        state["recv_h"].copy_(state["recv_d"], non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return state["recv_h"].numpy()
    return fn


def _assemble_chain(parts, N, L, num_alleles, args, record_latent):
    """The 23-field list of src/main.cpp:138-189 in the nesting R/summary.R indexes into."""
    tr, lad, dr, diag = parts["traces"], parts["ladder"], parts["draws"], parts["diagnostics"]
    order = np.argsort(dr["draw_step"], kind="stable")
    for k, v in dr.items():
        if v is not None:
            dr[k] = v[order]
    D = len(dr["draw_step"])
    res = dict(tr)
    res["data_llik"] = list(np.ascontiguousarray(dr["data_llik"].T))
    res["coi"] = list(np.ascontiguousarray(dr["coi"].T))
    res["lam_coi"] = dr["lam_coi"]
    af = np.ascontiguousarray(dr["allele_freqs"].transpose(1, 0, 2))     # [L][D][A]
    res["allele_freqs"] = [list(af[j, :, :num_alleles[j]]) for j in range(L)]
    res["eps_neg"] = list(np.ascontiguousarray(dr["eps_neg"].T))
    res["eps_pos"] = list(np.ascontiguousarray(dr["eps_pos"].T))
    res["relatedness"] = list(np.ascontiguousarray(dr["relatedness"].T))
    if record_latent and dr["latent"] is not None:
        lat = dr["latent"].reshape(D, N, L, -1)  # [..][W] mask words, word w = alleles 64 w ..
        res["latent_genotypes"] = [[[[64 * w + b for w in range(lat.shape[3]) for b in range(64)
                                      if int(lat[d, i, j, w]) >> b & 1]
                                     for d in range(D)] for j in range(L)] for i in range(N)]
    else:
        # nothing recorded: every (sample, locus) entry is an empty list, as in the reference's
        # result; the rows are one shared object (read-only by convention, like an R value)
        res["latent_genotypes"] = [[[] for _ in range(L)]] * N
    res["observed_coi"] = parts["observed_coi"]
    res["swap_store"] = lad["swap_store"]
    res["swap_acceptances"] = lad["swap_acceptances"]
    res["swap_barriers"] = lad["swap_barriers"]
    res["temp_gradient"] = lad["temp_gradient"]
    na = np.asarray(num_alleles)
    if (na == na[0]).all():          # one allele count: slice the matrix once per chain
        def per_locus(mat):
            return list(mat[:, :int(na[0])])
    else:
        def per_locus(mat):
            return [mat[j, :na[j]] for j in range(L)]
    res["acceptance_rates"] = [dict(allele_freq_accept=per_locus(g["p_accept"]),
                                    coi_accept=g["m_accept"], eps_neg_accept=g["eps_neg_accept"],
                                    eps_pos_accept=g["eps_pos_accept"], r_accept=g["r_accept"],
                                    m_r_accept=g["m_r_accept"], full_sample_accept=g["sample_accept"],
                                    mean_coi_accept=g["mean_coi_accept"]) for g in diag]
    res["sampling_variances"] = [dict(allele_freq_var=per_locus(g["p_prop_var"]),
                                      eps_neg_var=g["eps_neg_var"], eps_pos_var=g["eps_pos_var"],
                                      r_var=g["r_var"], m_r_var=g["m_r_var"],
                                      mean_coi_var=g["mean_coi_var"]) for g in diag]
    res["max_runtime_reached"] = parts["max_runtime_reached"]
    res["total_runtime"] = parts["total_runtime"]
    lam = np.asarray(res["lam_coi"])
    res["mean_coi"] = lam / (1.0 - np.exp(-lam))      # R/mcmc.R:154
    return res


def run_mcmc(data, is_missing=False, allow_relatedness=True, thin=1, burnin=1e4,
             samples_per_chain=1e3, verbose=True, use_message=False, eps_pos_alpha=1,
             eps_pos_beta=1, eps_neg_alpha=1, eps_neg_beta=1, r_alpha=1, r_beta=1,
             mean_coi_shape=.1, mean_coi_scale=10, max_eps_pos=2, max_eps_neg=2, max_coi=40,
             record_latent_genotypes=False, num_chains=1, num_cores=1, pt_chains=1, pt_grad=1,
             pt_num_threads=1, adapt_temp=True, pre_adapt_steps=25, temp_adapt_steps=25,
             max_initialization_tries=10000, max_runtime=math.inf, *, seed=1, device=0,
             distributed=None, progress=None, pipeline="auto", ref32_emulation=False):
    """Same arguments and result as ``moire::run_mcmc`` (R/mcmc.R:76-182): returns
    ``dict(runtime, chains=[...], args)`` whose ``chains[i]`` holds the 23 fields of
    src/main.cpp:138-189 plus ``mean_coi``.

    ``data`` is what ``load_long_form_data`` / ``simulate_data`` return (``data["data"][locus]``
    a [N][A_j] 0/1 matrix, ``sample_ids``, ``loci``) or an already packed ``PackedData``.
    ``num_cores`` / ``pt_num_threads`` are accepted and ignored: every chain runs on the GPU.
    Extra keyword-only arguments the R function has no counterpart for: ``seed`` (the reference
    seeds from ``std::random_device``), ``device``, ``distributed`` -- ``True`` to shard the ladder
    over the ranks of an initialised ``torch.distributed`` group (every rank returns the full
    result) --, ``progress(phase, step, posterior, hot_chain)`` and ``pipeline`` ("auto", "flat",
    "fused": the kernel generation, include/moire_engine.h) and ``ref32_emulation`` (float-rounded
    P(any missing) clamped at 1 - FLT_EPSILON like the shipped float reference; off = fp64).

    Engine limits the reference does not have (``ValueError`` / ``MoireError`` when exceeded): at
    most 128 alleles per locus (data with a locus of 65..128 alleles runs on the two-word build of
    the engine, ``libmoire_b200_w128.so``, picked here from the data), ``max_coi`` <=
    MOIRE_MAX_COI = 127 (the reference's own bound), and chains-per-GPU x samples x loci < 2^32.
    """
    t_start = time.time()
    args = dict(is_missing=is_missing, allow_relatedness=allow_relatedness, thin=thin,
                burnin=burnin, samples_per_chain=samples_per_chain, verbose=verbose,
                use_message=use_message, eps_pos_alpha=eps_pos_alpha, eps_pos_beta=eps_pos_beta,
                eps_neg_alpha=eps_neg_alpha, eps_neg_beta=eps_neg_beta, r_alpha=r_alpha,
                r_beta=r_beta, mean_coi_shape=mean_coi_shape, mean_coi_scale=mean_coi_scale,
                max_eps_pos=max_eps_pos, max_eps_neg=max_eps_neg, max_coi=max_coi,
                record_latent_genotypes=record_latent_genotypes, num_chains=num_chains,
                num_cores=num_cores, pt_chains=pt_chains, pt_grad=pt_grad,
                pt_num_threads=pt_num_threads, adapt_temp=adapt_temp,
                pre_adapt_steps=pre_adapt_steps, temp_adapt_steps=temp_adapt_steps,
                max_initialization_tries=max_initialization_tries, max_runtime=max_runtime)
    sample_ids = loci = None
    if isinstance(data, PackedData):
        pk = data
    else:
        sample_ids, loci = data.get("sample_ids"), data.get("loci")
        pk = pack_data(data["data"], is_missing)
    if max_coi < 1:
        raise ValueError("Max COI must be greater than 1")   # R/mcmc.R:135-137
    ladder = _ladder(pt_chains, pt_grad)
    rank, world, gather, native = 0, 1, None, False
    if distributed:
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else None
        gather = torch_allgather(world, dev)
        native = dev is not None and not os.environ.get("MOIRE_B200_CALLBACK_EXCHANGE")
    model = dict(allow_relatedness=bool(allow_relatedness), max_coi=int(max_coi),
                 mean_coi_shape=float(mean_coi_shape), mean_coi_scale=float(mean_coi_scale),
                 max_eps_pos=float(max_eps_pos), eps_pos_alpha=float(eps_pos_alpha),
                 eps_pos_beta=float(eps_pos_beta), max_eps_neg=float(max_eps_neg),
                 eps_neg_alpha=float(eps_neg_alpha), eps_neg_beta=float(eps_neg_beta),
                 r_alpha=float(r_alpha), r_beta=float(r_beta))
    chains = []
    prof = os.environ.get("MOIRE_B200_PROFILE")    # phase times of the call, to stderr
    marks = [("pack", time.time())]
    for chain_number in range(1, int(num_chains) + 1):
        mc = MCMC(pk, ladder, burnin=int(burnin), samples=int(round(samples_per_chain)),
                  thin=int(thin), adapt_temp=adapt_temp, pre_adapt_steps=int(pre_adapt_steps),
                  temp_adapt_steps=int(temp_adapt_steps),
                  max_initialization_tries=int(max_initialization_tries),
                  record_latent_genotypes=record_latent_genotypes, max_runtime=max_runtime,
                  seed=int(seed) + 7919 * (chain_number - 1), device=device, rank=rank,
                  world_size=world, allgather=gather, pipeline=pipeline,
                  ref32_emulation=ref32_emulation, **model)
        try:
            if native:
                mc.init_nccl(device)     # falls back to the callback if NCCL cannot be bound
            marks.append(("create", time.time()))
            ok = mc.init()
            marks.append(("init", time.time()))
            if not ok:
                rec = [tuple(int(x) for x in r) for r in mc.init_log()]
                ill, pri = mc.init_counts()          # this rank's chains; merged over ranks below
                info = dict(records=rec, per_chain=ill.tolist(), prior=pri.sum(axis=0).tolist())
                if world > 1:
                    import torch.distributed as dist
                    allinfo = [None] * world
                    dist.all_gather_object(allinfo, info)
                else:
                    allinfo = [info]
                records = [r for i in allinfo for r in i["records"]]
                per_chain = [c for i in allinfo for c in i["per_chain"]]
                prior = dict(zip(PRIOR_TERMS, np.sum([i["prior"] for i in allinfo], axis=0).tolist()))
                diag = classify_initialization(records, prior, int(max_initialization_tries),
                                               len(ladder), per_chain)
                raise InitializationFailure(
                    format_initialization_failure_message(diag, sample_ids, loci), diag)
            reached, minutes = mc.run(progress)
            marks.append(("run", time.time()))
            parts = dict(traces=mc.traces(), ladder=mc.ladder(), draws=mc.draws(),
                         diagnostics=mc.engine.diagnostics_all(),
                         observed_coi=pk.observed_coi.copy(), max_runtime_reached=reached,
                         total_runtime=minutes)
            if world > 1:
                import torch.distributed as dist
                allparts = [None] * world
                dist.all_gather_object(allparts, dict(draws=parts["draws"],
                                                      diagnostics=parts["diagnostics"]))
                merged = {}
                for k in parts["draws"]:
                    vs = [a["draws"][k] for a in allparts if a["draws"][k] is not None]
                    merged[k] = np.concatenate(vs) if vs else None
                parts["draws"] = merged
                parts["diagnostics"] = [g for a in allparts for g in a["diagnostics"]]
            chains.append(_assemble_chain(parts, pk.num_samples, pk.num_loci, pk.num_alleles, args,
                                          record_latent_genotypes))
            marks.append(("assemble", time.time()))
        finally:
            mc.close()
    if prof:
        import sys
        prev = t_start
        for name, t in marks + [("close", time.time())]:
            print(f"run_mcmc phase {name}: {1e3 * (t - prev):.1f} ms", file=sys.stderr)
            prev = t
    return dict(runtime=time.time() - t_start, chains=chains, args=args)
