"""Python face of the C-ABI engine: the calls the reference's MCMC driver makes on its ``Chain``
objects (src/mcmc.cpp), for a block of tempered chains resident on one GPU.

Every method is a thin ctypes call into libmoire_b200.so; no likelihood arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .data import PackedData

UPDATE_KINDS = dict(eps_neg=0, eps_pos=1, p=2, m=3, r=4, eff_coi=5, samples=6, mean_coi=7)
KIND_NAMES = ["eps_neg", "eps_pos", "p", "m", "r", "eff_coi", "samples", "mean_coi", "reduce"]

# model defaults of moire::run_mcmc (R/mcmc.R:76-106)
DEFAULT_MODEL = dict(allow_relatedness=True, burnin=10000, max_coi=40, mean_coi_shape=0.1,
                     mean_coi_scale=10.0, max_eps_pos=2.0, eps_pos_alpha=1.0, eps_pos_beta=1.0,
                     max_eps_neg=2.0, eps_neg_alpha=1.0, eps_neg_beta=1.0, r_alpha=1.0,
                     r_beta=1.0)


# moire_params.pipeline (include/moire_engine.h): kernel generation, "auto" = by problem size
PIPELINES = dict(auto=0, flat=1, fused=2)
NUMERICS_REF32 = 0x100   # MOIRE_NUMERICS_REF32: float-rounded PAM clamped at 1 - FLT_EPSILON
P_CACHE_ON, P_CACHE_OFF = 0x200, 0x400   # MOIRE_P_CACHE_*: update_p's PAM cache (default: by shape)


def pipeline_flags(pipeline, ref32_emulation=False):
    """moire_params.pipeline from "auto" | "flat" | "fused", optionally followed by "+cache" /
    "-cache" (update_p's PAM cache forced on / off; flat pipelines only)."""
    name, flags = pipeline, 0
    if pipeline.endswith("+cache"):
        name, flags = pipeline[:-6], P_CACHE_ON
    elif pipeline.endswith("-cache"):
        name, flags = pipeline[:-6], P_CACHE_OFF
    return PIPELINES[name] | flags | (NUMERICS_REF32 if ref32_emulation else 0)


class MoireError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """``num_chains`` tempered chains [chain_offset, chain_offset + num_chains) on one device."""

    def __init__(self, data: PackedData, temps, seed=1, device=0, chain_offset=0, pipeline="auto",
                 ref32_emulation=False, **model):
        self._h = None
        # one-word masks (libmoire_b200.so) unless a locus has more than 64 alleles
        self.W = data.mask_words
        self.lib = _lib.load(self.W)
        prm = dict(DEFAULT_MODEL)
        unknown = set(model) - set(prm)
        if unknown:
            raise TypeError(f"unknown model parameters: {sorted(unknown)}")
        prm.update(model)
        self.model = prm
        self.data = data
        self.N, self.L = data.num_samples, data.num_loci
        temps = np.ascontiguousarray(np.atleast_1d(np.asarray(temps, dtype=np.float64)))
        self.T = len(temps)
        self._keep = dict(
            na=np.ascontiguousarray(data.num_alleles, dtype=np.int32),
            obs=np.ascontiguousarray(data.obs_mask, dtype=np.uint64),
            mis=np.ascontiguousarray(data.is_missing, dtype=np.uint8),
            coi=np.ascontiguousarray(data.observed_coi, dtype=np.int32))
        d = _lib.MoireData(self.N, self.L, _ptr(self._keep["na"]), _ptr(self._keep["obs"]),
                           _ptr(self._keep["mis"]), _ptr(self._keep["coi"]))
        p = _lib.MoireParams(int(bool(prm["allow_relatedness"])), int(prm["burnin"]),
                             int(prm["max_coi"]), pipeline_flags(pipeline, ref32_emulation),
                             prm["mean_coi_shape"], prm["mean_coi_scale"],
                             prm["max_eps_pos"], prm["eps_pos_alpha"], prm["eps_pos_beta"],
                             prm["max_eps_neg"], prm["eps_neg_alpha"], prm["eps_neg_beta"],
                             prm["r_alpha"], prm["r_beta"])
        h = C.c_void_p()
        rc = self.lib.moire_engine_create(C.byref(d), C.byref(p), self.T, chain_offset,
                                          _ptr(temps), device, seed, C.byref(h))
        if rc != 0:
            raise MoireError(f"moire_engine_create failed ({rc}): "
                             f"{self.lib.moire_last_error().decode()}")
        self._h = h
        self.a_pad = self.lib.moire_engine_a_pad(h)

    def close(self):
        if self._h is not None:
            self.lib.moire_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise MoireError(f"{what} failed ({rc}): "
                             f"{self.lib.moire_engine_last_error(self._h).decode()}")

    # ---- MCMC::MCMC ----
    def init_chains(self, max_tries=10000):
        T = self.T
        out = dict(ill_count=np.zeros(T, np.int32), still_ill=np.zeros(T, np.int32),
                   first_bad_sample=np.zeros(T, np.int32), first_bad_locus=np.zeros(T, np.int32),
                   prior_nonfinite=np.zeros((T, 5), np.int32))
        self._ck(self.lib.moire_engine_init_chains(
            self._h, max_tries, _ptr(out["ill_count"]), _ptr(out["still_ill"]),
            _ptr(out["first_bad_sample"]), _ptr(out["first_bad_locus"]),
            _ptr(out["prior_nonfinite"])), "init_chains")
        return out

    def diagnose(self, chain=0):
        """(first_bad_sample, first_bad_locus, prior_nonfinite[5]) of the chain's current state:
        Chain::first_nonfinite_genotyping_term / collect_nonfinite_prior_terms, 1-based, 0 = none."""
        bs, bl = C.c_int32(), C.c_int32()
        pri = np.zeros(5, np.int32)
        self._ck(self.lib.moire_engine_diagnose(self._h, chain, C.byref(bs), C.byref(bl), _ptr(pri)),
                 "diagnose")
        return int(bs.value), int(bl.value), pri

    # ---- MCMC::burnin / sample loop body ----
    def step(self, iteration):
        self._ck(self.lib.moire_engine_step(self._h, iteration), "step")

    def update(self, kind, iteration):
        k = UPDATE_KINDS[kind] if isinstance(kind, str) else int(kind)
        self._ck(self.lib.moire_engine_update(self._h, k, iteration), "update")

    def eval_all(self):
        self._ck(self.lib.moire_engine_eval_all(self._h), "eval_all")

    def reduce(self):
        self._ck(self.lib.moire_engine_reduce(self._h), "reduce")

    def synchronize(self):
        self._ck(self.lib.moire_engine_synchronize(self._h), "synchronize")

    # ---- Chain::get_llik / get_prior / temps ----
    def scalars(self):
        llik, prior = np.zeros(self.T), np.zeros(self.T)
        self._ck(self.lib.moire_engine_get_scalars(self._h, _ptr(llik), _ptr(prior)), "get_scalars")
        return llik, prior

    def device_scalars(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self.lib.moire_engine_device_scalars(self._h, C.byref(a), C.byref(b)),
                 "device_scalars")
        return a.value, b.value

    def set_temps(self, temps):
        t = np.ascontiguousarray(np.asarray(temps, dtype=np.float64))
        assert len(t) == self.T
        self._ck(self.lib.moire_engine_set_temps(self._h, _ptr(t)), "set_temps")

    def get_temps(self):
        t = np.zeros(self.T)
        self._ck(self.lib.moire_engine_get_temps(self._h, _ptr(t)), "get_temps")
        return t

    # ---- state ----
    def _mask_shape(self):
        """[N][L] allele sets of one chain: uint64 [N][L], or [N][L][2] with two-word masks."""
        return (self.N, self.L) if self.W == 1 else (self.N, self.L, self.W)

    def dump_state(self, chain=0):
        N, L, A = self.N, self.L, self.a_pad
        st = dict(m=np.zeros(N, np.int32), r=np.zeros(N), eps_pos=np.zeros(N), eps_neg=np.zeros(N),
                  p=np.zeros((L, A)), latent=np.zeros(self._mask_shape(), np.uint64), lg_adj=np.zeros((N, L)),
                  mean_coi=np.zeros(1), cell_t=np.zeros((N, L)), cell_o=np.zeros((N, L)),
                  prior_eps_neg=np.zeros(N), prior_eps_pos=np.zeros(N), prior_coi=np.zeros(N),
                  prior_r=np.zeros(N), mean_coi_hyper_prior=np.zeros(1))
        cs = _lib.MoireChainState(*[_ptr(st[n]) for n, _ in _lib.MoireChainState._fields_])
        self._ck(self.lib.moire_engine_dump_state(self._h, chain, C.byref(cs)), "dump_state")
        st["mean_coi"] = float(st["mean_coi"][0])
        st["mean_coi_hyper_prior"] = float(st["mean_coi_hyper_prior"][0])
        return st

    def load_state(self, chain, st):
        N, L, A = self.N, self.L, self.a_pad
        keep = {}

        def arr(name, dtype, shape):
            if name not in st or st[name] is None:
                return None
            a = np.ascontiguousarray(np.asarray(st[name], dtype=dtype))
            if name == "p" and a.shape != shape:
                b = np.zeros(shape, dtype)
                b[:, :a.shape[1]] = a
                a = b
            assert a.shape == shape, (name, a.shape, shape)
            keep[name] = a
            return _ptr(a)

        vals = dict(m=arr("m", np.int32, (N,)), r=arr("r", np.float64, (N,)),
                    eps_pos=arr("eps_pos", np.float64, (N,)),
                    eps_neg=arr("eps_neg", np.float64, (N,)), p=arr("p", np.float64, (L, A)),
                    latent=arr("latent", np.uint64, self._mask_shape()),
                    lg_adj=arr("lg_adj", np.float64, (N, L)))
        if "mean_coi" in st and st["mean_coi"] is not None:
            keep["mc"] = np.array([st["mean_coi"]], dtype=np.float64)
            vals["mean_coi"] = _ptr(keep["mc"])
        cs = _lib.MoireChainState(*[vals.get(n) for n, _ in _lib.MoireChainState._fields_])
        self._ck(self.lib.moire_engine_load_state(self._h, chain, C.byref(cs)), "load_state")

    def diagnostics(self, chain=0):
        N, L, A = self.N, self.L, self.a_pad
        d = dict(eps_neg_accept=np.zeros(N, np.int32), eps_pos_accept=np.zeros(N, np.int32),
                 m_accept=np.zeros(N, np.int32), r_accept=np.zeros(N, np.int32),
                 m_r_accept=np.zeros(N, np.int32), sample_accept=np.zeros(N, np.int32),
                 eps_neg_var=np.zeros(N), eps_pos_var=np.zeros(N), r_var=np.zeros(N),
                 m_r_var=np.zeros(N), p_accept=np.zeros((L, A), np.int32),
                 p_attempt=np.zeros((L, A), np.int32), p_prop_var=np.zeros((L, A)),
                 mean_coi_accept=np.zeros(1), mean_coi_var=np.zeros(1))
        cd = _lib.MoireChainDiagnostics(*[_ptr(d[n]) for n, _ in
                                          _lib.MoireChainDiagnostics._fields_])
        self._ck(self.lib.moire_engine_read_diagnostics(self._h, chain, C.byref(cd)),
                 "read_diagnostics")
        d["mean_coi_accept"] = float(d["mean_coi_accept"][0])
        d["mean_coi_var"] = float(d["mean_coi_var"][0])
        return d

    def diagnostics_all(self):
        """``diagnostics`` of every chain in one device read: a list of per-chain dicts (views)."""
        T, N, L, A = self.T, self.N, self.L, self.a_pad
        i32, f64 = np.int32, np.float64
        d = dict(eps_neg_accept=np.zeros((T, N), i32), eps_pos_accept=np.zeros((T, N), i32),
                 m_accept=np.zeros((T, N), i32), r_accept=np.zeros((T, N), i32),
                 m_r_accept=np.zeros((T, N), i32), sample_accept=np.zeros((T, N), i32),
                 eps_neg_var=np.zeros((T, N), f64), eps_pos_var=np.zeros((T, N), f64),
                 r_var=np.zeros((T, N), f64), m_r_var=np.zeros((T, N), f64),
                 p_accept=np.zeros((T, L, A), i32), p_attempt=np.zeros((T, L, A), i32),
                 p_prop_var=np.zeros((T, L, A), f64), mean_coi_accept=np.zeros(T, f64),
                 mean_coi_var=np.zeros(T, f64))
        cd = _lib.MoireChainDiagnostics(*[_ptr(d[n]) for n, _ in
                                          _lib.MoireChainDiagnostics._fields_])
        self._ck(self.lib.moire_engine_read_diagnostics(self._h, -1, C.byref(cd)),
                 "read_diagnostics")
        return [{k: (float(v[c]) if v.ndim == 1 else v[c]) for k, v in d.items()} for c in range(T)]

    def sample_llik(self, chain=0):
        out = np.zeros(self.N)
        self._ck(self.lib.moire_engine_read_sample_llik(self._h, chain, _ptr(out)),
                 "read_sample_llik")
        return out

    def record_draw(self, chain=0, latent=False):
        N, L, A = self.N, self.L, self.a_pad
        d = dict(p=np.zeros((L, A)), m=np.zeros(N, np.int32), eps_neg=np.zeros(N),
                 eps_pos=np.zeros(N), r=np.zeros(N), sample_llik=np.zeros(N),
                 mean_coi=np.zeros(1), latent=np.zeros(self._mask_shape(), np.uint64) if latent else None)
        self._ck(self.lib.moire_engine_record_draw(
            self._h, chain, _ptr(d["p"]), _ptr(d["m"]), _ptr(d["eps_neg"]), _ptr(d["eps_pos"]),
            _ptr(d["r"]), _ptr(d["sample_llik"]), _ptr(d["mean_coi"]), _ptr(d["latent"])),
            "record_draw")
        d["mean_coi"] = float(d["mean_coi"][0])
        return d

    def subset_order(self, which, k):
        """Subset enumeration order for k latent alleles as decoded on the device from table
        ``which`` (0 tail table, 1 mask table, 2 unrolled): (masks[2^k - 1], negative[2^k - 1])."""
        out = np.zeros((1 << k) - 1, np.uint16)
        self._ck(self.lib.moire_engine_subset_order(self._h, which, k, _ptr(out)), "subset_order")
        return out & 0x7FFF, (out & 0x8000) != 0

    def counters(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.moire_engine_counters(self._h, C.byref(a), C.byref(b)), "counters")
        return int(a.value), int(b.value)

    def kind_evals(self):
        out = np.zeros(8, np.uint64)
        self._ck(self.lib.moire_engine_kind_evals(self._h, _ptr(out)), "kind_evals")
        return {k: int(out[i]) for i, k in enumerate(KIND_NAMES[:8])}

    def timer_start(self):
        self._ck(self.lib.moire_engine_timer_start(self._h), "timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.lib.moire_engine_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return float(ms.value)

    def set_timing(self, on=True):
        self._ck(self.lib.moire_engine_set_timing(self._h, int(on)), "set_timing")

    def timing(self):
        ms = np.zeros(9)
        n = np.zeros(9, np.uint64)
        self._ck(self.lib.moire_engine_read_timing(self._h, _ptr(ms), _ptr(n)), "read_timing")
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KIND_NAMES)}


def _evaluator_timing(self):
    """(ms, launches) of the transmission evaluator kernel alone since timing was enabled."""
    ms, n = C.c_double(), C.c_uint64()
    self._ck(self.lib.moire_engine_read_evaluator_timing(self._h, C.byref(ms), C.byref(n)),
             "read_evaluator_timing")
    return float(ms.value), int(n.value)


Engine.evaluator_timing = _evaluator_timing


def trim_memory():
    """Hand the device buffers cached from closed engines back to the CUDA driver."""
    _lib.load().moire_engine_trim_memory()


def philox4x32(counter, key):
    lib = _lib.load()
    c = np.ascontiguousarray(counter, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib.moire_philox4x32(_ptr(c), _ptr(k), _ptr(out))
    return out
