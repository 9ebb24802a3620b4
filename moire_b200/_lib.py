"""Loader for the CUDA engine's C-ABI shared library (libmoire_b200.so).

The product path has no CPU fallback: if the library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C moire_b200/csrc``) importing fails loudly, and if
there is no CUDA device ``moire_engine_create`` returns MOIRE_ECUDA.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MOIRE_B200_LIB: an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("MOIRE_B200_LIB") or os.path.join(HERE, "libmoire_b200.so")
# the same sources built with two-word allele masks (loci with 65..128 alleles, include/moire_engine.h
# "Mask width"); the one-word library serves everything else
LIB_PATH_W128 = os.environ.get("MOIRE_B200_LIB_W128") or os.path.join(HERE, "libmoire_b200_w128.so")

vp, i32, u64, dbl = C.c_void_p, C.c_int32, C.c_uint64, C.c_double


class MoireData(C.Structure):
    _fields_ = [("num_samples", i32), ("num_loci", i32), ("num_alleles", vp), ("obs_mask", vp),
                ("is_missing", vp), ("observed_coi", vp)]


class MoireParams(C.Structure):
    _fields_ = [("allow_relatedness", i32), ("burnin", i32), ("max_coi", i32), ("pipeline", i32),
                ("mean_coi_shape", dbl), ("mean_coi_scale", dbl),
                ("max_eps_pos", dbl), ("eps_pos_alpha", dbl), ("eps_pos_beta", dbl),
                ("max_eps_neg", dbl), ("eps_neg_alpha", dbl), ("eps_neg_beta", dbl),
                ("r_alpha", dbl), ("r_beta", dbl)]


class MoireChainState(C.Structure):
    _fields_ = [(n, vp) for n in ("m", "r", "eps_pos", "eps_neg", "p", "latent", "lg_adj",
                                  "mean_coi", "cell_t", "cell_o", "prior_eps_neg",
                                  "prior_eps_pos", "prior_coi", "prior_r",
                                  "mean_coi_hyper_prior")]


class MoireChainDiagnostics(C.Structure):
    _fields_ = [(n, vp) for n in ("eps_neg_accept", "eps_pos_accept", "m_accept", "r_accept",
                                  "m_r_accept", "sample_accept", "eps_neg_var", "eps_pos_var",
                                  "r_var", "m_r_var", "p_accept", "p_attempt", "p_prop_var",
                                  "mean_coi_accept", "mean_coi_var")]


class MoireMcmcParams(C.Structure):
    _fields_ = [("thin", i32), ("burnin", i32), ("samples", i32), ("adapt_temp", i32),
                ("pre_adapt_steps", i32), ("temp_adapt_steps", i32),
                ("max_initialization_tries", i32), ("record_latent_genotypes", i32),
                ("max_runtime", dbl)]


class MoireMcmcSizes(C.Structure):
    _fields_ = [(n, i32) for n in ("num_chains", "first_chain", "local_chains", "burnin_steps",
                                   "sample_steps", "num_draws", "num_swap_store", "a_pad")]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, vp, C.POINTER(dbl), i32, C.POINTER(dbl))
PROGRESS_FN = C.CFUNCTYPE(C.c_int, vp, i32, i32, dbl, i32)

# every symbol include/moire_mcmc.h declares
MCMC_SYMBOLS = {
    "moire_mcmc_create": (C.c_int, [C.POINTER(MoireData), C.POINTER(MoireParams),
                                    C.POINTER(MoireMcmcParams), vp, i32, i32, i32, i32, u64,
                                    C.POINTER(vp)]),
    "moire_mcmc_destroy": (C.c_int, [vp]),
    "moire_mcmc_last_error": (C.c_char_p, [vp]),
    "moire_mcmc_partition": (None, [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]),
    "moire_mcmc_set_allgather": (C.c_int, [vp, ALLGATHER_FN, vp]),
    "moire_mcmc_nccl_unique_id": (C.c_int, [vp, C.c_char_p]),
    "moire_mcmc_init_nccl": (C.c_int, [vp, vp, C.c_char_p]),
    "moire_mcmc_engine": (vp, [vp]),
    "moire_mcmc_init": (C.c_int, [vp, C.POINTER(i32)]),
    "moire_mcmc_get_init_counts": (C.c_int, [vp, vp, vp]),
    "moire_mcmc_burnin": (C.c_int, [vp, i32]),
    "moire_mcmc_sample": (C.c_int, [vp, i32]),
    "moire_mcmc_finalize": (C.c_int, [vp]),
    "moire_mcmc_run": (C.c_int, [vp, PROGRESS_FN, vp, C.POINTER(i32), C.POINTER(dbl)]),
    "moire_mcmc_get_sizes": (C.c_int, [vp, C.POINTER(MoireMcmcSizes)]),
    "moire_mcmc_get_mode": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32)]),
    "moire_mcmc_get_traces": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "moire_mcmc_get_ladder": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "moire_mcmc_get_hot": (C.c_int, [vp, vp]),
    "moire_mcmc_get_draws": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "moire_mcmc_swap_pass": (C.c_int, [vp, vp, i32, i32]),
    "moire_mcmc_set_swap_uniforms": (C.c_int, [vp, vp, i32]),
    "moire_mcmc_set_ladder": (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, i32]),
    "moire_mcmc_adapt_temp": (C.c_int, [vp]),
    "moire_mcmc_set_local_scalars": (C.c_int, [vp, vp, vp]),
    "moire_monotone_spline_solve": (dbl, [vp, vp, i32, dbl]),
}

# every symbol include/moire_engine.h declares: (restype, argtypes)
ENGINE_SYMBOLS = {
    "moire_last_error": (C.c_char_p, []),
    "moire_engine_last_error": (C.c_char_p, [vp]),
    "moire_engine_create": (C.c_int, [C.POINTER(MoireData), C.POINTER(MoireParams), i32, i32, vp,
                                      i32, u64, C.POINTER(vp)]),
    "moire_engine_destroy": (C.c_int, [vp]),
    "moire_engine_a_pad": (i32, [vp]),
    "moire_engine_mask_words": (i32, []),
    "moire_engine_max_alleles": (i32, []),
    "moire_engine_num_chains": (i32, [vp]),
    "moire_engine_init_chains": (C.c_int, [vp, i32, vp, vp, vp, vp, vp]),
    "moire_engine_read_init_log": (C.c_int, [vp, vp, i32, vp]),
    "moire_engine_diagnose": (C.c_int, [vp, i32, vp, vp, vp]),
    "moire_engine_step": (C.c_int, [vp, i32]),
    "moire_engine_update": (C.c_int, [vp, i32, i32]),
    "moire_engine_eval_all": (C.c_int, [vp]),
    "moire_engine_reduce": (C.c_int, [vp]),
    "moire_engine_get_scalars": (C.c_int, [vp, vp, vp]),
    "moire_engine_device_scalars": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
    "moire_engine_set_temps": (C.c_int, [vp, vp]),
    "moire_engine_get_temps": (C.c_int, [vp, vp]),
    "moire_engine_dump_state": (C.c_int, [vp, i32, C.POINTER(MoireChainState)]),
    "moire_engine_load_state": (C.c_int, [vp, i32, C.POINTER(MoireChainState)]),
    "moire_engine_read_diagnostics": (C.c_int, [vp, i32, C.POINTER(MoireChainDiagnostics)]),
    "moire_engine_read_sample_llik": (C.c_int, [vp, i32, vp]),
    "moire_engine_record_draw": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "moire_engine_counters": (C.c_int, [vp, vp, vp]),
    "moire_engine_kind_evals": (C.c_int, [vp, vp]),
    "moire_engine_timer_start": (C.c_int, [vp]),
    "moire_engine_timer_stop": (C.c_int, [vp, C.POINTER(dbl)]),
    "moire_engine_set_timing": (C.c_int, [vp, i32]),
    "moire_engine_trim_memory": (C.c_int, []),
    "moire_engine_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "moire_engine_read_timing": (C.c_int, [vp, vp, vp]),
    "moire_engine_read_evaluator_timing": (C.c_int, [vp, C.POINTER(dbl), C.POINTER(u64)]),
    "moire_engine_synchronize": (C.c_int, [vp]),
    "moire_engine_subset_order": (C.c_int, [vp, i32, i32, vp]),
    "moire_engine_device_iteration": (C.c_int, [vp, C.POINTER(vp)]),
    "moire_engine_device_temps": (C.c_int, [vp, C.POINTER(vp)]),
    "moire_engine_captured_step_launches": (C.c_int, [vp, C.POINTER(C.c_int64)]),
    "moire_engine_add_launches": (C.c_int, [vp, C.c_int64]),
    "moire_engine_get_timing": (C.c_int, [vp, C.POINTER(i32)]),
    "moire_engine_enqueue_record_draw": (C.c_int, [vp, vp, vp]),
    "moire_simulate_data": (C.c_int, [i32, i32, vp, vp, i32, vp, vp, dbl, dbl, dbl, u64, i32, vp, vp, vp]),
    "moire_philox4x32": (None, [vp, vp, vp]),
}

_libs = {}


def load(mask_words: int = 1):
    """Returns the ctypes handle of libmoire_b200.so (``mask_words`` = 1: up to 64 alleles per
    locus) or libmoire_b200_w128.so (2: up to 128) with every prototype set."""
    if mask_words in _libs:
        return _libs[mask_words]
    if mask_words not in (1, 2):
        raise ValueError("mask_words must be 1 or 2 (at most 128 alleles per locus)")
    path = LIB_PATH if mask_words == 1 else LIB_PATH_W128
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA engine first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in list(ENGINE_SYMBOLS.items()) + list(MCMC_SYMBOLS.items()):
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.moire_engine_mask_words() != mask_words:
        raise ImportError(f"{path} was built with {lib.moire_engine_mask_words()}-word masks")
    _libs[mask_words] = lib
    return lib


def words_for(max_alleles: int) -> int:
    """Mask words (= which library) a data set with this widest locus needs."""
    return 1 if max_alleles <= 64 else 2
