#!/usr/bin/env python
"""Mints tests/golden/driver_ref{32,64}.npz: known answers for the host-side parallel-tempering
logic -- MCMC::swap_chains (src/mcmc.cpp:190-240) and MCMC::adapt_temp (src/mcmc.cpp:244-304,
tk::spline) -- from the UNMODIFIED reference compiled under oracle/_ref (see make_golden.py).

    make -C oracle ref && python tests/golden/make_golden_driver.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from moire_b200.data import pack_data  # noqa: E402
from moire_b200.simulate import simulate_data  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def mint(prec):
    ref = Ref(prec)
    d = simulate_data(mean_coi=2, num_samples=6, epsilon_pos=.02, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(4)] * 3, seed=3)
    pk = pack_data(d["data"], d["is_missing"])
    ref.set_data(pk.num_alleles, pk.obs_mask, pk.is_missing)
    rng = np.random.default_rng(2024)
    g = {}
    case = 0
    for T in (2, 3, 5, 8, 13):
        ladder = np.linspace(1.0, 0.0, T) ** rng.choice([1.0, 2.0, 0.5])
        ref.set_params(pt_chains=ladder, pre_adapt_steps=5, temp_adapt_steps=3, burnin=100)
        h = ref.mcmc_new(seed=17)
        # ---- swap passes: a sequence, state carried from pass to pass like the real driver
        si = np.arange(T)
        temps = ladder.copy()
        sb, sa, ns = np.zeros(T - 1), np.zeros(T - 1, np.int64), 0
        even = 1
        for rep in range(12):
            llik = -rng.gamma(50.0, 4.0, size=T) * (1.0 + 2.0 * temps)
            if rep == 7:
                llik[rng.integers(T)] = np.nan          # NaN ratio => no swap, barrier += 0
            step, burnin = int(rng.integers(0, 12)), int(rep % 3 != 2)
            seed = 1000 + 37 * case
            ref.seed_runif(seed)
            logu = np.array([ref.draw_log_runif() for _ in range(T)])
            ref.seed_runif(seed)
            out = ref.mcmc_swap_pass(h, llik, temps, si, even, step, burnin, sb, sa, ns)
            pre = f"swap{case}."
            g[pre + "in"] = np.array([T, even, step, burnin, ns], np.int64)
            g[pre + "llik"], g[pre + "temps"], g[pre + "si"] = llik, temps.copy(), si.copy()
            g[pre + "sb"], g[pre + "sa"], g[pre + "logu"] = sb.copy(), sa.copy(), logu
            for k, v in out.items():
                g[pre + "out_" + k] = np.asarray(v)
            si, temps = out["swap_indices"], out["chain_temps"]
            sb, sa, ns = out["swap_barriers"], out["swap_acceptances"], out["num_swaps"]
            even = 1 - even
            case += 1
        ref.mcmc_free(h)
    g["num_swap_cases"] = np.array(case)
    # ---- adapt_temp
    case = 0
    for T in (3, 4, 6, 10, 20, 40):
        for kind in range(4):
            ladder = np.linspace(1.0, 0.0, T) ** (1.0 + kind)
            ref.set_params(pt_chains=ladder)
            h = ref.mcmc_new(seed=5)
            num_swaps = int(rng.integers(4, 60))
            barriers = rng.uniform(0, 1, T - 1) * num_swaps / 2
            if kind == 1:
                barriers[rng.integers(T - 1)] = 0.0     # a flat segment
            if kind == 2:
                barriers *= rng.uniform(0, 1, T - 1) ** 4  # very uneven
            if kind == 3:
                barriers[:] = 0.0                       # nothing rejected yet -> NaN grid -> bail
            out = ref.mcmc_adapt_temp(h, ladder, barriers, num_swaps)
            pre = f"adapt{case}."
            g[pre + "gradient"], g[pre + "barriers"] = ladder, barriers
            g[pre + "num_swaps"], g[pre + "out"] = np.array(num_swaps), out
            ref.mcmc_free(h)
            case += 1
    g["num_adapt_cases"] = np.array(case)
    return g


if __name__ == "__main__":
    for prec in (32, 64):
        np.savez_compressed(os.path.join(OUT, f"driver_ref{prec}.npz"), **mint(prec))
        print(prec, os.path.getsize(os.path.join(OUT, f"driver_ref{prec}.npz")))
