"""Host-side parallel-tempering logic (moire_b200/csrc/mcmc_driver.cpp) against known answers
minted from the reference's own MCMC::swap_chains / MCMC::adapt_temp (tests/golden/
make_golden_driver.py), plus the sharded ladder on two gloo processes.  All of it runs in the
driver's ladder-only mode: no GPU, no likelihood evaluation."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny_pack():
    from moire_b200.data import pack_data
    from moire_b200.simulate import simulate_data
    d = simulate_data(mean_coi=2, num_samples=6, epsilon_pos=.02, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(4)] * 3, seed=3)
    return pack_data(d["data"], d["is_missing"])


def ladder_only(T, temps=None, **kw):
    from moire_b200.mcmc import MCMC
    temps = np.linspace(1, 0, T) if temps is None else temps
    base = dict(burnin=100, samples=10, pre_adapt_steps=5, temp_adapt_steps=3, device=-1)
    base.update(kw)
    return MCMC(tiny_pack(), temps, **base)


def test_partition_is_contiguous_and_balanced():
    import ctypes as C
    from moire_b200 import _lib
    lib = _lib.load()
    for T in (1, 7, 40, 64):
        for W in range(1, min(T, 8) + 1):
            nxt, sizes = 0, []
            for r in range(W):
                f, c = C.c_int32(), C.c_int32()
                lib.moire_mcmc_partition(T, r, W, C.byref(f), C.byref(c))
                assert f.value == nxt and c.value >= 1
                nxt += c.value
                sizes.append(c.value)
            assert nxt == T and max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("prec", [64, 32])
def test_swap_pass_matches_reference(prec):
    """MCMC::swap_chains, src/mcmc.cpp:190-240: index map, temperatures, hot flag, barrier
    accumulators, acceptance counters, num_swaps -- decisions bit-exact (fed the reference's own
    uniforms), accumulators to rounding."""
    g = load_golden(f"driver_ref{prec}.npz")
    tol = 1e-12 if prec == 64 else 2e-5
    mc, curT, flips = None, None, 0
    for case in range(int(g["num_swap_cases"])):
        pre = f"swap{case}."
        T, even, step, burnin, ns = (int(x) for x in g[pre + "in"])
        if T != curT:
            if mc is not None:
                mc.close()
            mc, curT = ladder_only(T), T
        llik = g[pre + "llik"]
        mc.set_ladder(swap_barriers=g[pre + "sb"], swap_indices=g[pre + "si"],
                      chain_temps=g[pre + "temps"], num_swaps=ns, even_swap=even)
        lad0 = mc.ladder()
        # acceptances are not settable; track the increment instead
        logu = g[pre + "logu"]
        mc.swap_pass(llik, step, burnin, uniforms=np.exp(logu))
        lad = mc.ladder()
        assert np.array_equal(lad["swap_indices"], g[pre + "out_swap_indices"]), case
        assert np.allclose(lad["chain_temps"], g[pre + "out_chain_temps"], rtol=0, atol=1e-7 if prec == 32 else 0), case
        assert np.allclose(lad["swap_barriers"], g[pre + "out_swap_barriers"], rtol=tol, atol=tol), case
        d_acc = lad["swap_acceptances"] - lad0["swap_acceptances"]
        assert np.array_equal(d_acc, g[pre + "out_swap_acceptances"] - g[pre + "sa"]), case
        assert lad["swap_store"][-1] == g[pre + "out_swap_indices"][0]
        flips += int((lad["swap_indices"] != g[pre + "si"]).any())
    mc.close()
    assert flips > 10   # the set does exercise accepted swaps


def test_hot_flag_quirk():
    """No chain is hot until the first accepted swap at ladder position 0 or the first
    adapt_temp (src/mcmc.cpp:221-225, 294)."""
    mc = ladder_only(3)
    assert mc.hot().tolist() == [0, 0, 0]
    # pass 1 starts at pair 1 (even_swap = true -> i = 1): a certain swap there sets no hot flag
    mc.swap_pass([-100.0, -50.0, -10.0], 0, True, uniforms=[0.5])
    assert mc.ladder()["swap_indices"].tolist() == [0, 2, 1] and mc.hot().tolist() == [0, 0, 0]
    # pass 2 starts at pair 0: chain 2 (now at position 1) has the better llik -> swap -> hot
    mc.swap_pass([-100.0, -50.0, -10.0], 1, True, uniforms=[0.5])
    assert mc.ladder()["swap_indices"].tolist() == [2, 0, 1] and mc.hot().tolist() == [0, 0, 1]
    assert mc.ladder()["chain_temps"].tolist() == [0.5, 0.0, 1.0]
    mc.close()


@pytest.mark.parametrize("prec", [64, 32])
def test_adapt_temp_matches_reference(prec):
    """MCMC::adapt_temp (src/mcmc.cpp:244-304) incl. tk::spline's monotone cubic + solve_one."""
    g = load_golden(f"driver_ref{prec}.npz")
    tol = 1e-9 if prec == 64 else 5e-4
    changed = bailed = 0
    for case in range(int(g["num_adapt_cases"])):
        pre = f"adapt{case}."
        grad, barriers = g[pre + "gradient"], g[pre + "barriers"]
        T = len(grad)
        mc = ladder_only(T, temps=grad)
        mc.set_ladder(temp_gradient=grad, swap_barriers=barriers, chain_temps=grad,
                      num_swaps=int(g[pre + "num_swaps"]))
        mc.adapt_temp()
        lad = mc.ladder()
        want = g[pre + "out"]
        assert np.allclose(lad["temp_gradient"], want, rtol=0, atol=tol), (case, lad["temp_gradient"], want)
        if np.array_equal(want, grad):
            bailed += 1
        else:
            changed += 1
            assert np.all(np.diff(lad["temp_gradient"]) <= 0)       # still ordered (ties as the reference)
            assert np.allclose(lad["chain_temps"], lad["temp_gradient"])  # identity index map
            assert mc.hot()[0] == 1 and not lad["swap_barriers"].any()
        mc.close()
    assert changed + bailed == 24 and changed >= 12


def test_spline_solve_is_monotone_inverse():
    from moire_b200.mcmc import monotone_spline_solve
    rng = np.random.default_rng(3)
    for n in (3, 5, 17):
        x = np.sort(rng.uniform(0, 1, n))
        x[0], x[-1] = 0.0, 1.0
        y = np.r_[0.0, np.cumsum(rng.uniform(0, 1, n - 1) ** 3)]
        ts = np.linspace(0, y[-1], 50)[:-1]
        xs = np.array([monotone_spline_solve(x, y, t) for t in ts])
        assert np.all(np.isfinite(xs)) and np.all(np.diff(xs) >= -1e-12)
        for k in range(n - 1):   # interpolates the knots
            assert abs(monotone_spline_solve(x, y, y[k]) - x[k]) < 1e-9


def test_single_rank_loop_bookkeeping():
    """Traces, swap_store and the adaptation schedule of MCMC::burnin / MCMC::sample
    (src/mcmc.cpp:157-188, 306-355) with injected scalars."""
    T = 4
    mc = ladder_only(T, burnin=40, samples=7, pre_adapt_steps=5, temp_adapt_steps=3)
    assert mc.init()
    rng = np.random.default_rng(0)
    for step in range(40):
        mc.set_local_scalars(-rng.gamma(30, 3, T), -rng.gamma(3, 1, T))
        mc.burnin(step)
    for step in range(7):
        mc.set_local_scalars(-rng.gamma(30, 3, T), -rng.gamma(3, 1, T))
        mc.sample(step)
    mc.finalize()
    s, tr, lad = mc.sizes(), mc.traces(), mc.ladder()
    assert s["burnin_steps"] == 40 and s["sample_steps"] == 7 and s["num_swap_store"] == 47
    assert np.all(np.isfinite(tr["posterior_burnin"])) and len(tr["llik_sample"]) == 7
    assert lad["temp_gradient"][0] == 1.0 and lad["temp_gradient"][-1] == 0.0
    assert np.all(np.diff(lad["temp_gradient"]) < 0)
    assert not np.allclose(lad["temp_gradient"], np.linspace(1, 0, T))   # the ladder adapted
    assert sorted(lad["swap_indices"].tolist()) == list(range(T))
    assert sorted(lad["chain_temps"].tolist(), reverse=True) == lad["temp_gradient"].tolist()
    mc.close()


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from test_mcmc_driver import ladder_only
from moire_b200.mcmc import torch_allgather
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, T = dist.get_rank(), 7
mc = ladder_only(T, burnin=30, samples=6, rank=rank, world_size=2, allgather=torch_allgather(2))
assert mc.init()
first, count = mc.sizes()["first_chain"], mc.sizes()["local_chains"]
rng = np.random.default_rng(5)
for step in range(36):
    llik, prior = -rng.gamma(30, 3, T), -rng.gamma(3, 1, T)    # same on both ranks; each reports its block
    mc.set_local_scalars(llik[first:first + count], prior[first:first + count])
    (mc.burnin if step < 30 else mc.sample)(step if step < 30 else step - 30)
lad, tr = mc.ladder(), mc.traces()
np.savez({out!r} + str(rank), first=first, count=count, draws=mc.sizes()["num_draws"], hot=mc.hot(), **lad, **tr)
mc.close()
dist.destroy_process_group()
"""


def test_sharded_ladder_two_gloo_ranks_agree_with_one(tmp_path):
    """The N>1 path on CPU: two processes own 3 + 4 chains of a 7-chain ladder, exchange scalars
    through a gloo all-gather, and must hold the same ladder as each other and as a single rank
    fed the same scalars; exactly one rank stores each draw."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port, out=out))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    r0, r1 = (dict(np.load(out + f"{r}.npz")) for r in range(2))
    # single rank, same scalars
    T = 7
    mc = ladder_only(T, burnin=30, samples=6)
    assert mc.init()
    rng = np.random.default_rng(5)
    for step in range(36):
        mc.set_local_scalars(-rng.gamma(30, 3, T), -rng.gamma(3, 1, T))
        (mc.burnin if step < 30 else mc.sample)(step if step < 30 else step - 30)
    lad, tr, hot, draws1 = mc.ladder(), mc.traces(), mc.hot(), mc.sizes()["num_draws"]
    mc.close()
    assert (int(r0["first"]), int(r0["count"]), int(r1["first"]), int(r1["count"])) == (0, 3, 3, 4)
    for k, v in list(lad.items()) + list(tr.items()):
        assert np.array_equal(r0[k], v) and np.array_equal(r1[k], v), k
    assert np.array_equal(r0["hot"], hot) and np.array_equal(r1["hot"], hot)
    assert int(r0["draws"]) + int(r1["draws"]) == draws1


def test_run_loop_progress_and_interrupt():
    """The two loops of run_mcmc (src/main.cpp:52-93): the progress callback sees every step of
    both phases; a non-zero return stops the run like Rcpp::checkUserInterrupt unwinding."""
    from moire_b200.engine import MoireError
    mc = ladder_only(3, burnin=6, samples=4)
    assert mc.init()
    seen = []
    reached, minutes = mc.run(lambda phase, step, post, hot: seen.append((phase, step)) or False)
    assert seen == [(0, s) for s in range(1, 7)] + [(1, s) for s in range(1, 5)]
    assert reached is False and minutes == 0.0
    s = mc.sizes()
    assert (s["burnin_steps"], s["sample_steps"], s["num_swap_store"]) == (6, 4, 10)
    mc.close()
    mc = ladder_only(3, burnin=6, samples=4)
    assert mc.init()
    with pytest.raises(MoireError, match="interrupted"):
        mc.run(lambda phase, step, post, hot: phase == 0 and step == 3)
    assert mc.sizes()["burnin_steps"] == 3
    mc.close()
