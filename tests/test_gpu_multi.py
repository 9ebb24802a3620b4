"""The temperature ladder sharded over two GPUs (one process per GPU, NCCL all-gather of the
per-chain scalars) reproduces the single-GPU run bit for bit: chains never move, streams are keyed
by global chain id, and every rank makes the same swap decisions."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, pickle
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from test_gpu_mcmc import small_data
from moire_b200.mcmc import run_mcmc
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = small_data()
res = run_mcmc(d, d["is_missing"], burnin=60, samples_per_chain=30, pt_chains=6, verbose=False,
               seed=4, device=local, distributed=True)
if dist.get_rank() == 0:
    ch = res["chains"][0]
    keep = {{k: ch[k] for k in ("llik_burnin", "llik_sample", "prior_sample", "posterior_sample",
                                "lam_coi", "swap_store", "temp_gradient", "swap_acceptances",
                                "swap_barriers", "coi", "allele_freqs", "relatedness")}}
    keep["n_accept_chains"] = len(ch["acceptance_rates"])
    pickle.dump(keep, open({out!r}, "wb"))
dist.destroy_process_group()
"""


def test_two_gpu_ladder_reproduces_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import pickle
    from test_gpu_mcmc import small_data
    from moire_b200.mcmc import run_mcmc
    out = str(tmp_path / "r0.pkl")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    two = pickle.load(open(out, "rb"))
    d = small_data()
    one = run_mcmc(d, d["is_missing"], burnin=60, samples_per_chain=30, pt_chains=6,
                   verbose=False, seed=4)["chains"][0]
    assert two["n_accept_chains"] == 6
    for k in ("llik_burnin", "llik_sample", "prior_sample", "posterior_sample", "lam_coi",
              "swap_store", "temp_gradient", "swap_acceptances", "swap_barriers"):
        assert np.array_equal(np.asarray(two[k]), np.asarray(one[k])), k
    for k in ("coi", "relatedness"):
        assert all(np.array_equal(a, b) for a, b in zip(two[k], one[k])), k
    assert all(np.array_equal(a, b) for la, lb in zip(two["allele_freqs"], one["allele_freqs"])
               for a, b in zip(la, lb))
