"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints one
JSON line with the agreed keys (it runs the unmodified reference's C++ sampler from oracle/_ref on
the host cores), ranks other than 0 stay silent, and our arm fails loudly -- no JSON line, non-zero
exit -- when there is no device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmoire_ref32.so"))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "c1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mcmc_steps_per_sec_all_temps"
    assert d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["config"]["workload"].startswith("c1")


def test_reference_arm_is_silent_on_other_ranks():
    r = run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0"], env=dict(RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = run_bench(["--steps", "1", "--warmup", "1", "--workload", "c1", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
