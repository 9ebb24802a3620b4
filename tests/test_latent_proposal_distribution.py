"""The latent-genotype proposal (a8, src/sampler.cpp:230-340): the engine draws it from its own
Philox streams with selection sampling instead of the reference's shuffle, so the two cannot be
compared draw by draw.  They must be the same DISTRIBUTION: here the engine-semantics sampler (the
CPU stepper, which the GPU kernels match bit for bit in tests/test_gpu_parity.py) and the
reference's own `Sampler::sample_latent_genotype` (oracle/_ref, unmodified) are each drawn tens of
thousands of times on the same inputs and compared -- joint counts of (false negatives, false
positives), per-allele inclusion frequencies, and the log proposal density attached to each
outcome."""
import numpy as np
import pytest

from oracle.pyoracle import Oracle, Ref, ref_available

pytestmark = pytest.mark.skipif(not ref_available(64), reason="oracle/_ref not built")

CASES = [  # observed alleles (0/1 per allele), coi, eps_pos, eps_neg
    ([1, 0, 1, 0, 0, 0, 0, 0, 0, 0], 4, 0.3, 0.6),
    ([0, 0, 0, 0, 0, 0], 2, 0.2, 0.9),                  # nothing observed: one forced false negative
    ([1, 1, 1, 1, 0, 0, 0, 0], 2, 0.8, 0.4),            # more observed than the COI allows
    ([1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0], 7, 0.5, 1.2),
]


def popcount(a):
    return np.array([bin(int(x)).count("1") for x in a])


@pytest.mark.parametrize("obs,coi,eps_pos,eps_neg", CASES)
def test_engine_and_reference_proposals_have_the_same_distribution(obs, coi, eps_pos, eps_neg):
    n = 30000
    A = len(obs)
    obs_mask = sum(1 << a for a, o in enumerate(obs) if o)
    orc, ref = Oracle(), Ref(64)
    masks, lp = orc.stepper_sample_latent(12345, n, obs_mask, A, coi, eps_pos, eps_neg)
    ref_masks, ref_lp = np.zeros(n, np.uint64), np.zeros(n)
    for i in range(n):
        idx, l = ref.sample_latent_genotype(1000 + i, obs, coi, eps_pos, eps_neg)
        ref_masks[i] = sum(1 << int(a) for a in idx)
        ref_lp[i] = l
    # outcomes without a valid genotype are reported the same way (empty set, -inf)
    ok, ref_ok = np.isfinite(lp), np.isfinite(ref_lp)
    assert abs(ok.mean() - ref_ok.mean()) < 4 * np.sqrt(0.25 / n) + 1e-12
    masks, lp, ref_masks, ref_lp = masks[ok], lp[ok], ref_masks[ref_ok], ref_lp[ref_ok]
    # (fn, fp): latent alleles not observed / observed alleles not latent
    def counts(ms):
        m = ms.astype(np.uint64)
        fn = popcount(m & ~np.uint64(obs_mask))
        fp = popcount(np.uint64(obs_mask) & ~m)
        return fn, fp
    (fn, fp), (rfn, rfp) = counts(masks), counts(ref_masks)
    assert (popcount(masks) <= coi).all() and (popcount(masks) >= 1).all()
    top = max(fn.max(), rfn.max(), fp.max(), rfp.max()) + 1
    h = np.zeros((top, top))
    rh = np.zeros((top, top))
    np.add.at(h, (fn, fp), 1)
    np.add.at(rh, (rfn, rfp), 1)
    p, rp = h / len(fn), rh / len(rfn)
    se = np.sqrt((p + rp) / 2 * (1 - (p + rp) / 2) * (1 / len(fn) + 1 / len(rfn)))
    assert (np.abs(p - rp) <= 4.5 * se + 2e-4).all(), (p, rp)
    # per-allele inclusion
    inc = np.array([((masks >> np.uint64(a)) & np.uint64(1)).mean() for a in range(A)])
    rinc = np.array([((ref_masks >> np.uint64(a)) & np.uint64(1)).mean() for a in range(A)])
    se = np.sqrt((inc + rinc) / 2 * (1 - (inc + rinc) / 2) * (2 / n)) + 1e-4
    assert (np.abs(inc - rinc) <= 4.5 * se).all(), (inc, rinc)
    # the density attached to an outcome is a function of (fn, fp) only, and the same function
    for a in range(top):
        for b in range(top):
            mine, theirs = lp[(fn == a) & (fp == b)], ref_lp[(rfn == a) & (rfp == b)]
            if len(mine) and len(theirs):
                assert np.ptp(mine) < 1e-9 and np.ptp(theirs) < 1e-9
                assert abs(mine[0] - theirs[0]) < 1e-9 * max(1.0, abs(theirs[0])), (a, b)


def test_coi_step_has_the_reference_distribution():
    """Sampler::sample_coi_delta(2) (src/sampler.cpp:152-156: a fair sign times a
    std::geometric_distribution with success probability 1/3) against the engine's draw by
    inversion: the same law on the integers."""
    n = 200000
    mine = Oracle().stepper_coi_delta(99, n)
    ref = Ref(64)
    if not hasattr(ref.lib, "ref_sample_coi_delta_many"):
        pytest.skip("oracle/_ref predates ref_sample_coi_delta_many; rebuild with make -C oracle")
    theirs = ref.sample_coi_delta(7, 2.0, n)
    lo, hi = min(mine.min(), theirs.min()), max(mine.max(), theirs.max())
    for v in range(lo, hi + 1):
        p, q = (mine == v).mean(), (theirs == v).mean()
        pooled = (p + q) / 2
        se = np.sqrt(pooled * (1 - pooled) * 2 / n)
        assert abs(p - q) <= 4.5 * se + 2e-5, (v, p, q)
    # and the law itself: P(|X| = k) = (1/3)(2/3)^k, the sign fair (zero keeps its mass)
    for k in range(0, 8):
        want = (1 / 3) * (2 / 3) ** k
        got = (np.abs(mine) == k).mean()
        assert abs(got - want) < 4.5 * np.sqrt(want * (1 - want) / n), (k, got, want)
    assert abs((mine > 0).mean() - (mine < 0).mean()) < 4.5 * np.sqrt(0.5 / n)
