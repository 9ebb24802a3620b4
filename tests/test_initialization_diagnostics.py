"""The reference's initialization-failure tests (tests/testthat/test-initialization-diagnostics.R)
restated against the Python mirror of its formatter, plus the classifier
(InitializationDiagnostics::classify, src/initialization_diagnostics.h:62-133) on the engine's
per-attempt failure records.  (Like the reference's own test file says, a live failing start is
not reliable to provoke: in double precision the starting likelihood is finite for every input we
could construct.)"""
import re

from moire_b200.mcmc import (InitializationFailure, classify_initialization,
                             format_initialization_failure_message)


def test_initialization_failure_carries_diagnostics_and_message():
    diagnostics = dict(max_initialization_tries=10, chains_attempted=1, total_ill_conditioned=10,
                       unknown_genotyping_count=0, locus_counts=dict(locus=[2], n=[9]),
                       sample_counts=dict(sample=[1], n=[9]),
                       prior_nonfinite_counts=dict(term=[], n=[]), examples=dict(sample=[1], locus=[2]),
                       ill_conditioned_per_chain=[10], classification="consistent_failure_cause",
                       concentration="locus", dominant_locus=2, dominant_sample=None,
                       dominant_share=0.9, secondary_sample=1)
    msg = format_initialization_failure_message(diagnostics, sample_ids=["S1", "S2"], loci=["L1", "L2"])
    err = InitializationFailure(msg, diagnostics)
    assert err.diagnostics["classification"] == "consistent_failure_cause"
    assert re.search(r"Consistent failure cause: locus 2 \(L2\)", str(err))
    assert "Guidance: inspect this locus" in str(err)
    assert re.search(r"Secondary concentration on sample 1 \(S1\)", str(err))
    assert "in 90% of Ill-conditioned starts" in str(err)


def test_hard_starting_set_guidance_is_emitted():
    diagnostics = dict(max_initialization_tries=5, chains_attempted=1, total_ill_conditioned=5,
                       unknown_genotyping_count=0, locus_counts=dict(locus=[1, 2], n=[2, 3]),
                       sample_counts=dict(sample=[1, 2], n=[2, 3]),
                       prior_nonfinite_counts=dict(term=[], n=[]), examples=dict(sample=[], locus=[]),
                       ill_conditioned_per_chain=[5], classification="hard_starting_set",
                       concentration="", dominant_locus=2, dominant_sample=None, dominant_share=0.6,
                       secondary_sample=None)
    msg = format_initialization_failure_message(diagnostics)
    assert "hard starting set" in msg and "increase max_initialization_tries" in msg


def test_classifier_follows_the_reference_rules():
    # 9 of 10 failures at locus 2 (and at sample 1): locus preferred, sample reported as secondary
    rec = [(0, 1, 2)] * 9 + [(0, 3, 4)]
    d = classify_initialization(rec, {}, 10, 1, [10])
    assert (d["classification"], d["concentration"], d["dominant_locus"]) == ("consistent_failure_cause", "locus", 2)
    assert d["secondary_sample"] == 1 and abs(d["dominant_share"] - 0.9) < 1e-12
    assert d["locus_counts"] == dict(locus=[2, 4], n=[9, 1]) and d["examples"]["sample"][:2] == [1, 1]
    # one sample everywhere, loci spread: sample concentration
    rec = [(0, 7, j) for j in range(1, 11)]
    d = classify_initialization(rec, {}, 10, 2, [5, 5])
    assert (d["concentration"], d["dominant_sample"], d["dominant_locus"]) == ("sample", 7, None)
    # nothing dominates; ties go to the locus (best_locus_count >= best_sample_count)
    rec = [(0, 1, 1), (0, 2, 2), (0, 3, 3), (0, 0, 0)]
    d = classify_initialization(rec, dict(coi=2, eps_neg=0), 4, 1, [4])
    assert d["classification"] == "hard_starting_set" and d["unknown_genotyping_count"] == 1
    assert d["dominant_locus"] == 1 and d["dominant_sample"] is None and abs(d["dominant_share"] - 0.25) < 1e-12
    assert d["prior_nonfinite_counts"] == dict(term=["coi"], n=[2])
    assert classify_initialization([], {}, 3, 1, [0])["classification"] == "hard_starting_set"
