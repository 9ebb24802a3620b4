"""The reference's own loader tests (tests/testthat/test-utils.R:1-33) restated, plus the
``.rda`` reader against the packaged data files where /root/reference exists and against the
fixture minted from them (tests/golden/make_packaged_fixtures.py) everywhere."""
import os

import numpy as np
import pytest

from conftest import load_golden
from moire_b200.data import pack_data
from moire_b200.loaders import load_delimited_data, load_long_form_data, read_rda

REF_DATA = "/root/reference/data"


def test_loading_long_form_data_works():
    rng = np.random.default_rng(0)
    rows = list(zip(["S1", "S1", "S1", "S2", "S2", "S3", "S3"], ["A", "A", "B", "A", "B", "A", "A"],
                    ["1", "2", "1", "1", "2", "3", "1"]))
    rows = [rows[i] for i in rng.permutation(len(rows))]          # slice_sample(prop = 1)
    data = dict(sample_id=[r[0] for r in rows], locus=[r[1] for r in rows], allele=[r[2] for r in rows])
    res = load_long_form_data(data)
    assert res["sample_ids"] == ["S1", "S2", "S3"]
    assert res["loci"] == ["A", "B"]
    assert res["data"][0][0].tolist() == [1, 1, 0]
    assert res["data"][0][1].tolist() == [1, 0, 0]
    assert res["data"][0][2].tolist() == [1, 0, 1]
    assert int(res["is_missing"].sum()) == 1
    assert res["is_missing"].shape == (2, 3) and res["is_missing"][1, 2]   # S3 has nothing at B


def test_loading_delimited_data_works():
    data = dict(sample_id=["S1", "S2", "S3"], A=["1;2", "1", "1;3"], B=[None, "1", "2;3"])
    res = load_delimited_data(data)
    assert res["sample_ids"] == ["S1", "S2", "S3"]
    assert res["loci"] == ["A", "B"]
    assert res["data"][0][0].tolist() == [1, 1, 0]
    assert res["data"][0][1].tolist() == [1, 0, 0]
    assert res["data"][0][2].tolist() == [1, 0, 1]
    assert int(res["is_missing"].sum()) == 1


def test_pandas_frames_and_uninformative_loci(capsys):
    pd = pytest.importorskip("pandas")
    df = pd.DataFrame(dict(sample_id=["b", "a", "a", "b"], locus=["x", "x", "mono", "mono"],
                           allele=[2, 10, 7, 7]))
    res = load_long_form_data(df)
    assert "Uninformative loci" in capsys.readouterr().out
    assert res["uninformative_loci"] == ["mono"] and res["loci"] == ["x"]
    assert res["sample_ids"] == ["a", "b"]
    assert res["data"][0].tolist() == [[0, 1], [1, 0]]            # alleles sorted numerically: 2, 10
    pk = pack_data(res["data"], res["is_missing"])
    assert pk.obs_mask.tolist() == [[2], [1]]


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="packaged .rda files only in the build container")
def test_read_rda_decodes_the_packaged_files():
    sd = read_rda(os.path.join(REF_DATA, "simulated_data.rda"))["simulated_data"]
    assert list(sd)[:4] == ["data", "sample_ids", "loci", "is_missing"]
    assert len(sd["data"]) == 30 and len(sd["data"]["L1"]) == 100
    assert sd["is_missing"].shape == (30, 100) and int(sd["is_missing"].sum()) == 2
    nd = read_rda(os.path.join(REF_DATA, "namibia_data.rda"))["namibia_data"]
    assert len(nd["sample_id"]) == 97214
    mr = read_rda(os.path.join(REF_DATA, "mcmc_results.rda"))["mcmc_results"]
    assert abs(np.mean(mr["chains"][0]["llik_sample"]) + 10483.1) < 0.1
    g = load_golden("packaged_data.npz")                            # the fixture is what we decode now
    pk = pack_data(sd["data"], sd["is_missing"] != 0)
    assert np.array_equal(pk.obs_mask, g["sim.obs_mask"]) and np.array_equal(pk.is_missing, g["sim.is_missing"])
    assert read_rda(os.path.join(REF_DATA, "regional_allele_frequencies.rda"))


def test_packaged_fixture_matches_the_survey_facts():
    g = load_golden("packaged_data.npz")
    assert g["sim.obs_mask"].shape == (100, 30) and int(g["sim.is_missing"].sum()) == 2
    assert g["sim.num_alleles"].tolist() == [5] * 15 + [10] * 15
    assert g["nam.obs_mask"].shape == (2585, 26) and int(g["nam.is_missing"].sum()) == 7651
    assert int(g["nam.num_alleles"].sum()) == 451 and g["nam.num_alleles"].min() == 7 and g["nam.num_alleles"].max() == 35
    pop = np.vectorize(lambda x: bin(int(x)).count("1"))(g["nam.obs_mask"])
    assert pop.max() == 15 and abs((pop[pop > 0] == 1).mean() - 0.626) < 0.002
    assert len(g["res.pt_chains"]) == 80 and abs(g["res.llik_sample"].mean() + 10483.1) < 0.1
