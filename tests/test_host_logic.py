import numpy as np
import pytest

from moire_b200.data import pack_data
from moire_b200.simulate import simulate_data


def test_pack_data_masks_and_observed_coi():
    loc0 = np.array([[1, 0, 1], [0, 0, 0], [0, 1, 0]])
    loc1 = np.array([[1, 1, 1, 1], [0, 0, 0, 1], [0, 0, 0, 0]])
    pk = pack_data([loc0, loc1], np.array([[0, 1, 0], [0, 0, 1]], bool))
    assert pk.num_samples == 3 and pk.num_loci == 2
    assert pk.obs_mask.tolist() == [[0b101, 0b1111], [0, 0b1000], [0b010, 0]]
    assert pk.is_missing.tolist() == [[0, 0], [1, 0], [0, 1]]
    assert pk.observed_coi.tolist() == [4, 1, 1]
    assert pk.num_alleles.tolist() == [3, 4]


def test_pack_data_rejects_single_allele_locus_and_wide_locus():
    with pytest.raises(ValueError, match="less than 2 alleles"):
        pack_data([np.ones((3, 1))])
    assert pack_data([np.ones((3, 65))]).mask_words == 2      # past one mask word: the two-word build
    with pytest.raises(ValueError, match="128"):
        pack_data([np.ones((3, 129))])
    with pytest.raises(ValueError, match="is_missing"):
        pack_data([np.ones((3, 2))], np.zeros((3, 1), bool))


def test_simulate_data_shapes_and_model():
    d = simulate_data(mean_coi=4, num_samples=300, epsilon_pos=0.0, epsilon_neg=0.0,
                      locus_freq_alphas=[np.ones(6)] * 5, seed=2)
    assert len(d["data"]) == 5 and d["data"][0].shape == (300, 6)
    assert d["is_missing"].shape == (5, 300) and not d["is_missing"].any()
    assert d["sample_cois"].min() >= 1
    assert abs(d["sample_cois"].mean() - 4 / (1 - np.exp(-4))) < 0.4
    # without observation noise the observed alleles are exactly the present ones
    for j in range(5):
        assert np.array_equal(d["data"][j] != 0, d["true_genotypes"][j] > 0)
        assert (d["data"][j].sum(1) <= d["sample_cois"]).all()
    d2 = simulate_data(mean_coi=4, num_samples=50, epsilon_pos=.1, epsilon_neg=.1,
                       locus_freq_alphas=[np.ones(6)] * 2, missingness=1.0, seed=2)
    assert d2["is_missing"].all()
