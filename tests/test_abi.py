"""The C-ABI library loads and exports every symbol include/moire_engine.h declares; without a
GPU the product fails loudly instead of falling back to anything."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(moire_[a-z0-9_]+)\s*\(", src))
    return names


@pytest.mark.parametrize("words", [1, 2])
def test_library_exports_every_declared_symbol(words):
    """Both builds of the one ABI: libmoire_b200.so (one-word allele masks, up to 64 alleles per
    locus) and libmoire_b200_w128.so (two words, up to 128)."""
    import ctypes
    from moire_b200 import _lib
    lib = _lib.load(words)
    assert lib.moire_engine_mask_words() == words and lib.moire_engine_max_alleles() == 64 * words
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert decl >= set(_lib.ENGINE_SYMBOLS) | set(_lib.MCMC_SYMBOLS), "binding names a symbol the header does not declare"
    assert isinstance(lib, ctypes.CDLL)


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    from moire_b200.engine import philox4x32
    kat = [([0] * 4, [0] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        assert [int(x) for x in philox4x32(c, k)] == want


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from moire_b200.data import pack_data
    from moire_b200.engine import Engine, MoireError
    from moire_b200.simulate import simulate_data
    d = simulate_data(mean_coi=3, num_samples=5, epsilon_pos=.01, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(4)] * 3, seed=1)
    with pytest.raises(MoireError, match="no CPU fallback"):
        Engine(pack_data(d["data"], d["is_missing"]), [1.0])


def test_wide_loci_pack_into_two_mask_words():
    """pack_data: a locus with more than 64 alleles switches the whole data set to [N][L][2] masks
    (word 0 = alleles 0..63) and with that to the two-word library; more than 128 is an error."""
    from moire_b200 import _lib
    from moire_b200.data import pack_data
    rng = np.random.default_rng(3)
    mats = [(rng.random((7, a)) < .1).astype(int) for a in (100, 5, 128)]
    pk = pack_data(mats)
    assert pk.obs_mask.shape == (7, 3, 2) and pk.mask_words == 2 and _lib.words_for(pk.max_alleles) == 2
    for j, m in enumerate(mats):
        for i in range(7):
            x = int(pk.obs_mask[i, j, 0]) | (int(pk.obs_mask[i, j, 1]) << 64)
            assert [a for a in range(m.shape[1]) if x >> a & 1] == list(np.flatnonzero(m[i]))
    assert np.array_equal(pk.observed_coi, np.max([m.sum(1) for m in mats], axis=0))
    narrow = pack_data(mats[1:2] * 2)
    assert narrow.obs_mask.shape == (7, 2) and narrow.mask_words == 1
    assert np.array_equal(narrow.widened(2).obs_mask[:, :, 0], narrow.obs_mask)
    with pytest.raises(ValueError, match="two-word"):
        pack_data([np.ones((2, 129), int)])
    with pytest.raises(ValueError):
        _lib.load(3)


def test_wide_build_needs_a_device_too():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from moire_b200.data import pack_data
    from moire_b200.engine import Engine, MoireError
    with pytest.raises(MoireError, match="no CPU fallback"):
        Engine(pack_data([(np.arange(200).reshape(2, 100) % 7 == 0).astype(int)]), [1.0])


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "moire_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, fn)
