"""Repack ``GenotypingData`` into the engine's SoA bit-mask layout.

The reference keeps ``observed_alleles[locus][sample][allele] in {0,1}`` as nested
``vector<int>`` and ``is_missing_[locus][sample]`` (src/genotyping_data.cpp:14-58).  The engine
wants, sample-major like the reference's ``genotyping_llik_new[sample * L + locus]``
(src/chain.cpp:1116):

* ``obs_mask``    uint64 [N][L]   bit a = allele a observed (A_j <= 64), or [N][L][2] (word 0 =
                  alleles 0..63) when a locus has 65..128 alleles: the two-word build of the engine
* ``is_missing``  uint8  [N][L]
* ``num_alleles`` int32  [L]
* ``observed_coi`` int32 [N]      max over loci of #observed alleles (genotyping_data.cpp:33-43)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MAX_ALLELES = 128   # two mask words; one word (64) is the common case and the fast library


@dataclass
class PackedData:
    num_samples: int
    num_loci: int
    num_alleles: np.ndarray
    obs_mask: np.ndarray
    is_missing: np.ndarray
    observed_coi: np.ndarray

    @property
    def max_alleles(self) -> int:
        return int(self.num_alleles.max())

    @property
    def mask_words(self) -> int:
        """64-bit words per allele set: 1, or 2 when some locus has more than 64 alleles."""
        return 1 if self.obs_mask.ndim == 2 else int(self.obs_mask.shape[2])

    def widened(self, words: int = 2) -> "PackedData":
        """The same data with ``words`` mask words per allele set (zero-padded): runs a data set
        whose loci all fit one word on the two-word build of the engine (tests do)."""
        obs = self.obs_mask.reshape(self.num_samples, self.num_loci, -1)
        out = np.zeros((self.num_samples, self.num_loci, words), dtype=np.uint64)
        out[:, :, :obs.shape[2]] = obs
        return PackedData(self.num_samples, self.num_loci, self.num_alleles,
                          out if words > 1 else out[:, :, 0].copy(), self.is_missing, self.observed_coi)


def pack_data(data, is_missing=False) -> PackedData:
    """``data``: list over loci of [N][A_j] 0/1 arrays (or list of lists of vectors, as R hands
    them over); ``is_missing``: False or a [L][N] boolean matrix (R/mcmc.R:115-125)."""
    if hasattr(data, "values"):          # a named list keyed by locus (read_rda, R's data$data)
        data = list(data.values())
    L = len(data)
    if L == 0:
        raise ValueError("no loci")
    mats = [np.asarray(np.stack([np.asarray(v) for v in loc]) if not isinstance(loc, np.ndarray)
                       else loc) for loc in data]
    N = mats[0].shape[0]
    num_alleles = np.array([m.shape[1] for m in mats], dtype=np.int32)
    if (num_alleles < 2).any():
        # R/mcmc.R:127-133
        raise ValueError("Loci with less than 2 alleles present, remove these loci")
    if (num_alleles > MAX_ALLELES).any():
        raise ValueError(f"at most {MAX_ALLELES} alleles per locus fit the engine's two-word masks")
    W = 1 if int(num_alleles.max()) <= 64 else 2
    obs = np.zeros((N, L, W), dtype=np.uint64)
    coi = np.zeros(N, dtype=np.int32)
    for j, m in enumerate(mats):
        if m.shape[0] != N:
            raise ValueError("ragged sample dimension")
        bits = (m != 0).astype(np.uint64)
        for w in range(W):
            part = bits[:, 64 * w:64 * (w + 1)]
            weights = (np.uint64(1) << np.arange(part.shape[1], dtype=np.uint64))
            obs[:, j, w] = (part * weights[None, :]).sum(axis=1, dtype=np.uint64)
        coi = np.maximum(coi, bits.sum(axis=1).astype(np.int32))
    if W == 1:
        obs = obs[:, :, 0].copy()
    if is_missing is False or is_missing is None:
        mis = np.zeros((N, L), dtype=np.uint8)
    else:
        mis = np.asarray(is_missing)
        if mis.shape != (L, N):
            raise ValueError(f"is_missing must be loci x samples = {(L, N)}, got {mis.shape}")
        mis = np.ascontiguousarray(mis.T.astype(np.uint8))
    return PackedData(N, L, num_alleles, obs, mis, coi)
