import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from moire_b200.data import pack_data
from moire_b200.simulate import simulate_data
from moire_b200.engine import Engine
As = [100, 2, 65, 128, 64]
d = simulate_data(mean_coi=3, num_samples=30, epsilon_pos=.02, epsilon_neg=.1,
                  locus_freq_alphas=[np.ones(a) for a in As], internal_relatedness_alpha=.2,
                  internal_relatedness_beta=1, seed=3)
pk = pack_data(d["data"], d["is_missing"])
print("words", pk.mask_words, flush=True)
eng = Engine(pk, [1.0, 0.0], seed=2, burnin=30, max_coi=20)
print(eng.init_chains(), flush=True)
eng.synchronize(); print("init ok", flush=True)
st = eng.dump_state(0); print("dump ok", st["latent"][:2], flush=True)
for it in range(3):
    for kind in range(8):
        eng.update(kind, it)
        eng.synchronize()
        print("it", it, "kind", kind, "ok", flush=True)
print(eng.scalars())
