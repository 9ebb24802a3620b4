import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from moire_b200.data import pack_data
from moire_b200.simulate import simulate_data
from moire_b200.engine import Engine
for As, miss in (([60, 60, 30], 1.0), ([100, 100, 70], 1.0), ([100, 100, 70], 0.0), ([60, 60, 30], 0.0)):
    d = simulate_data(mean_coi=3, num_samples=6, epsilon_pos=.02, epsilon_neg=.1,
                      locus_freq_alphas=[np.ones(a) for a in As], internal_relatedness_alpha=.2,
                      internal_relatedness_beta=1, seed=3, missingness=miss)
    pk = pack_data(d["data"], d["is_missing"])
    eng = Engine(pk, [1.0], seed=5, burnin=1000, max_coi=10, pipeline="flat")
    eng.init_chains()
    for it in range(20):
        eng.update(2, it)
    dg = eng.diagnostics(0)
    att = dg["p_attempt"]
    print(As, "miss", miss, "W", eng.W, "attempts per locus", att.sum(axis=1), "per step", att.sum(axis=1) / 20.0,
          "min p", eng.dump_state(0)["p"][0, :As[0]].min(), flush=True)
    eng.close()
