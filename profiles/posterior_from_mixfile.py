"""Population sampler on a proposal read from a <stem>_mix.data file (amx_sampler_save_proposal / the reference's
writer), at increasing burn-in lengths: does a poor proposal converge to the reference posterior, just slowly?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from automix_b200 import _lib as amx, workloads as W
from gen_golden_posterior import read_mix

g = np.load(os.path.join(ROOT, "tests", "golden", "coalmine_posterior.npz"))
truth, tse = g["truth_p"], g["truth_se"]
wl = W.coalmine()
T = amx.Target(wl["target"])
mix = read_mix(sys.argv[1])
print("L", mix["ncomp"], "sig model 1", np.round(mix["sig"][3:8], 5))
P = amx.Proposal(mix)
for burn in [int(a) for a in sys.argv[2:]] or [10000, 100000, 400000]:
    pop = amx.RjPopulation(P, T, 16384, wl["init"], seed=5)
    pop.set_pk_mode(True)
    pop.init_chains()
    pop.sweeps(burn, burning=True)
    pop.collect(reset=True)
    pop.sweeps(4000)
    vis, st = pop.collect()
    p, se, _ = pop.visit_se()
    z = (p - truth) / np.sqrt(se ** 2 + tse ** 2)
    print(f"burn {burn}: P {np.round(p, 4)} se {np.round(se, 5)} z {np.round(z, 1)} jump acc {st['acc_jump'] / st['try_jump']:.3f}", flush=True)
    pop.close()
