"""Coal-mining posterior model probabilities from the population sampler for proposals fitted on the device with
different seeds: z-scores against the reference's long-run posterior (tests/golden/coalmine_posterior.npz), at two
burn-in lengths.  Separates slow mixing under a poor proposal from a biased kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from automix_b200 import _lib as amx, workloads as W

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "coalmine_posterior.npz"))
truth, tse = g["truth_p"], g["truth_se"]
wl = W.coalmine()
T = amx.Target(wl["target"])
seeds = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
for s in seeds:
    ncomp, wt, mean, tri, sig, off = [], [], [], [], [], 0
    for k, d in enumerate(wl["dims"]):
        d = int(d)
        r = amx.rwm_adapt(T, k, 100000, 1, wl["init"][off:off + d], seed=1000 * s + k)
        off += d
        xs = r["samples"][0]
        idx, _ = amx.em_draw_init(len(xs), 30, W.splitmix_uniforms_fast(77 * s + k, 4096))
        e = amx.em_fit(xs, idx, Lmax=30, maxit=5000)
        ncomp.append(e["L"]); wt.append(e["lam"]); mean.append(e["mu"].ravel()); tri.append(e["B"].ravel()); sig.append(r["sig"][0])
    mix = dict(dims=np.asarray(wl["dims"], np.int32), ncomp=np.array(ncomp, np.int32), wt=np.concatenate(wt),
               mean=np.concatenate(mean), tri=np.concatenate(tri), sig=np.concatenate(sig))
    P = amx.Proposal(mix)
    for burn in (10000, 100000):
        pop = amx.RjPopulation(P, T, 16384, wl["init"], seed=31 + s)
        pop.set_pk_mode(True)
        pop.init_chains()
        pop.sweeps(burn, burning=True)
        pop.collect(reset=True)
        pop.sweeps(4000)
        vis, st = pop.collect()
        p, se, _ = pop.visit_se()
        z = (p - truth) / np.sqrt(se ** 2 + tse ** 2)
        print(f"seed {s} L={ncomp} burn {burn}: P {np.round(p, 4)} se {np.round(se, 5)} z {np.round(z, 1)} jump acc {st['acc_jump'] / st['try_jump']:.3f}", flush=True)
        pop.close()
