"""RJ sweeps alone (C2 toy1, reference-fitted proposal) -- the command profiled by ncu for the K3 kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from automix_b200 import _lib as amx, workloads as W
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50
name = sys.argv[3] if len(sys.argv) > 3 else "toy1"
wl = getattr(W, name)()
if name in ("toy1", "toy2"):
    g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz")))
    mix = {k[4:]: g[k] for k in g if k.startswith("mix_")}
    init = g["init"]
elif wl["target"]["kind"] == "gaussmix":
    mix, init = W.ideal_proposal(wl), wl["init"]
else:  # no closed-form proposal (coal-mining): run stages 1-2 on the device, as the pipeline does
    T0 = amx.Target(wl["target"])
    init = wl["init"]
    ncomp, wt, mean, tri, sig, off = [], [], [], [], [], 0
    for k, d in enumerate(wl["dims"]):
        d = int(d)
        r = amx.rwm_adapt(T0, k, 1000, 1, init[off:off + d], seed=11 + k)
        off += d
        xs = r["samples"][0]
        idx, _ = amx.em_draw_init(len(xs), 30, W.splitmix_uniforms_fast(40 + k, 4096))
        e = amx.em_fit(xs, idx, Lmax=30, maxit=300)
        ncomp.append(e["L"]); wt.append(e["lam"]); mean.append(e["mu"].ravel()); tri.append(e["B"].ravel()); sig.append(r["sig"][0])
    mix = dict(dims=np.asarray(wl["dims"], np.int32), ncomp=np.array(ncomp, np.int32), wt=np.concatenate(wt),
               mean=np.concatenate(mean), tri=np.concatenate(tri), sig=np.concatenate(sig))
T, P = amx.Target(wl["target"]), amx.Proposal(mix)
pop = amx.RjPopulation(P, T, C, init, seed=1)
pop.init_chains()
pop.sweeps(100, burning=True)
pop.collect(reset=True)
for _ in range(3):
    pop.sweeps(S)
vis, st = pop.collect()
print(f"{name}: C={C} S={S} kernel_ms/launch={st['kernel_ms']/3:.3f} chain-sweeps/s={3*C*S/(st['kernel_ms']*1e-3):.4g} flops/sweep={st['flops']/(3*C*S):.1f} p={np.round(vis/vis.sum(),4)}")
