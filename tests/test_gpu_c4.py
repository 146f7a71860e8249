"""BASELINE config 4 (finite mixture of normals with an unknown number of components, "enzyme-style" data) through the
whole pipeline on the device: stage 1 (adaptive RWM per model), stage 2 (mixture fit), stage 3 (population RJ sweeps).
The reference ships no such example, so there is no reference posterior to compare with; the plug-in itself is checked
against the host callback at 1e-12 (test_gpu_targets.py), stage 1 and the sweeps against the oracle on injected uniforms
(test_gpu_rwm.py, test_gpu_rj.py).  Here: the pipeline runs, is reproducible, and gives the answer the data dictate --
the sample is from TWO well separated components, so one component has no support and two or three carry the mass."""
import numpy as np
import pytest

from automix_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _pipeline(amx, nmodels, seed):
    wl = W.c4_mixnorm(nmodels=nmodels)
    T = amx.Target(wl["target"])
    ncomp, wt, mean, tri, sig, off = [], [], [], [], [], 0
    for k, d in enumerate(wl["dims"]):
        d = int(d)
        r = amx.rwm_adapt(T, k, 20000, 1, wl["init"][off:off + d], seed=seed + k)
        off += d
        xs = r["samples"][0]
        assert xs.shape == (1000 * d, d) and np.isfinite(xs).all()
        idx, _ = amx.em_draw_init(len(xs), 30, W.splitmix_uniforms_fast(seed + 100 + k, 4096))
        e = amx.em_fit(xs, idx, Lmax=30, maxit=5000)
        assert 1 <= e["L"] <= 30
        ncomp.append(e["L"]); wt.append(e["lam"]); mean.append(e["mu"].ravel()); tri.append(e["B"].ravel()); sig.append(r["sig"][0])
    mix = dict(dims=np.asarray(wl["dims"], np.int32), ncomp=np.array(ncomp, np.int32), wt=np.concatenate(wt),
               mean=np.concatenate(mean), tri=np.concatenate(tri), sig=np.concatenate(sig))
    pop = amx.RjPopulation(amx.Proposal(mix), T, 8192, wl["init"], seed=seed)
    pop.set_pk_mode(True)
    pop.init_chains()
    pop.sweeps(3000, burning=True)
    pop.collect(reset=True)
    pop.sweeps(1000)
    vis, st = pop.collect()
    p, se, _ = pop.visit_se()
    pop.close()
    return vis, p, se, ncomp, st


def test_c4_pipeline_four_models(amx):
    vis, p, se, ncomp, st = _pipeline(amx, 4, 17)   # K = 1 .. 4, d = 2, 5, 8, 11
    print("C4 fitted L", ncomp, "P(K)", np.round(p, 4), "se", np.round(se, 4), "jump acceptance", st["acc_jump"] / st["try_jump"])
    assert int(vis.sum()) == 8192 * 1000
    assert p[0] < 1e-3, "one normal component cannot carry a sample from two separated ones"
    assert p[1] + p[2] > 0.5 and np.all(se < 0.02)
    vis2, p2, _, ncomp2, _ = _pipeline(amx, 4, 17)
    assert np.array_equal(vis, vis2) and ncomp == ncomp2, "a seeded pipeline is reproducible"
