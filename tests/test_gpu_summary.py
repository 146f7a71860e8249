"""Posterior summaries on the device against the oracle and the reference's
own sokal() (golden vectors in tests/golden/sokal.npz, generated from user_examples/logwrite.c).

Bar: the window length m exactly; var and tau within 1e-10 relative (the reference reaches the
autocovariance through two FFTs, the kernel through direct sums: same quantity, different rounding);
counts exact and moments within 1e-12 of the oracle's two-pass formulas."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


def test_sokal_against_reference_golden(amx):
    g = cases.load_golden("sokal")
    for c in cases.SOKAL_CASES:
        x = cases.sokal_case(c)
        x0 = x.copy()
        var, tau, m = amx.sokal(x)
        gv, gt, gm = g[c[0] + "_vtm"]
        assert np.array_equal(x, x0)  # input untouched
        assert int(m[0]) == int(gm), (c[0], int(m[0]), int(gm))
        assert _rel(var[0], gv) < RTOL, (c[0], var[0], gv)
        if np.isnan(gt):
            assert np.isnan(tau[0]), c[0]
        else:
            assert _rel(tau[0], gt) < RTOL, (c[0], tau[0], gt)


def test_sokal_batched_against_oracle(amx, orc):
    """Many series in one call, every persistence regime, one CTA per series."""
    n = 2048
    x = np.stack([cases.sokal_series(100 + i, n, stay, 2 + i % 5) for i, stay in enumerate(np.linspace(0.0, 0.995, 37))])
    var, tau, m = amx.sokal(x)
    again = amx.sokal(x)
    for a, b in zip((var, tau, m), again):
        assert np.array_equal(a, b)  # fixed reduction order
    for i in range(len(x)):
        ov, ot, om = orc.sokal(x[i])
        assert int(m[i]) == om, (i, int(m[i]), om)
        assert _rel(var[i], ov) < RTOL and _rel(tau[i], ot) < RTOL, (i, tau[i], ot)


def test_sokal_rejects_bad_length(amx):
    for n in (3, 100, 6):
        with pytest.raises(amx.AmxError):
            amx.sokal(np.zeros(n))


def _population(amx, name, nchains, n_trace=0, seed=3):
    from automix_b200 import workloads as W

    wl = cases.workload(name)
    mix = W.ideal_proposal(wl)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, nchains, cases.default_init(wl, 5), seed=seed, n_trace=n_trace)
    pop.init_chains()
    return wl, pop, (T, P)


def test_trace_chain_sokal(amx, orc):
    """amx_rj_sokal over the trace chains' model-index series equals the oracle on the downloaded trace."""
    wl, pop, keep = _population(amx, "toy1", 64, n_trace=5)
    pop.sweeps(500, burning=True)
    nsweeps, nkeep = 3000, 2048
    pop.sweeps(nsweeps)
    var, tau, m = pop.sokal(nkeep)
    tr = pop.trace()
    for c in range(5):
        ov, ot, om = orc.sokal(tr["k"][c][nsweeps - nkeep:].astype(float))
        assert int(m[c]) == om
        assert _rel(var[c], ov) < RTOL and _rel(tau[c], ot) < RTOL
    assert np.all(tau > 0.5)
    with pytest.raises(amx.AmxError):
        pop.sokal(4096)  # more than was recorded


@pytest.mark.parametrize("name,nchains", [("toy1", 5000), ("toy2", 777), ("c5_rj", 1500)])
def test_population_moments(amx, orc, name, nchains):
    wl, pop, keep = _population(amx, name, nchains)
    dims = [int(d) for d in wl["dims"]]
    pop.sweeps(60, burning=True)
    pop.moments_reset()
    ks, ths, lps = [], [], []
    for rep in range(3):  # thinned accumulation: three snapshots of the population
        pop.sweeps(7)
        pop.moments_accumulate()
        s = pop.get_state()
        ks.append(s["k"]); ths.append(s["theta"]); lps.append(s["lp"])
    k, th, lp = np.concatenate(ks), np.concatenate(ths), np.concatenate(lps)
    total = 0
    for mdl, d in enumerate(dims):
        got = pop.moments(mdl, d)
        cnt, mean, cov = orc.model_moments(k, th, mdl, d)
        assert got["count"] == cnt == int(np.sum(k == mdl))
        total += cnt
        if cnt == 0:
            continue
        scale = np.maximum(1.0, np.abs(mean))
        assert np.max(np.abs(got["mean"] - mean) / scale) < 1e-12, (mdl, got["mean"], mean)
        assert abs(got["mean_lp"] - lp[k == mdl].mean()) < 1e-11 * max(1.0, abs(lp[k == mdl].mean()))
        if cnt > 1:
            cs = np.maximum(1.0, np.abs(cov).max())
            assert np.max(np.abs(got["cov"] - cov)) / cs < 1e-11, (mdl, np.max(np.abs(got["cov"] - cov)))
            assert np.array_equal(got["cov"], got["cov"].T)
    assert total == 3 * nchains
    # reproducible bit for bit: same population, same snapshots
    first = [pop.moments(mdl, d) for mdl, d in enumerate(dims)]
    pop.moments_reset()
    pop.moments_accumulate()
    a = pop.moments(0, dims[0])
    pop.moments_reset()
    pop.moments_accumulate()
    b = pop.moments(0, dims[0])
    assert a["count"] == b["count"] and np.array_equal(a["mean"], b["mean"]) and np.array_equal(a["cov"], b["cov"])
    assert first[0]["count"] >= a["count"]


def test_moments_match_the_target(amx):
    """Statistical sanity at population scale: toy1's model 0 is a two-component 1-d mixture with known mean."""
    wl, pop, keep = _population(amx, "toy1", 1 << 16, seed=11)
    pop.sweeps(300, burning=True)
    pop.moments_reset()
    for _ in range(8):
        pop.sweeps(25)
        pop.moments_accumulate()
    vis, _ = pop.collect()
    tg = wl["target"]
    off_w = 0
    tot = 0
    for mdl, d in enumerate(wl["dims"]):
        L = int(tg["ncomp"][mdl])
        w = np.asarray(tg["wt"][off_w:off_w + L], float)
        off_w += L
        got = pop.moments(mdl, int(d))
        tot += got["count"]
        if mdl == 0:
            mu = np.asarray(tg["mean"][:L], float)  # d = 1: one coordinate per component
            want = float(np.sum(w * mu) / np.sum(w))
            sd = np.sqrt(got["cov"][0, 0] / got["count"])
            assert abs(got["mean"][0] - want) < 6 * sd + 1e-3, (got["mean"][0], want, sd)
    assert tot == 8 * (1 << 16)
