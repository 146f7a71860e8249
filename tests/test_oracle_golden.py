"""The CPU restatement against the committed golden vectors (tests/golden/*.npz), which are
outputs of the unmodified reference produced by oracle/gen_golden.py.  Runs anywhere."""
import numpy as np
import pytest

import cases


def test_helpers(orc, po):
    g = cases.load_golden("helpers")
    seed = np.array([12345], dtype=np.uint64)
    import ctypes as C

    orc.lib.orc_sdrni.argtypes = [C.POINTER(C.c_ulong)]
    orc.lib.orc_sdrand.restype = C.c_double
    orc.tape(np.zeros(0))  # clear any tape left by another test
    orc.lib.orc_tape_set(None, 0)
    orc.lib.orc_sdrni(seed.ctypes.data_as(C.POINTER(C.c_ulong)))
    got = np.array([orc.lib.orc_sdrand() for _ in range(8)])
    assert np.array_equal(got, g["sdrand_12345"])
    # SURVEY.md appendix C prints the first five
    assert np.allclose(got[:3], [0.25515066366218653, 0.10890139173535314, 0.14093326031927483], rtol=0, atol=0)
    orc.tape(cases.tape(101, 256))
    assert np.array_equal(orc.gauss(7), g["gauss7"])
    assert np.array_equal(orc.gauss(4), g["gauss4"])
    assert np.array_equal(orc.rt(5, 5), g["rt5_dof5"])
    assert np.array_equal(orc.rt(3, 1), g["rt3_dof1"])
    assert np.array_equal(orc.rt(3, 2), g["rt3_dof2"])
    assert np.array_equal(orc.perm(np.arange(6.0)), g["perm6"])
    assert orc.tape_used() == int(g["tape_used"][0])
    assert np.array_equal(orc.chol(g["chol_in"], 6), g["chol_out"])
    assert orc.det(g["chol_out"], 6) == g["det"][0]
    got = np.array([orc.lnormprob(g["ln_mu"], g["chol_out"], x) for x in g["ln_x"]])
    assert np.array_equal(got, g["ln_out"])
    got = np.array([orc.loggamma(v) for v in g["loggamma_x"]])
    assert np.allclose(got, g["loggamma"], rtol=4e-16, atol=4e-16)
    assert abs(orc.ltprob(5, 0.7) - g["ltprob_5_0.7"][0]) < 1e-15
    # known answers printed in SURVEY.md appendix C
    assert np.allclose(orc.chol(np.array([4.0, 2, 5, -1, 0.5, 3]), 3), [2, 1, 2, -0.5, 0.5, 1.5811388300841898], rtol=0, atol=1e-16)
    B = orc.chol(np.array([4.0, 2, 5, -1, 0.5, 3]), 3)
    assert orc.lnormprob([0.5, -1, 2], B, [1, 0.25, -0.75]) == -6.4106303266709856


@pytest.mark.parametrize("name", ["toy1", "toy2"])
def test_pipeline_against_reference_outputs(orc, ht, name):
    g = cases.load_golden(name)
    wl = cases.workload(name)
    seed = int(g["seed"][0])
    init = cases.default_init(wl, seed)
    assert np.array_equal(init, g["init"])
    mix, stages = cases.fit_pipeline(orc, ht, wl, init, seed)
    for key in ("dims", "ncomp", "wt", "mean", "tri", "sig"):
        assert np.array_equal(mix[key], g["mix_" + key]), key
    for k, st in enumerate(stages):
        r, e = st["rwm"], st["em"]
        assert np.array_equal(r["sig"], g[f"rwm{k}_sig"])
        assert np.array_equal(r["samples"][:64], g[f"rwm{k}_samples_head"])
        assert np.array_equal(r["samples"][-64:], g[f"rwm{k}_samples_tail"])
        assert np.array_equal(r["samples"].sum(0), g[f"rwm{k}_samples_sum"])
        assert np.array_equal(r["sig_trace"][::10], g[f"rwm{k}_sig_trace"])
        for key in ("trace_L", "trace_loglik", "trace_cost", "trace_ann"):
            assert np.array_equal(e[key], g[f"em{k}_{key}"]), key
    ptr = ht.select(wl["target"])
    dmax = int(max(wl["dims"]))
    n1, n2 = (int(v) for v in g["rj_nsweeps"])
    orc.tape(cases.tape(seed * 1000 + 999, cases.rj_tape_len(dmax, n1 + n2) + 8))
    s0 = orc.chain_init(wl["dims"], init, ptr)
    assert s0["k"] == int(g["rj_init_k"][0]) and s0["lp"] == g["rj_init_lp"][0]
    a = orc.rj_sweeps(mix, ptr, s0, n1, burning=True)
    b = orc.rj_sweeps(mix, ptr, a["state"], n2)
    for tag, r in (("burn", a), ("run", b)):
        for key in ("k", "lp", "theta", "pk", "counters", "visits"):
            assert np.array_equal(r[key], g[f"rj_{tag}_{key}"]), (tag, key)
    assert orc.tape_used() == int(g["rj_tape_used"][0])


def test_em_against_reference_outputs(orc):
    g = cases.load_golden("em3d")
    for maxit in (0, 1, 2, 5, 5000):
        orc.tape(cases.tape(77, 4096))
        e = orc.fit_mixture(g["x"], Lmax=12, maxit=maxit)
        for key in ("lam", "mu", "B", "trace_L", "trace_loglik", "trace_cost", "trace_ann"):
            assert np.array_equal(e[key], g[f"m{maxit}_{key}"]), (maxit, key)
    fa = orc.fit_autorj(g["x"])
    assert np.array_equal(fa["mu"], g["autorj_mu"]) and np.array_equal(fa["B"], g["autorj_B"])


def test_tape_overrun_is_reported(orc, ht):
    wl = cases.workload("toy1")
    ptr = ht.select(wl["target"])
    orc.tape(cases.tape(1, 10))
    s0 = orc.chain_init(wl["dims"], np.zeros(3), ptr)
    g = cases.load_golden("toy1")
    mix = {k[4:]: g[k] for k in g if k.startswith("mix_")}
    orc.rj_sweeps(mix, ptr, s0, 50)
    assert orc.tape_overrun()


def test_sokal_oracle_against_reference_golden(orc):
    """orc_sokal (own radix-2 transform) against the reference's sokal() outputs (logwrite.c:354-403)."""
    g = cases.load_golden("sokal")
    for c in cases.SOKAL_CASES:
        x = cases.sokal_case(c)
        var, tau, m = orc.sokal(x)
        gv, gt, gm = g[c[0] + "_vtm"]
        assert m == int(gm), c[0]
        assert abs(var - gv) <= 1e-12 * max(1.0, abs(gv)), c[0]
        if np.isnan(gt):
            assert np.isnan(tau)
        else:
            assert abs(tau - gt) <= 1e-10 * max(1.0, abs(gt)), (c[0], tau, gt)
    with pytest.raises(ValueError):
        orc.sokal(np.zeros(100))


def test_model_moments_oracle(orc):
    rng = np.random.default_rng(3)
    k = rng.integers(0, 3, 500).astype(np.int32)
    th = rng.normal(size=(500, 4)) * [1, 2, 3, 4] + [10, -5, 0, 1]
    for mdl, d in ((0, 2), (1, 4), (2, 1)):
        cnt, mean, cov = orc.model_moments(k, th, mdl, d)
        sel = th[k == mdl][:, :d]
        assert cnt == len(sel)
        assert np.allclose(mean, sel.mean(0), rtol=1e-13, atol=1e-13)
        assert np.allclose(cov, np.cov(sel.T).reshape(d, d), rtol=1e-12, atol=1e-13)
