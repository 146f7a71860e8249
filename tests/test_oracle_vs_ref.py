"""The CPU restatement (oracle/amx_oracle.c) against the unmodified reference compiled from
/root/reference (oracle/_ref), bit for bit, on seeded tapes.  Skipped where _ref is absent."""
import numpy as np
import pytest

import cases


def test_helpers_bit_exact(orc, ref, po):
    u = cases.tape(3, 4096)
    orc.tape(u)
    ref.tape(u)
    for n in (1, 2, 3, 7, 8):
        assert np.array_equal(orc.gauss(n), ref.gauss(n))
    for n, dof in ((5, 5), (3, 1), (3, 2), (4, 7), (1, 3)):
        assert np.array_equal(orc.rt(n, dof), ref.rt(n, dof))
    for n in (2, 5, 9):
        assert np.array_equal(orc.perm(np.arange(float(n))), ref.perm(np.arange(float(n))))
    for s in (0.4, 1.0, 1.7, 3.5):
        assert orc.rgamma(s) == ref.rgamma(s)
    assert orc.tape_used() == ref.tape_used()
    rng = np.random.default_rng(0)
    for d in (1, 2, 5, 13, 20):
        A = rng.normal(size=(d, d))
        Sp = po.pack_lower(np.tril(A @ A.T + d * np.eye(d)))
        Bo, Br = orc.chol(Sp, d), ref.chol(Sp, d)
        assert np.array_equal(Bo, Br)
        assert orc.det(Bo, d) == ref.det(Bo, d)
        mu = rng.normal(size=d)
        for _ in range(5):
            x = rng.normal(size=d) * 3
            assert orc.lnormprob(mu, Bo, x) == ref.lnormprob(mu, Bo, x)
    for v in (0.3, 0.5, 1.0, 2.5, 7.25, 33.0, 100.5):  # lgamma vs the reference's Cody-Hillstrom
        assert abs(orc.loggamma(v) - ref.loggamma(v)) <= 4e-16 * max(1.0, abs(ref.loggamma(v)))
    assert abs(orc.ltprob(5, 0.7) - ref.ltprob(5, 0.7)) < 1e-15


def test_workload_targets_match_reference_examples(po, ht):
    if not po.have_ref():
        pytest.skip("no oracle/_ref")
    ru = po.RefUserTargets()
    rng = np.random.default_rng(1)
    for name, nm, dmax, scale in (("toy1", 2, 2, 4.0), ("toy2", 5, 5, 6.0)):
        ht.select(cases.workload(name)["target"])
        for _ in range(500):
            k = int(rng.integers(0, nm))
            x = rng.normal(size=dmax) * scale
            a, b = ht.logpost(k, x), ru.eval(name, k, x)
            if np.isfinite(b):
                assert abs(a - b) <= 1e-14 * max(1.0, abs(b))
    wl = cases.workload("coalmine")
    ht.select(wl["target"])
    off = 0
    for k in range(6):
        d = 2 * k + 3
        x0 = wl["init"][off:off + d]
        off += d
        assert np.array_equal(x0, ru.cpt_init(k, d))
        for _ in range(100):
            x = x0 * (1 + 0.3 * rng.normal(size=d))
            a, b = ht.logpost(k, x), ru.eval("cpt", k, x)
            assert abs(a - b) <= 1e-14 * max(1.0, abs(b))


@pytest.mark.parametrize("name,seed", [("toy1", 21), ("toy2", 22)])
def test_stages_bit_exact(orc, ref, ht, name, seed):
    wl = cases.workload(name)
    init = cases.default_init(wl, seed)
    mo, so = cases.fit_pipeline(orc, ht, wl, init, seed)
    mr, sr = cases.fit_pipeline(ref, ht, wl, init, seed)
    for key in mo:
        assert np.array_equal(mo[key], mr[key]), key
    for a, b in zip(so, sr):
        for key in ("samples", "sig", "sig_trace"):
            assert np.array_equal(a["rwm"][key], b["rwm"][key]), key
        assert np.array_equal(a["rwm"]["acc_trace"], b["rwm"]["acc_trace"], equal_nan=True)
        for key in ("trace_L", "trace_loglik", "trace_cost", "trace_ann"):
            assert np.array_equal(a["em"][key], b["em"][key]), key
        fa, fb = orc.fit_autorj(a["rwm"]["samples"]), ref.fit_autorj(a["rwm"]["samples"])
        assert all(np.array_equal(fa[q], fb[q]) for q in fa)
    ptr = ht.select(wl["target"])
    dmax = int(max(wl["dims"]))
    for dof, perm in ((0, 0), (4, 0), (0, 1), (3, 1)):
        u = cases.tape(seed + 100 * dof + perm, 3 * cases.rj_tape_len(dmax, 1200))
        res = []
        for chk in (orc, ref):
            chk.tape(u)
            s0 = chk.chain_init(wl["dims"], init, ptr)
            a = chk.rj_sweeps(mo, ptr, s0, 400, burning=True, do_perm=perm, dof=dof)
            b = chk.rj_sweeps(mo, ptr, a["state"], 800, do_perm=perm, dof=dof)
            assert not chk.tape_overrun()
            res.append((s0, a, b, chk.tape_used()))
        (s0o, ao, bo, uo), (s0r, ar, br, ur) = res
        assert s0o["k"] == s0r["k"] and s0o["lp"] == s0r["lp"] and uo == ur
        for x, y in ((ao, ar), (bo, br)):
            for key in ("k", "lp", "theta", "pk", "counters", "visits"):
                assert np.array_equal(x[key], y[key]), (dof, perm, key)


def test_em_iteration_by_iteration(orc, ref):
    g = cases.load_golden("em3d")
    for maxit in (0, 1, 2, 3, 7):
        u = cases.tape(77, 4096)
        orc.tape(u)
        a = orc.fit_mixture(g["x"], Lmax=12, maxit=maxit)
        ref.tape(u)
        b = ref.fit_mixture(g["x"], Lmax=12, maxit=maxit)
        for key in ("lam", "mu", "B", "trace_L", "trace_loglik", "trace_cost", "trace_ann"):
            assert np.array_equal(a[key], b[key]), (maxit, key)


def test_sokal_oracle_vs_reference(po, orc):
    """Fresh series (not the committed ones) through the reference's sokal() compiled as it lies."""
    if not po.have_ref_logwrite():
        pytest.skip("oracle/_ref/libref_logwrite.so not built")
    lw = po.RefLogwrite()
    for i, (n, stay) in enumerate([(8, 0.5), (128, 0.0), (512, 0.8), (2048, 0.95), (16384, 0.98), (65536, 0.9)]):
        x = cases.sokal_series(500 + i, n, stay, 4)
        rv, rt, rm, rho = lw.sokal(x)
        ov, ot, om = orc.sokal(x)
        assert om == rm, (n, om, rm)
        assert abs(ov - rv) <= 1e-12 * max(1.0, abs(rv))
        assert abs(ot - rt) <= 1e-10 * max(1.0, abs(rt)), (n, ot, rt)
        assert abs(rho[0] - 1.0) < 1e-12
