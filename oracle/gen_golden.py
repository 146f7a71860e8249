"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).  TEST INFRASTRUCTURE.

Run in the container that holds /root/reference:
    make -C oracle all && python oracle/gen_golden.py
The reference's own tests hold no per-step vectors, so these files --
outputs of the reference itself on seeded inputs -- are what pins the oracle and, through it,
the CUDA path.  Inputs are SplitMix64 tapes identified by seed; only seeds and outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import pyoracle as po  # noqa: E402


def main():
    assert po.have_ref(), "oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists"
    os.makedirs(cases.GOLDEN_DIR, exist_ok=True)
    ref = po.Checker("ref")
    ht = po.HostTargets()
    pristine = po.RefPristine()

    # ---- helper known answers -----------------------------------------------------------------
    g = {}
    pristine.seed(12345)
    g["sdrand_12345"] = pristine.uniforms(8)
    u = cases.tape(101, 256)
    ref.tape(u)
    g["gauss7"] = ref.gauss(7)
    g["gauss4"] = ref.gauss(4)
    g["rt5_dof5"] = ref.rt(5, 5)
    g["rt3_dof1"] = ref.rt(3, 1)
    g["rt3_dof2"] = ref.rt(3, 2)
    g["perm6"] = ref.perm(np.arange(6.0))
    g["tape_used"] = np.array([ref.tape_used()])
    rng = np.random.default_rng(7)
    A = rng.normal(size=(6, 6))
    S = A @ A.T + 6 * np.eye(6)
    g["chol_in"] = po.pack_lower(np.tril(S))
    g["chol_out"] = ref.chol(g["chol_in"], 6)
    g["det"] = np.array([ref.det(g["chol_out"], 6)])
    g["ln_mu"] = rng.normal(size=6)
    g["ln_x"] = rng.normal(size=(16, 6)) * 2
    g["ln_out"] = np.array([ref.lnormprob(g["ln_mu"], g["chol_out"], x) for x in g["ln_x"]])
    g["loggamma_x"] = np.array([0.3, 0.5, 1.0, 1.5, 2.5, 3.0, 7.25, 33.0, 100.5])
    g["loggamma"] = np.array([ref.loggamma(v) for v in g["loggamma_x"]])
    g["ltprob_5_0.7"] = np.array([ref.ltprob(5, 0.7)])
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "helpers.npz"), **g)

    # ---- stages 1-3 on toy1 and toy2 -------------------------------------------------------------
    for name, seed, nrj in (("toy1", 11, 2000), ("toy2", 12, 1500)):
        wl = cases.workload(name)
        init = cases.default_init(wl, seed)
        mix, stages = cases.fit_pipeline(ref, ht, wl, init, seed)
        out = {"init": init, "seed": np.array([seed])}
        for key in ("dims", "ncomp", "wt", "mean", "tri", "sig"):
            out["mix_" + key] = mix[key]
        for k, st in enumerate(stages):
            r, e = st["rwm"], st["em"]
            out[f"rwm{k}_sig"] = r["sig"]
            out[f"rwm{k}_samples_head"] = r["samples"][:64]
            out[f"rwm{k}_samples_tail"] = r["samples"][-64:]
            out[f"rwm{k}_samples_sum"] = r["samples"].sum(0)
            out[f"rwm{k}_sig_trace"] = r["sig_trace"][::10]
            out[f"em{k}_trace_L"] = e["trace_L"]
            out[f"em{k}_trace_loglik"] = e["trace_loglik"]
            out[f"em{k}_trace_cost"] = e["trace_cost"]
            out[f"em{k}_trace_ann"] = e["trace_ann"]
            if wl["dims"][k] <= 2:
                out[f"em{k}_samples"] = r["samples"]  # inputs for the GPU EM parity test
        ptr = ht.select(wl["target"])
        dmax = int(max(wl["dims"]))
        ref.tape(cases.tape(seed * 1000 + 999, cases.rj_tape_len(dmax, nrj) + 8))
        s0 = ref.chain_init(wl["dims"], init, ptr)
        a = ref.rj_sweeps(mix, ptr, s0, nrj // 2, burning=True)
        b = ref.rj_sweeps(mix, ptr, a["state"], nrj - nrj // 2)
        assert not ref.tape_overrun()
        out["rj_nsweeps"] = np.array([nrj // 2, nrj - nrj // 2])
        out["rj_init_k"] = np.array([s0["k"]])
        out["rj_init_lp"] = np.array([s0["lp"]])
        for tag, r in (("burn", a), ("run", b)):
            out[f"rj_{tag}_k"] = r["k"]
            out[f"rj_{tag}_lp"] = r["lp"]
            out[f"rj_{tag}_theta"] = r["theta"]
            out[f"rj_{tag}_pk"] = r["pk"]
            out[f"rj_{tag}_counters"] = r["counters"]
            out[f"rj_{tag}_visits"] = r["visits"]
        out["rj_tape_used"] = np.array([ref.tape_used()])
        np.savez_compressed(os.path.join(cases.GOLDEN_DIR, f"{name}.npz"), **out)
        print(name, "ncomp", mix["ncomp"], "visits", b["visits"], "tape used", ref.tape_used())

    # ---- EM on a 3-d three-cluster sample set, iteration by iteration ---------------------------------
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.normal(size=(700, 3)) * [1, .5, 2] + [4, 0, -2],
                        rng.normal(size=(900, 3)) @ np.array([[1, 0, 0], [.6, .8, 0], [-.3, .2, .5]]).T + [-3, 2, 1],
                        rng.normal(size=(400, 3)) * .4 + [0, -5, 5]])
    out = {"x": x}
    for maxit in (0, 1, 2, 5, 5000):
        ref.tape(cases.tape(77, 4096))
        e = ref.fit_mixture(x, Lmax=12, maxit=maxit)
        for key in ("lam", "mu", "B", "trace_L", "trace_loglik", "trace_cost", "trace_ann"):
            out[f"m{maxit}_{key}"] = e[key]
    fa = ref.fit_autorj(x)
    out["autorj_mu"], out["autorj_B"] = fa["mu"], fa["B"]
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "em3d.npz"), **out)
    print("em3d: iters", len(out["m5000_trace_L"]), "L", len(out["m5000_lam"]))
    for f in sorted(os.listdir(cases.GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(cases.GOLDEN_DIR, f)))


def sokal_golden():
    """tests/golden/sokal.npz: the reference's sokal() (user_examples/logwrite.c:354-403, compiled as it lies
    into oracle/_ref/libref_logwrite.so) on the seeded series of cases.SOKAL_CASES."""
    assert po.have_ref_logwrite(), "oracle/_ref/libref_logwrite.so is missing: run `make -C oracle ref`"
    lw = po.RefLogwrite()
    out = {}
    for c in cases.SOKAL_CASES:
        x = cases.sokal_case(c)
        var, tau, m, rho = lw.sokal(x)
        out[c[0] + "_vtm"] = np.array([var, tau, float(m)])
        out[c[0] + "_rho8"] = rho[:8]  # the first autocorrelations the reference left in its input array
        print(c[0], var, tau, m)
    np.savez_compressed(os.path.join(cases.GOLDEN_DIR, "sokal.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sokal":
        sokal_golden()
    else:
        main()
        sokal_golden()
