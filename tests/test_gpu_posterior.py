"""Posterior model probabilities against the reference's sampler, within Monte-Carlo error (north_star level 2).

Fixtures: tests/golden/coalmine_posterior.npz, produced from the UNMODIFIED reference by
oracle/gen_golden_posterior.py (the truth from 9.6e7 sweeps of its fixed-pk chain, its adaptive chain at 4.8e7
sweeps, and its adaptive chain run as 4000 short chains).  The Monte-Carlo error of a population estimate is
computed from the between-chain variance of the traced chains (chains are independent):
se = sd_over_chains(time-averaged frequency) / sqrt(number of chains).  Every bound is 4 combined standard errors.
"""
import os

import numpy as np
import pytest

import cases
from automix_b200 import workloads as W

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coalmine_posterior.npz")


def _run(amx, wl, mix, init, nchains, nburn, nsweep, population, seed, ntrace=256):
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, nchains, init, seed=seed, n_trace=ntrace)
    pop.set_pk_mode(population)
    pop.init_chains()
    pop.sweeps(nburn, burning=True)
    pop.collect(reset=True)
    pop.sweeps(nsweep)
    vis, st = pop.collect()
    assert int(vis.sum()) == nchains * nsweep
    p = vis / vis.sum()
    p_g, se_g, ngroups = pop.visit_se()  # the library's own error estimate: spread between 64 groups of chains
    assert np.allclose(p_g, p, atol=1e-12) and ngroups == min(64, (nchains + 127) // 128)
    k = pop.trace()["k"]  # [ntrace, nsweep]
    nm = len(wl["dims"])
    freq = np.stack([(k == m).mean(1) for m in range(nm)], 1)  # per traced chain
    se = freq.std(0, ddof=1) / np.sqrt(nchains)
    # two estimates of the same standard error (64 groups: ~9 % relative noise; 256 chains: ~4.5 %)
    assert np.all(se_g < 1.6 * se + 1e-6) and np.all(se_g > se / 1.6 - 1e-6), (se_g, se)
    out = dict(p=p, se=se, se_groups=se_g, chain_sd=freq.std(0, ddof=1), pk=pop.pk_shared()[0], stats=st)
    pop.close()
    return out


def test_coalmine_population_mode_matches_the_reference_posterior(amx):
    """BASELINE config 3.  16384 chains x (2000 burn-in + 2000 sweeps), shared pk adapted from the population's visit
    histogram, on the proposal the reference fitted: every P(k) within 4 standard errors of the reference's
    long-run posterior (0.0581 0.2513 0.2973 0.2336 0.1163 0.0434)."""
    g = np.load(GOLD)
    mix = {k[4:]: g[k] for k in g.files if k.startswith("mix_")}
    wl = W.coalmine()
    r = _run(amx, wl, mix, wl["init"], 16384, 2000, 2000, True, seed=1851)
    tol = 4.0 * np.sqrt(r["se"] ** 2 + g["truth_se"] ** 2)
    print("population mode P(k)", np.round(r["p"], 4), "se", np.round(r["se"], 5), "truth", np.round(g["truth_p"], 4),
          "z", np.round((r["p"] - g["truth_p"]) / (tol / 4), 2), "shared pk", np.round(r["pk"], 4))
    assert np.all(r["se"] < 0.0012)  # the test can fail: 4 se is at most 0.005 on probabilities of 0.04 .. 0.30
    assert np.all(np.abs(r["p"] - g["truth_p"]) < tol), (r["p"], g["truth_p"], tol)
    # the adapted pk itself tracks the posterior model probabilities (that is what the rule converges to)
    assert np.all(np.abs(r["pk"] - g["truth_p"]) < 0.03), r["pk"]
    assert abs(r["pk"].sum() - 1.0) < 1e-12


def test_coalmine_per_chain_mode_reproduces_the_reference_under_the_same_schedule(amx):
    """The kernel against the reference's ESTIMATOR: per-chain pk adaptation (the reference's rule) with the schedule
    many-short-chains gives the same -- biased -- probabilities as 4000 chains of the reference itself on the same
    proposal and schedule (P(k=5) = 0.101 against a posterior of 0.116), within 4 standard errors; and that is
    measurably not the posterior."""
    g = np.load(GOLD)
    mix = {k[4:]: g[k] for k in g.files if k.startswith("mix_")}
    wl = W.coalmine()
    r = _run(amx, wl, mix, wl["init"], 16384, int(g["sched_runs"][1]), int(g["sched_runs"][2]), False, seed=7)
    tol = 4.0 * np.sqrt(r["se"] ** 2 + g["sched_se"] ** 2)
    print("per-chain mode P(k)", np.round(r["p"], 4), "reference, same schedule", np.round(g["sched_p"], 4),
          "z", np.round((r["p"] - g["sched_p"]) / (tol / 4), 2))
    assert np.all(np.abs(r["p"] - g["sched_p"]) < tol), (r["p"], g["sched_p"], tol)
    # between-chain spread agrees with the reference's chains too (a second moment of the same estimator)
    assert np.all(np.abs(r["chain_sd"] / g["sched_chain_sd"] - 1.0) < 0.25), (r["chain_sd"], g["sched_chain_sd"])
    assert abs(r["p"][4] - g["truth_p"][4]) > 0.008  # the schedule's bias is real and resolved by this test


def test_toy2_population_mode_against_the_known_model_probabilities(amx):
    """usertoy2.c: model probabilities 0.5 / 0.25 / 0.125 / 0.0625 / 0.0625 by construction."""
    wl = cases.workload("toy2")
    mix = W.ideal_proposal(wl)
    r = _run(amx, wl, mix, cases.default_init(wl, 3), 32768, 1000, 1000, True, seed=99)
    print("toy2 P(k)", np.round(r["p"], 4), "se", np.round(r["se"], 5))
    assert np.all(np.abs(r["p"] - wl["true_probs"]) < 4.0 * r["se"] + 1e-4), (r["p"], r["se"])
    assert np.all(r["se"] < 0.002)


def test_toy1_both_modes_against_the_known_model_probabilities(amx):
    """usertoy1.c:96-100: 0.3 / 0.7 (thesis p.167: 0.2997 / 0.7003)."""
    wl = cases.workload("toy1")
    mix = W.ideal_proposal(wl)
    for population in (True, False):
        r = _run(amx, wl, mix, cases.default_init(wl, 3), 1 << 16, 500, 1000, population, seed=2024)
        print("toy1", "population" if population else "per-chain", np.round(r["p"], 5), "se", np.round(r["se"], 5))
        if population:
            assert abs(r["p"][0] - 0.3) < 4.0 * r["se"][0] + 1e-4, (r["p"], r["se"])
        else:  # two models, fast mixing: the per-chain rule's transient is small here, but it is not asserted at 4 se
            assert abs(r["p"][0] - 0.3) < 0.004


def test_population_pk_is_partitioned_by_segments_not_by_calls(amx):
    """The shared pk moves after every `segment` sweeps of a call: a call of 100 sweeps and four calls of 25 give the
    same chains bit for bit (segment = 25), and the histogram adds up."""
    wl = cases.workload("toy1")
    mix = W.ideal_proposal(wl)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    outs = []
    for split in ((100,), (25, 25, 25, 25), (50, 50)):
        pop = amx.RjPopulation(P, T, 4096, cases.default_init(wl, 3), seed=5)
        pop.set_pk_mode(True, 25)
        pop.init_chains()
        pop.sweeps(50, burning=True)
        for n in split:
            pop.sweeps(n)
        vis, st = pop.collect()
        outs.append((vis.copy(), pop.get_state()["theta"].copy(), pop.pk_shared()[0]))
        pop.close()
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
    assert outs[0][0].sum() == 4096 * 150
