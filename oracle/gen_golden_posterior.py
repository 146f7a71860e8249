"""Posterior model probabilities of the coal-mining example from the UNMODIFIED reference -> tests/golden/coalmine_posterior.npz.
TEST INFRASTRUCTURE; runs only where /root/reference exists (oracle/_ref/ref_population, built by oracle/Makefile).

Three things are recorded (all from the reference's own library, generator and usercpt.c log-posterior):

  truth_p, truth_se   16 independent chains x 6e6 sweeps with doAdapt = 0 (fixed uniform jump probabilities: a plain
                      Metropolis-Hastings chain, no adaptation transient).  The posterior model probabilities to
                      +- 3e-4.  The reference's default adaptive chain converges to the same numbers, slowly:
  adapt_p, adapt_se   8 chains x 6e6 sweeps with the reference's defaults (doAdapt = 1).
  sched_p, sched_se   the reference's adaptive chain run with the POPULATION SCHEDULE of the GPU tests: 4000
                      independent chains x (2000 burn-in + 2000 sweeps), all on the proposal fitted with seed 1851
                      (mix_*: that proposal in the flat layout of include/amx_layout.h).  Its deviation from truth_p
                      (P(k=5) 0.101 vs 0.116) is the finite-time bias of per-chain pk adaptation, a property of the
                      reference's estimator under that schedule -- the per-chain mode of the GPU kernel must reproduce
                      it, the population mode must not show it.

usage: python oracle/gen_golden_posterior.py [--reuse]     (--reuse: parse outputs already in oracle/_build)
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "_ref", "ref_population")
OUT = os.path.join(HERE, "_build")
NPROC = 8


def run_wave(jobs):
    """jobs: list of (tag, env, args).  Runs them NPROC at a time."""
    for i in range(0, len(jobs), NPROC):
        procs = []
        for tag, env, args in jobs[i:i + NPROC]:
            fo = open(os.path.join(OUT, tag + ".txt"), "w")
            fe = open(os.path.join(OUT, tag + ".err"), "w")
            procs.append(subprocess.Popen([BIN] + [str(a) for a in args], stdout=fo, stderr=fe, env=dict(os.environ, **env)))
        for p in procs:
            assert p.wait() == 0


def rows(tags, nsweep):
    out = []
    for t in tags:
        for line in open(os.path.join(OUT, t + ".txt")):
            v = line.split()
            out.append([int(x) / nsweep for x in v[1:7]])
    return np.array(out)


def read_mix(path):
    tok = open(path).read().split()
    it = iter(tok)
    nm = int(next(it))
    dims = [int(next(it)) for _ in range(nm)]
    ncomp, wt, mean, tri, sig = [], [], [], [], []
    for d in dims:
        sig += [float(next(it)) for _ in range(d)]
        L = int(next(it))
        ncomp.append(L)
        for _ in range(L):
            wt.append(float(next(it)))
            mean += [float(next(it)) for _ in range(d)]
            tri += [float(next(it)) for _ in range(d * (d + 1) // 2)]
    return dict(dims=np.array(dims, np.int32), ncomp=np.array(ncomp, np.int32), wt=np.array(wt), mean=np.array(mean),
                tri=np.array(tri), sig=np.array(sig))


def main():
    reuse = "--reuse" in sys.argv
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
    na = [f"na_{w}_{i}" for w in (1, 2) for i in range(1, 9)]
    ad = [f"ad_{i}" for i in range(1, 9)]
    pc = [f"popc_{i}" for i in range(1, 9)]
    if not reuse:
        run_wave([(t, {"NOADAPT": "1"}, [1, 10000, 6000000, 200 + 10 * int(t.split("_")[1]) + int(t.split("_")[2]), 0]) for t in na])
        run_wave([(t, {}, [1, 10000, 6000000, 300 + int(t.split("_")[1]), 0]) for t in ad])
        run_wave([(t, {}, [500, 2000, 2000, 1851, 5000 + int(t.split("_")[1])] +
                   ([os.path.join(OUT, "cpt_mix_1851.data")] if t == "popc_1" else [])) for t in pc])
    A, B, S = rows(na, 6000000), rows(ad, 6000000), rows(pc, 2000)
    mix = read_mix(os.path.join(OUT, "cpt_mix_1851.data"))
    se = lambda a: a.std(0, ddof=1) / np.sqrt(len(a))
    out = dict(truth_p=A.mean(0), truth_se=se(A), truth_runs=np.array([len(A), 6000000]),
               adapt_p=B.mean(0), adapt_se=se(B), adapt_runs=np.array([len(B), 6000000]),
               sched_p=S.mean(0), sched_se=se(S), sched_chain_sd=S.std(0, ddof=1),
               sched_runs=np.array([len(S), 2000, 2000]), **{"mix_" + k: v for k, v in mix.items()})
    np.savez(os.path.join(ROOT, "tests", "golden", "coalmine_posterior.npz"), **out)
    for k in ("truth", "adapt", "sched"):
        print(k, np.round(out[k + "_p"], 5), "+-", np.round(out[k + "_se"], 5))
    print("fitted L", mix["ncomp"])


if __name__ == "__main__":
    main()
