"""Seeded parity cases shared by oracle/gen_golden.py and tests/.  TEST INFRASTRUCTURE.

A case is fully determined by (workload name, seed): the uniform tapes are SplitMix64 streams,
so the committed golden files only need to hold the seeds and the expected outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from automix_b200 import workloads as W  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def tape(seed: int, n: int) -> np.ndarray:
    return W.splitmix_uniforms_fast(seed, n)


def workload(name: str):
    return getattr(W, name)()


def default_init(wl, seed: int) -> np.ndarray:
    if wl["init"] is not None:
        return np.asarray(wl["init"], dtype=np.float64)
    rng = W.SplitMix64(seed)
    return rng.normals(int(np.sum(wl["dims"])))


def rwm_sweeps(d: int, nsweep2: int) -> int:
    n = max(nsweep2, 10000 * d)
    return n + n // 10


def rwm_tape_len(d: int, nsweep2: int) -> int:
    # per sweep: 1 + max(block: 2*ceil(d/2)+1, single: 3d) uniforms (dof = 0)
    return rwm_sweeps(d, nsweep2) * (1 + 3 * d + 2) + 16


def rj_tape_len(dmax: int, nsweeps: int) -> int:
    # [3d | 2ceil(d/2)+1] + 1 + 1 + 1 + 2ceil(dmax/2) + 1
    return nsweeps * (3 * dmax + 2 * ((dmax + 1) // 2) + 6) + 16


def fit_pipeline(chk, ht, wl, init, seed: int, nsweep2: int = 1000, Lmax: int = 30):
    """Stages 1+2 for every model through checker `chk` (oracle or reference).
    Returns the flat proposal mixture plus the per-model stage outputs."""
    ptr = ht.select(wl["target"])
    dims = wl["dims"]
    wt, mean, tri, sig, ncomp, stages = [], [], [], [], [], []
    off = 0
    for k, d in enumerate(dims):
        d = int(d)
        chk.tape(tape(seed * 1000 + 2 * k, rwm_tape_len(d, nsweep2)))
        r = chk.rwm_within_model(k, d, nsweep2, ptr, init[off:off + d])
        assert not chk.tape_overrun()
        off += d
        chk.tape(tape(seed * 1000 + 2 * k + 1, 4096))
        e = chk.fit_mixture(r["samples"], Lmax=Lmax, maxit=5000)
        ncomp.append(e["L"])
        wt.append(e["lam"])
        mean.append(e["mu"].ravel())
        tri.append(e["B"].ravel())
        sig.append(r["sig"])
        stages.append(dict(rwm=r, em=e))
    mix = dict(dims=np.asarray(dims, np.int32), ncomp=np.array(ncomp, np.int32), wt=np.concatenate(wt),
               mean=np.concatenate(mean), tri=np.concatenate(tri), sig=np.concatenate(sig))
    return mix, stages


def load_golden(name: str):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    return dict(np.load(path, allow_pickle=False))


def sokal_series(seed: int, n: int, stay: float, nmodels: int) -> np.ndarray:
    """A model-index series with persistence ``stay`` (the kind runStats.xr holds, automix.c:122-124):
    the chain keeps its model with probability ``stay``, else draws one uniformly."""
    u = tape(seed, 2 * n)
    x = np.zeros(n)
    k = 0
    for i in range(n):
        if u[2 * i] > stay:
            k = int(u[2 * i + 1] * nmodels)
        x[i] = k
    return x


def ar1_series(seed: int, n: int, phi: float) -> np.ndarray:
    """A real-valued AR(1) series (sokal() takes doubles; the report writer only feeds it model indices)."""
    u = tape(seed, 2 * n)
    z = np.sqrt(-2.0 * np.log(1.0 - u[0::2])) * np.cos(2 * np.pi * u[1::2])
    x = np.zeros(n)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + z[i]
    return x


SOKAL_CASES = [  # (name, kind, seed, n, parameter, nmodels)
    ("k16", "k", 11, 16, 0.3, 2), ("k64", "k", 12, 64, 0.5, 3), ("k1024", "k", 13, 1024, 0.9, 3),
    ("k4096", "k", 14, 4096, 0.97, 6), ("k32768", "k", 15, 32768, 0.99, 2), ("k32768b", "k", 16, 32768, 0.2, 6),
    ("ar256", "ar", 17, 256, 0.8, 0), ("ar8192", "ar", 18, 8192, 0.95, 0), ("ar4", "ar", 19, 4, 0.1, 0),
    ("const64", "k", 20, 64, 1.0, 2),
]


def sokal_case(c):
    name, kind, seed, n, par, nm = c
    return sokal_series(seed, n, par, nm) if kind == "k" else ar1_series(seed, n, par)
