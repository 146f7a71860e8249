"""Workload definitions for the BASELINE.json configurations.

Everything here is plain numpy on flat arrays (include/amx_layout.h); nothing computes on
the hot path.  Random inputs come from a documented SplitMix64 stream so that the host, the
oracle and the device see bit-identical inputs for a given seed.

Target spec dicts:
  {"kind": "gaussmix", dims, ncomp, modw, wt, mean, tri, flags}
  {"kind": "quad", dims, center, scale, lo, hi}
  {"kind": "coalmine", dims}
  {"kind": "mixnorm", dims, ncomp, y, prior}
Proposal (mixture) dicts: {dims, ncomp, wt, mean, tri, sig}.
"""
from __future__ import annotations

import math

import numpy as np

MASK64 = (1 << 64) - 1


class SplitMix64:
    """Steele/Lea/Flood SplitMix64; uniforms are (x >> 11 + 0.5) * 2^-53 in (0,1)."""

    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def uniform(self) -> float:
        return ((self.next_u64() >> 11) + 0.5) * (1.0 / 9007199254740992.0)

    def uniforms(self, n: int) -> np.ndarray:
        return np.array([self.uniform() for _ in range(n)], dtype=np.float64)

    def normal(self) -> float:
        u1, u2 = self.uniform(), self.uniform()
        return math.sqrt(-2.0 * math.log(u1)) * math.sin(2.0 * math.pi * u2)

    def normals(self, n: int) -> np.ndarray:
        return np.array([self.normal() for _ in range(n)], dtype=np.float64)


def splitmix_uniforms_fast(seed: int, n: int) -> np.ndarray:
    """Vectorised SplitMix64 uniforms (same stream as SplitMix64.uniforms)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & MASK64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def tri_len(d: int) -> int:
    return d * (d + 1) // 2


def pack_lower(M) -> np.ndarray:
    M = np.asarray(M, dtype=np.float64)
    d = M.shape[0]
    return np.array([M[i, j] for i in range(d) for j in range(i + 1)], dtype=np.float64)


def _flat(dims, ncomp, wt, mean, tri):
    return dict(dims=np.array(dims, np.int32), ncomp=np.array(ncomp, np.int32),
                wt=np.concatenate([np.ravel(w) for w in wt]).astype(np.float64),
                mean=np.concatenate([np.ravel(m) for m in mean]).astype(np.float64),
                tri=np.concatenate([np.ravel(t) for t in tri]).astype(np.float64))


# --------------------------------------------------------------------------------------
# C1: README 1-D Normal (README.md:50-73; tests/test_automix.c:257-265)
# --------------------------------------------------------------------------------------
def c1_normal():
    target = dict(kind="quad", dims=np.array([1], np.int32), center=np.array([0.5]),
                  scale=np.array([1.0]), lo=None, hi=None)
    return dict(name="c1_normal", target=target, init=np.array([0.5]), dims=np.array([1], np.int32))


def truncnormal():
    """tests/test_automix.c:242-255."""
    target = dict(kind="quad", dims=np.array([1], np.int32), center=np.array([1.0]),
                  scale=np.array([1.0]), lo=np.array([0.0]), hi=np.array([10.0]))
    return dict(name="truncnormal", target=target, init=np.array([1.0]), dims=np.array([1], np.int32))


# --------------------------------------------------------------------------------------
# C2: toy1 (src/user_examples/usertoy1.c:34-104): 2 models, d=1,2
# --------------------------------------------------------------------------------------
def toy1():
    dims = [1, 2]
    ncomp = [2, 3]
    wt = [[0.2, 0.8], [1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0]]
    mean = [[[-3.0], [2.0]], [[0.0, 3.0], [-4.0, 1.0], [4.0, 1.0]]]
    tri = [[[2.0], [1.0]],
           [[2.0, 0.0, 0.7071068], [1.414214, 1.060660, 0.9354143], [1.414214, -1.060660, 0.9354143]]]
    t = _flat(dims, ncomp, wt, mean, tri)
    t.update(kind="gaussmix", modw=np.array([0.3, 0.7]), flags=0)
    return dict(name="toy1", target=t, dims=t["dims"], init=None, true_probs=np.array([0.3, 0.7]))


# --------------------------------------------------------------------------------------
# toy2 (src/user_examples/usertoy2.c:34-79): 5 models, d=1..5
# --------------------------------------------------------------------------------------
def toy2():
    dims = [1, 2, 3, 4, 5]
    ncomp = [2] * 5
    wt, mean, tri = [], [], []
    for d in dims:
        wt.append([0.3, 0.7])
        mean.append([[5.0] * d, [-5.0] * d])
        tri.append([pack_lower(np.eye(d) * 1.0), pack_lower(np.eye(d) * 2.0)])
    t = _flat(dims, ncomp, wt, mean, tri)
    modw = np.array([0.5, 0.25, 0.125, 0.0625, 0.0625])
    t.update(kind="gaussmix", modw=modw, flags=0)
    return dict(name="toy2", target=t, dims=t["dims"], init=np.zeros(sum(dims)), true_probs=modw)


# --------------------------------------------------------------------------------------
# C3: coal-mining change points (src/user_examples/usercpt.c)
# --------------------------------------------------------------------------------------
def coalmine():
    dims = np.array([2 * k + 3 for k in range(6)], np.int32)
    init = []
    for k in range(6):
        v = np.zeros(2 * k + 3)
        v[: k + 2] = 1.0 / 200.0  # alpha / beta (usercpt.c:35-37)
        for j in range(1, k + 2):
            v[k + 1 + j] = (40907.0 * j) / (k + 2)  # usercpt.c:38-40
        init.append(v)
    return dict(name="coalmine", target=dict(kind="coalmine", dims=dims), dims=dims,
                init=np.concatenate(init),
                thesis_probs=np.array([0.058, 0.250, 0.296, 0.234, 0.118, 0.044]))


# --------------------------------------------------------------------------------------
# C4: finite mixture of normals with an unknown number of components ("enzyme-style" data; not in the
# reference -- SURVEY.md 8d gives the synthetic definition)
# --------------------------------------------------------------------------------------
def c4_mixnorm(nmodels: int = 10, n: int = 245, seed: int = 7):
    """y_i, i = 1..245, from 0.6 N(0.19, 0.08^2) + 0.4 N(1.3, 0.5^2) (SplitMix64 stream `seed`: one uniform picks the
    component, then one Box-Muller normal).  Model k (0-based) is a K = k+1 component normal mixture with
    theta = (a_1..a_{K-1} | m_1..m_K | s_1..s_K): stick-breaking logits (w_j = v_j prod_{i<j}(1 - v_i), v = sigmoid(a)),
    means, log standard deviations; d = 3K - 1 = 2, 5, ..., 29.  Normalised independent priors a ~ N(0, 1.5^2),
    m ~ N(0.7, 1), s ~ N(log 0.3, 1); models a priori equally likely.  Start: equal weights, means spread over the
    data's quantiles, log sd = log(sd(y) / K)."""
    rng = SplitMix64(seed)
    y = np.empty(n)
    for i in range(n):
        u = rng.uniform()
        z = rng.normal()
        y[i] = 0.19 + 0.08 * z if u < 0.6 else 1.3 + 0.5 * z
    ncomp = np.arange(1, nmodels + 1, dtype=np.int32)
    dims = (3 * ncomp - 1).astype(np.int32)
    prior = np.array([1.5, 0.7, 1.0, math.log(0.3), 1.0])
    ys = np.sort(y)
    init = []
    for K in ncomp:
        K = int(K)
        # stick-breaking logits of equal weights: v_j = 1 / (K - j)
        a = [math.log((1.0 / (K - j)) / (1.0 - 1.0 / (K - j))) for j in range(K - 1)]
        m = [float(ys[int((j + 0.5) / K * n)]) for j in range(K)]
        s = [math.log(float(np.std(y)) / K)] * K
        init.append(np.array(a + m + s))
    return dict(name="c4_mixnorm", target=dict(kind="mixnorm", dims=dims, ncomp=ncomp, y=y, prior=prior), dims=dims,
                init=np.concatenate(init), y=y)


# --------------------------------------------------------------------------------------
# C5-RJ: synthetic scaling targets: 10 models, d_k = 2k, 3 Gaussians each
# --------------------------------------------------------------------------------------
def c5_rj(seed: int = 2024, nmodels: int = 10):
    rng = SplitMix64(seed)
    dims = [2 * (k + 1) for k in range(nmodels)]
    scales = [0.7, 1.0, 1.5]
    wt, mean, tri = [], [], []
    for d in dims:
        wt.append([0.5, 0.3, 0.2])
        mm, tt = [], []
        for g in range(3):
            mm.append([-4.0 + 8.0 * rng.uniform() for _ in range(d)])
            M = np.eye(d) * scales[g]
            for i in range(d):
                for j in range(i):
                    M[i, j] = 0.3 * rng.normal()
            tt.append(pack_lower(M))
        mean.append(mm)
        tri.append(tt)
    t = _flat(dims, [3] * nmodels, wt, mean, tri)
    modw = np.array([1.0 / (k + 1) for k in range(nmodels)])
    modw /= modw.sum()
    t.update(kind="gaussmix", modw=modw, flags=1)
    # start every model at the mean of its heaviest component
    init = np.concatenate([np.asarray(mean[k][0]) for k in range(nmodels)])
    return dict(name="c5_rj", target=t, dims=t["dims"], init=init, true_probs=modw)


def ideal_proposal(wl, sig_scale: float = 2.38):
    """Proposal mixture equal to a gaussmix target's own components (the fit an exact EM would
    approach), with RWM scales sig = sig_scale/sqrt(d) * (weighted marginal sd)."""
    t = wl["target"]
    assert t["kind"] == "gaussmix"
    dims, ncomp = t["dims"], t["ncomp"]
    sig = []
    iw = im = it = 0
    for k, d in enumerate(dims):
        L = int(ncomp[k])
        nt = tri_len(int(d))
        w = t["wt"][iw:iw + L]
        mu = t["mean"][im:im + L * d].reshape(L, d)
        var = np.zeros(d)
        gm = (w[:, None] * mu).sum(0) / w.sum()
        for l in range(L):
            T = np.zeros((d, d))
            T[np.tril_indices(d)] = t["tri"][it + l * nt: it + (l + 1) * nt]
            cov = T @ T.T
            var += w[l] / w.sum() * (np.diag(cov) + (mu[l] - gm) ** 2)
        sig.append(sig_scale / math.sqrt(d) * np.sqrt(var))
        iw += L
        im += L * d
        it += L * nt
    return dict(dims=dims.copy(), ncomp=ncomp.copy(), wt=t["wt"].copy(), mean=t["mean"].copy(),
                tri=t["tri"].copy(), sig=np.concatenate(sig))


# --------------------------------------------------------------------------------------
# C5-EM: n samples in d dims from a G-component Gaussian mixture
# --------------------------------------------------------------------------------------
def c5_em_samples(n: int = 1_000_000, d: int = 10, G: int = 6, seed: int = 2025):
    """Returns (x [n,d] row-major float64, truth dict).  Vectorised SplitMix64 stream."""
    rng = SplitMix64(seed)
    means = np.array([[-4.0 + 8.0 * rng.uniform() for _ in range(d)] for _ in range(G)])
    scales = [0.7, 1.0, 1.5]
    Ts = []
    for g in range(G):
        M = np.eye(d) * scales[g % 3]
        for i in range(d):
            for j in range(i):
                M[i, j] = 0.3 * rng.normal()
        Ts.append(M)
    w = np.array([1.0 / (g + 1) for g in range(G)])
    w /= w.sum()
    u = splitmix_uniforms_fast(seed + 1, n * (2 * d + 1)).reshape(n, 2 * d + 1)
    comp = np.searchsorted(np.cumsum(w), u[:, 0]).clip(0, G - 1)
    z = np.sqrt(-2.0 * np.log(u[:, 1:d + 1])) * np.sin(2.0 * np.pi * u[:, d + 1:])
    x = np.empty((n, d))
    for g in range(G):
        m = comp == g
        x[m] = means[g] + z[m] @ Ts[g].T
    return np.ascontiguousarray(x), dict(means=means, tris=Ts, wt=w, comp=comp)


def em_init_indices(n: int, Lmax: int, uniforms) -> tuple[np.ndarray, int]:
    """Distinct start rows exactly as automix.c:682-697; returns (indices, uniforms used)."""
    idx = []
    used = 0
    while len(idx) < Lmax:
        v = int(math.floor(n * float(uniforms[used])))
        used += 1
        if v not in idx:
            idx.append(v)
    return np.array(idx, np.int32), used
