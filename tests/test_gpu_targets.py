"""Device log-posterior plug-ins against the host callbacks (oracle/host_targets.c)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["toy1", "toy2", "c5_rj", "c1_normal", "truncnormal", "coalmine", "c4_mixnorm"])
def test_plugin_matches_host_callback(amx, ht, name):
    wl = cases.workload(name)
    spec = wl["target"]
    ht.select(spec)
    T = amx.Target(spec)
    rng = np.random.default_rng(4)
    dims = np.asarray(wl["dims"])
    nm, dmax = len(dims), int(dims.max())
    n = 4000
    k = rng.integers(0, nm, size=n).astype(np.int32)
    if name == "coalmine":
        x = np.zeros((n, dmax))
        offs = np.concatenate([[0], np.cumsum(dims)])
        for i in range(n):
            x0 = wl["init"][offs[k[i]]:offs[k[i] + 1]]
            x[i, : dims[k[i]]] = x0 * (1 + 0.3 * rng.normal(size=dims[k[i]]))
    elif name == "c4_mixnorm":  # around the start values and far out (logits +-40, log sds down to -8: softplus tails)
        x = np.zeros((n, dmax))
        offs = np.concatenate([[0], np.cumsum(dims)])
        for i in range(n):
            x0 = wl["init"][offs[k[i]]:offs[k[i] + 1]]
            x[i, : dims[k[i]]] = x0 + rng.normal(size=dims[k[i]]) * (0.3 if i % 4 else 12.0)
    elif name == "c5_rj":
        x = rng.normal(size=(n, dmax)) * 2.5
    else:
        x = rng.normal(size=(n, dmax)) * 4
    got = T.eval(k, x)
    want = np.array([ht.logpost(int(k[i]), x[i, : dims[k[i]]].copy()) for i in range(n)])
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    err = np.abs(got[fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))
    assert err.max() < 1e-12, err.max()


def test_coalmine_fast_path_and_walk_agree_with_the_reference_walk(amx, ht):
    """The device plug-in replaces the 191-step data walk by binary searches when that is exactly equivalent,
    and keeps the walk otherwise (two change points inside one data gap, change points beyond the last datum,
    before the first one).  The host callback always walks, as usercpt.c does."""
    wl = cases.workload("coalmine")
    ht.select(wl["target"])
    T = amx.Target(wl["target"])
    rng = np.random.default_rng(11)
    y = np.array([74.0, 231, 354, 356, 480, 492, 40623])
    ks, xs = [], []
    for trial in range(3000):
        k = int(rng.integers(0, 6))
        ns = k + 1
        h = rng.random(ns + 1) * 0.02 + 1e-4
        mode = trial % 5
        if mode == 0:    # generic positions
            s = np.sort(rng.random(ns) * 40907.0)
        elif mode == 1:  # several change points inside one data gap (1145 .. 1971, 36673 .. 39039)
            lo, hi = ((1146.0, 1970.0), (36674.0, 39038.0))[trial % 2]
            s = np.sort(lo + rng.random(ns) * (hi - lo))
        elif mode == 2:  # beyond the last datum
            s = np.sort(40624.0 + rng.random(ns) * 280.0)
        elif mode == 3:  # before the first datum / on top of data values
            s = np.sort(np.concatenate([rng.random(1) * 74.0, rng.choice(y, ns - 1) + rng.integers(0, 2, ns - 1)]))[:ns] if ns > 1 else rng.random(1) * 74.0
            s = np.sort(s + np.arange(ns) * 1e-3)
        else:            # clustered
            c = rng.random() * 40000.0
            s = np.sort(c + rng.random(ns) * 300.0)
        x = np.zeros(13)
        x[: ns + 1] = h
        x[ns + 1: 2 * ns + 1] = s
        ks.append(k)
        xs.append(x)
    ks = np.array(ks, np.int32)
    xs = np.array(xs)
    got = T.eval(ks, xs)
    dims = wl["dims"]
    want = np.array([ht.logpost(int(k), x[: dims[k]].copy()) for k, x in zip(ks, xs)])
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert err.max() < 1e-12, (err.max(), int(err.argmax()))
