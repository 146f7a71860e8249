"""Device log-posterior plug-ins against the host callbacks (oracle/host_targets.c)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["toy1", "toy2", "c5_rj", "c1_normal", "truncnormal", "coalmine"])
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
