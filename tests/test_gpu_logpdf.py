"""K4 (batched mixture log-density) against the oracle: 1e-12 relative, fp64."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _mixture(rng, d, L):
    wt = rng.random(L) + 0.1
    wt /= wt.sum()
    mean = rng.normal(size=(L, d)) * 3
    tri = []
    for _ in range(L):
        A = rng.normal(size=(d, d))
        tri.append(np.linalg.cholesky(A @ A.T + d * np.eye(d))[np.tril_indices(d)])
    return wt, mean, np.array(tri)


@pytest.mark.parametrize("d,L,n", [(1, 1, 1), (1, 3, 100), (2, 5, 1000), (5, 30, 777), (10, 30, 4096), (13, 7, 333), (20, 30, 2048), (32, 32, 65)])
def test_against_oracle(amx, orc, d, L, n):
    rng = np.random.default_rng(d * 100 + L)
    wt, mean, tri = _mixture(rng, d, L)
    x = rng.normal(size=(n, d)) * 4
    comp, mix = amx.mix_logpdf(wt, mean.ravel(), tri.ravel(), x)
    oc, om = orc.mix_logpdf(wt, mean.ravel(), tri.ravel(), x)
    assert np.max(np.abs(comp - oc) / np.maximum(1.0, np.abs(oc))) < RTOL
    ok = np.isfinite(om)
    assert np.array_equal(np.isfinite(mix), ok)
    assert np.max(np.abs(mix[ok] - om[ok]) / np.maximum(1.0, np.abs(om[ok]))) < RTOL


def test_against_reference_golden(amx):
    g = cases.load_golden("helpers")
    comp, _ = amx.mix_logpdf([1.0], g["ln_mu"], g["chol_out"], g["ln_x"])
    assert np.max(np.abs(comp[:, 0] - g["ln_out"]) / np.abs(g["ln_out"])) < RTOL


def test_linearity_of_weights_at_full_size(amx):
    """Size-independent property at 2^20 points: log-mixture density with weights w equals
    logsumexp(log w + component log-densities); checked against the component output."""
    rng = np.random.default_rng(3)
    d, L, n = 10, 30, 1 << 20
    wt, mean, tri = _mixture(rng, d, L)
    x = rng.normal(size=(n, d)) * 3
    comp, mix = amx.mix_logpdf(wt, mean.ravel(), tri.ravel(), x)
    a = comp + np.log(wt)
    m = a.max(1)
    lse = m + np.log(np.exp(a - m[:, None]).sum(1))
    ok = np.isfinite(mix) & (m > -700)
    assert ok.mean() > 0.5
    assert np.max(np.abs(mix[ok] - lse[ok]) / np.maximum(1.0, np.abs(lse[ok]))) < 1e-11


def test_rejects_bad_shapes(amx):
    with pytest.raises(amx.AmxError):
        amx.mix_logpdf(np.ones(40) / 40, np.zeros(40), np.ones(40), np.zeros((3, 1)))
