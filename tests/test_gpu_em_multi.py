"""Sample-sharded EM over several GPUs of one box (amx_em_fit_multi): the shards exchange only the per-pass
partial sufficient statistics through NVLink peer memory inside the persistent kernel.  The discrete trace
must equal the single-GPU fit's exactly (every GPU takes the same annihilation / convergence branches) and
the continuous results agree to reduction-order rounding."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def _data(n, d, seed):
    from automix_b200 import workloads as W

    x, _ = W.c5_em_samples(n=n, d=d, G=4, seed=seed)
    return x


def test_single_device_through_the_multi_entry_point(amx):
    x = _data(30000, 5, 3)
    idx = np.arange(12, dtype=np.int32) * 997
    a = amx.em_fit(x, idx, Lmax=12, maxit=8, want_state=True)
    b = amx.em_fit(x, idx, Lmax=12, maxit=8, want_state=True, devices=[0])
    assert np.array_equal(a["trace_L"], b["trace_L"]) and np.array_equal(a["trace_loglik"], b["trace_loglik"])
    assert np.array_equal(a["cur_w"], b["cur_w"])


@pytest.mark.parametrize("n,d,L,maxit", [(40000, 3, 10, 12), (100003, 10, 30, 6), (60000, 20, 12, 4)])
def test_sharded_fit_equals_single_gpu_fit(amx, n, d, L, maxit):
    ndev = amx.device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    x = _data(n, d, 11)
    idx = (np.arange(L, dtype=np.int64) * (n // L) + 5).astype(np.int32)
    one = amx.em_fit(x, idx, Lmax=L, maxit=maxit, want_state=True)
    for g in sorted({2, min(ndev, 4), ndev}):
        many = amx.em_fit(x, idx, Lmax=L, maxit=maxit, want_state=True, devices=list(range(g)))
        assert many["status"] == 0
        assert np.array_equal(one["trace_L"], many["trace_L"]), "component-count trace differs"
        assert np.array_equal(one["trace_ann"], many["trace_ann"]), "annihilation trace differs"
        assert _rel(many["trace_loglik"], one["trace_loglik"]) < 1e-11
        assert _rel(many["cur_mu"], one["cur_mu"]) < 1e-10 and _rel(many["cur_B"], one["cur_B"]) < 1e-10
        assert np.max(np.abs(many["cur_w"] - one["cur_w"])) < 1e-10
        assert many["comp_steps"] == one["comp_steps"]


@pytest.mark.parametrize("ranks", [2, 3, 8])
@pytest.mark.parametrize("n,d,L,maxit", [(40000, 3, 10, 12), (100003, 10, 30, 6), (52000, 12, 12, 4)])
def test_emulated_ranks_on_one_gpu_equal_the_single_fit(amx, ranks, n, d, L, maxit):
    """The sharded exchange without a second GPU: naming one GPU several times makes the library run the ranks as
    slices of one cooperative grid on it -- same kernel, same per-rank reduction, same rows posted to every rank and
    summed in rank order, same per-rank tile loops; only NVLink is missing.  Runs on the driver's single-GPU box."""
    x = _data(n, d, 11)
    idx = (np.arange(L, dtype=np.int64) * (n // L) + 5).astype(np.int32)
    one = amx.em_fit(x, idx, Lmax=L, maxit=maxit, want_state=True)
    many = amx.em_fit(x, idx, Lmax=L, maxit=maxit, want_state=True, devices=[0] * ranks)
    assert many["status"] == 0
    assert np.array_equal(one["trace_L"], many["trace_L"]), "component-count trace differs"
    assert np.array_equal(one["trace_ann"], many["trace_ann"]), "annihilation trace differs"
    assert _rel(many["trace_loglik"], one["trace_loglik"]) < 1e-11
    assert _rel(many["cur_mu"], one["cur_mu"]) < 1e-10 and _rel(many["cur_B"], one["cur_B"]) < 1e-10
    assert np.max(np.abs(many["cur_w"] - one["cur_w"])) < 1e-10
    assert many["comp_steps"] == one["comp_steps"]
    again = amx.em_fit(x, idx, Lmax=L, maxit=maxit, want_state=True, devices=[0] * ranks)
    assert np.array_equal(again["cur_w"], many["cur_w"]), "a sharded fit is bitwise reproducible"


def test_emulated_ranks_against_oracle(amx, orc):
    g = cases.load_golden("em3d")
    x = g["x"]
    orc.tape(cases.tape(77, 4096))
    o = orc.fit_mixture(x, Lmax=12, maxit=3, want_state=True)
    r = amx.em_fit(x, o["init_idx"], Lmax=12, maxit=3, want_state=True, devices=[0, 0, 0, 0])
    assert np.array_equal(r["trace_L"], o["trace_L"]) and np.array_equal(r["trace_ann"], o["trace_ann"])
    assert _rel(r["trace_loglik"], o["trace_loglik"]) < 1e-12
    assert _rel(r["cur_mu"], o["cur_mu"]) < 1e-12 and _rel(r["cur_B"], o["cur_B"]) < 1e-12
    assert np.max(np.abs(r["cur_w"] - o["cur_w"])) < 1e-12


def test_sharded_fit_against_oracle(amx, orc):
    if amx.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    g = cases.load_golden("em3d")
    x = g["x"]
    orc.tape(cases.tape(77, 4096))
    o = orc.fit_mixture(x, Lmax=12, maxit=3, want_state=True)
    r = amx.em_fit(x, o["init_idx"], Lmax=12, maxit=3, want_state=True, devices=[0, 1])
    assert np.array_equal(r["trace_L"], o["trace_L"]) and np.array_equal(r["trace_ann"], o["trace_ann"])
    assert _rel(r["trace_loglik"], o["trace_loglik"]) < 1e-12
    assert _rel(r["cur_mu"], o["cur_mu"]) < 1e-12 and _rel(r["cur_B"], o["cur_B"]) < 1e-12
    assert np.max(np.abs(r["cur_w"] - o["cur_w"])) < 1e-12
