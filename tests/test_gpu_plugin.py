"""A USER-SUPPLIED __device__ log-posterior (the device variant of the reference's `double f(int, double*)` contract,
automix.h:46): tests/plugins/toy1_user.cuh is written against the SDK headers the way a user would, built into a shared
object by `python -m automix_b200.plugin build` (here by __graft_entry__.build(), so that it travels to the GPU box),
loaded with amx_target_plugin, and must behave exactly like the built-in Gaussian-mixture family on the same
parameters: evaluation, stage-1 chains and reversible-jump populations all bit for bit."""
import os

import numpy as np
import pytest

import cases
from automix_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "plugins", "toy1_user.cuh")


def _plugin():
    from automix_b200 import plugin

    return plugin.build(SRC)


def test_plugin_builds_and_exports_its_entry_point():
    """CPU check: nvcc cross-compiles the user's source against the library's kernel templates."""
    import subprocess

    so = _plugin()
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True, check=True).stdout
    assert " T amx_plugin_entry" in out


@pytest.mark.gpu
def test_user_plugin_equals_the_builtin_family_bit_for_bit(amx):
    so = _plugin()
    wl = W.toy1()
    spec = wl["target"]
    Tb = amx.Target(spec)
    Tu = amx.Target(dict(kind="plugin", so=so, dims=spec["dims"], blob=amx.family_blob(spec), flags=0))
    rng = np.random.default_rng(3)
    k = rng.integers(0, 2, size=5000).astype(np.int32)
    x = rng.normal(size=(5000, 2)) * 4
    assert np.array_equal(Tu.eval(k, x), Tb.eval(k, x))
    # stage 1: three adaptive chains of model 1 (d = 2), Philox streams
    init = cases.default_init(wl, 8)
    a, b = amx.rwm_adapt(Tu, 1, 1000, 3, init[1:3], seed=5), amx.rwm_adapt(Tb, 1, 1000, 3, init[1:3], seed=5)
    assert np.array_equal(a["samples"], b["samples"]) and np.array_equal(a["sig"], b["sig"])
    # stage 3: a population on the mixtures the reference fitted
    g = cases.load_golden("toy1")
    mix = {q[4:]: g[q] for q in g if q.startswith("mix_")}
    out = []
    for T in (Tu, Tb):
        pop = amx.RjPopulation(amx.Proposal(mix), T, 8192, g["init"], seed=9, n_trace=4)
        pop.init_chains()
        pop.sweeps(50, burning=True)
        pop.sweeps(150)
        vis, st = pop.collect()
        out.append((vis.copy(), pop.get_state()["theta"].copy(), pop.trace()["lp"].copy(), st["acc_jump"], st["draws"]))
        pop.close()
    for q in range(5):
        assert np.array_equal(out[0][q], out[1][q]), q
    assert out[0][0].sum() == 8192 * 200 and abs(out[0][0][0] / out[0][0].sum() - 0.3) < 0.02
