"""K3 (reversible-jump sweeps) against the oracle, the reference's golden traces and, where
oracle/_ref travelled to this box, the reference itself -- on identical injected uniforms.

Bar (BASELINE.json north_star): accept decisions and model visits bit-exact, continuous state
within 1e-12 relative."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-12   # per-step bar: states re-synchronised with the checker at least every RESYNC sweeps
DRIFT_RTOL = 1e-9  # free-running trajectories accumulate rounding (different libm, FMA); decisions stay bit-exact
RESYNC = 40


def _golden_mix(g):
    return {k[4:]: g[k] for k in g if k.startswith("mix_")}


def _close(a, b, what, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() < rtol, (what, err.max(), int(err.argmax()))


@pytest.mark.parametrize("name", ["toy1", "toy2"])
def test_trace_chain_against_reference_golden(amx, name):
    g = cases.load_golden(name)
    wl = cases.workload(name)
    seed = int(g["seed"][0])
    mix = _golden_mix(g)
    dmax = int(max(wl["dims"]))
    n1, n2 = (int(v) for v in g["rj_nsweeps"])
    tape = cases.tape(seed * 1000 + 999, cases.rj_tape_len(dmax, n1 + n2) + 8)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, 3, g["init"], n_trace=3)
    pop.set_tape(np.tile(tape, (3, 1)))  # three chains fed the same tape must stay identical
    pop.init_chains()
    s0 = pop.get_state()
    assert np.all(s0["k"] == int(g["rj_init_k"][0]))
    _close(s0["lp"], np.full(3, g["rj_init_lp"][0]), "init lp")
    for tag, n, burning in (("burn", n1, True), ("run", n2, False)):
        pop.sweeps(n, burning=burning)
        vis, st = pop.collect(reset=True)
        tr = pop.trace()
        for c in range(3):
            assert np.array_equal(tr["k"][c], g[f"rj_{tag}_k"]), (tag, "model sequence")
            if tag == "burn":  # the first sweeps after the common start: per-step bar
                _close(tr["lp"][c][:RESYNC], g[f"rj_{tag}_lp"][:RESYNC], tag + " lp (first sweeps)")
                _close(tr["theta"][c][:RESYNC], g[f"rj_{tag}_theta"][:RESYNC], tag + " theta (first sweeps)")
            _close(tr["lp"][c], g[f"rj_{tag}_lp"], tag + " lp", DRIFT_RTOL)
            _close(tr["theta"][c], g[f"rj_{tag}_theta"], tag + " theta", DRIFT_RTOL)
            _close(tr["pk"][c], g[f"rj_{tag}_pk"], tag + " pk", DRIFT_RTOL)
        cnt = g[f"rj_{tag}_counters"]
        got = [st[f] for f in ("acc_block", "try_block", "acc_single", "try_single", "acc_jump", "try_jump")]
        assert got == [3 * int(v) for v in cnt], (tag, "accept counters")
        assert np.array_equal(vis, 3 * g[f"rj_{tag}_visits"].astype(np.uint64))
    assert st["draws"] > 0


def _fitted_proposal(amx, wl):
    """Stages 1-2 on the device, as the pipeline runs them: the proposal a real run jumps with (ill-scaled for the
    coal-mining model: rates ~1e-3 next to change points ~1e4, components with log-densities far in the tail)."""
    T = amx.Target(wl["target"])
    dims = np.asarray(wl["dims"])
    ncomp, wt, mean, tri, sig = [], [], [], [], []
    off = 0
    for k, d in enumerate(dims):
        d = int(d)
        r = amx.rwm_adapt(T, k, 1000, 1, wl["init"][off:off + d], seed=11 + k)
        off += d
        x = r["samples"][0]
        idx, _ = amx.em_draw_init(len(x), 30, cases.tape(40 + k, 4096))
        e = amx.em_fit(x, idx, Lmax=30, maxit=300)
        assert 1 <= e["L"] <= 30
        ncomp.append(e["L"])
        wt.append(e["lam"])
        mean.append(e["mu"].ravel())
        tri.append(e["B"].ravel())
        sig.append(r["sig"][0])
    return dict(dims=dims.astype(np.int32), ncomp=np.array(ncomp, np.int32), wt=np.concatenate(wt),
                mean=np.concatenate(mean), tri=np.concatenate(tri), sig=np.concatenate(sig))


def _handmade_proposal(name, dims, init):
    """A simple hand-made proposal: two components per model around the start point."""
    wt, mean, tri, sig, ncomp = [], [], [], [], []
    off = 0
    for d in dims:
        x0 = init[off:off + d]
        off += d
        sc = np.maximum(np.abs(x0) * 0.15, 0.05) if name == "coalmine" else (np.full(d, 0.25) if name == "c4_mixnorm" else np.ones(d))
        L = 2
        ncomp.append(L)
        wt.append([0.6, 0.4])
        mean.append(np.concatenate([x0, x0 + 0.5 * sc]))
        tri.append(np.concatenate([np.diag(sc)[np.tril_indices(d)], np.diag(1.5 * sc)[np.tril_indices(d)]]))
        sig.append(0.5 * sc)
    return dict(dims=np.asarray(dims).astype(np.int32), ncomp=np.array(ncomp, np.int32), wt=np.concatenate(wt),
                mean=np.concatenate(mean), tri=np.concatenate(tri), sig=np.concatenate(sig))


@pytest.mark.parametrize("name,nchains,nsweeps", [("toy1", 64, 300), ("toy2", 48, 200), ("c5_rj", 40, 120), ("c1_normal", 33, 250), ("coalmine", 24, 150),
                                                  ("coalmine_fitted", 24, 160), ("c4_mixnorm", 20, 80)])
def test_population_against_oracle(amx, orc, ht, name, nchains, nsweeps):
    """Every chain gets its own tape; the oracle replays each chain on the CPU."""
    fitted = name.endswith("_fitted")
    name = name.replace("_fitted", "")
    wl = cases.workload(name)
    spec = wl["target"]
    ptr = ht.select(spec)
    dims = np.asarray(wl["dims"])
    dmax = int(dims.max())
    init = cases.default_init(wl, 5)
    if fitted:
        mix = _fitted_proposal(amx, wl)
    elif spec["kind"] == "gaussmix":
        from automix_b200 import workloads as W

        mix = W.ideal_proposal(wl)
    else:
        mix = _handmade_proposal(name, dims, init)
    tlen = cases.rj_tape_len(dmax, nsweeps) + 8
    tapes = np.stack([cases.tape(1000 + c, tlen) for c in range(nchains)])
    T, P = amx.Target(spec), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, nchains, init, n_trace=nchains)
    pop.set_tape(tapes)
    pop.init_chains()
    # the oracle replays every chain; after each segment of RESYNC sweeps the device population is
    # re-synchronised to the oracle's states, so every compared sweep starts from identical inputs
    states = []
    for c in range(nchains):
        orc.tape(tapes[c])
        states.append(orc.chain_init(dims, init, ptr))
    s0 = pop.get_state()
    assert np.array_equal(s0["k"], [s["k"] for s in states])
    _close(s0["lp"], [s["lp"] for s in states], "init lp")
    used = [1] * nchains
    done = 0
    tot_vis = np.zeros(len(dims), np.int64)
    tot_cnt = np.zeros(6, np.int64)
    while done < nsweeps:
        seg = min(RESYNC, nsweeps - done)
        burning = done < nsweeps // 2
        pop.sweeps(seg, burning=burning)
        vis, st = pop.collect(reset=True)
        tr = pop.trace()
        seg_vis = np.zeros(len(dims), np.int64)
        seg_cnt = np.zeros(6, np.int64)
        seg_draws = 0
        for c in range(nchains):
            orc.tape(tapes[c][used[c]:])
            r = orc.rj_sweeps(mix, ptr, states[c], seg, burning=burning)
            assert not orc.tape_overrun()
            used[c] += orc.tape_used()
            seg_draws += orc.tape_used()
            states[c] = r["state"]
            assert np.array_equal(tr["k"][c], r["k"]), (c, done, "model sequence")
            _close(tr["lp"][c], r["lp"], "lp")
            _close(tr["theta"][c], r["theta"], "theta")
            _close(tr["pk"][c], r["pk"], "pk")
            seg_vis += r["visits"]
            seg_cnt += r["counters"].astype(np.int64)
        assert np.array_equal(vis.astype(np.int64), seg_vis)
        got = [st[f] for f in ("acc_block", "try_block", "acc_single", "try_single", "acc_jump", "try_jump")]
        assert got == list(seg_cnt)
        assert st["draws"] == seg_draws, "uniforms consumed"
        done += seg
        fin = pop.get_state()
        assert fin["sweep_i"] == 1 + done
        assert np.array_equal(fin["nreinit"], [s["nreinit"] for s in states])
        pop.set_state(states, 1 + done)


def test_against_reference_binary_if_present(amx, po, ht):
    if not po.have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    ref = po.Checker("ref")
    g = cases.load_golden("toy1")
    wl = cases.workload("toy1")
    mix = _golden_mix(g)
    ptr = ht.select(wl["target"])
    tape = cases.tape(4242, cases.rj_tape_len(2, 800) + 8)
    ref.tape(tape)
    s0 = ref.chain_init(wl["dims"], g["init"], ptr)
    r = ref.rj_sweeps(mix, ptr, s0, 800)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, 1, g["init"], n_trace=1)
    pop.set_tape(tape[None, :])
    pop.init_chains()
    pop.sweeps(800)
    pop.collect()
    tr = pop.trace()
    assert np.array_equal(tr["k"][0], r["k"])
    _close(tr["lp"][0], r["lp"], "lp")
    _close(tr["theta"][0], r["theta"], "theta")


def test_philox_streams_are_reproducible_and_partition_invariant(amx):
    """Counter-based RNG keyed by (seed, chain id): the histogram for a fixed (seed, chains) must
    not depend on how the launch is split in time, and chains must differ from one another."""
    wl = cases.workload("toy1")
    from automix_b200 import workloads as W

    mix = W.ideal_proposal(wl)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    init = cases.default_init(wl, 9)
    out = []
    for split in ((400,), (150, 250), (1, 399)):
        pop = amx.RjPopulation(P, T, 1000, init, seed=77)
        pop.init_chains()
        for n in split:
            pop.sweeps(n)
        vis, st = pop.collect()
        out.append((vis.copy(), st["acc_jump"], st["draws"], pop.get_state()["theta"].copy()))
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and o[1] == out[0][1] and o[2] == out[0][2]
        assert np.array_equal(o[3], out[0][3])
    assert len(np.unique(out[0][3][:, 0])) > 900
    assert out[0][0].sum() == 1000 * 400


def test_posterior_model_probabilities_toy1(amx):
    """End-to-end statistical check: toy1's true model probabilities are 0.3 / 0.7
    (usertoy1.c:96-100; thesis p.167 reports 0.2997 / 0.7003)."""
    wl = cases.workload("toy1")
    from automix_b200 import workloads as W

    mix = W.ideal_proposal(wl)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, 1 << 16, cases.default_init(wl, 3), seed=2024)
    pop.init_chains()
    pop.sweeps(300, burning=True)
    pop.collect(reset=True)
    pop.sweeps(400)
    vis, st = pop.collect()
    p = vis / vis.sum()
    assert vis.sum() == (1 << 16) * 400
    assert abs(p[0] - 0.3) < 0.004, p
    assert 0.5 < st["acc_jump"] / st["try_jump"] <= 1.0


@pytest.mark.parametrize("mode", ["scalar", "batched"])
def test_host_callback_mode_against_oracle(amx, orc, ht, mode):
    """The reference's scalar callback contract (and the batched variant) through the split
    propose | host evaluates | finish kernels: same tape, same answers as the oracle."""
    wl = cases.workload("toy2")
    ptr = ht.select(wl["target"])
    g = cases.load_golden("toy2")
    mix = _golden_mix(g)
    dims = np.asarray(wl["dims"])
    nchains, nsweeps = 6, 60
    tapes = np.stack([cases.tape(7000 + c, cases.rj_tape_len(5, nsweeps) + 8) for c in range(nchains)])
    T = amx.Target(wl["target"], host_fn=ht.ptr) if mode == "scalar" else amx.Target(wl["target"], host_batched=ht.batched_ptr)
    P = amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, nchains, g["init"], n_trace=nchains)
    pop.set_tape(tapes)
    pop.init_chains()
    pop.sweeps(nsweeps // 2, burning=True)
    pop.collect(reset=True)
    tr1 = pop.trace()
    pop.sweeps(nsweeps - nsweeps // 2)
    vis, st = pop.collect()
    tr2 = pop.trace()
    tot = np.zeros(len(dims), np.int64)
    for c in range(nchains):
        orc.tape(tapes[c])
        s0 = orc.chain_init(dims, g["init"], ptr)
        a = orc.rj_sweeps(mix, ptr, s0, nsweeps // 2, burning=True)
        b = orc.rj_sweeps(mix, ptr, a["state"], nsweeps - nsweeps // 2)
        assert np.array_equal(tr1["k"][c], a["k"]) and np.array_equal(tr2["k"][c], b["k"])
        _close(tr2["lp"][c], b["lp"], "lp")
        _close(tr2["theta"][c], b["theta"], "theta")
        _close(tr2["pk"][c], b["pk"], "pk")
        tot += b["visits"]
    assert np.array_equal(vis.astype(np.int64), tot)


@pytest.mark.parametrize("dof,perm", [(5, 0), (0, 1), (3, 1), (1, 0)])
def test_student_t_and_permutation_modes_against_oracle(amx, orc, ht, dof, perm):
    """The optional modes of the sweep (student_T_dof, doPerm): the gamma rejection sampler consumes a
    data-dependent number of uniforms, so a single wrong draw desynchronises everything after it."""
    wl = cases.workload("toy2")
    ptr = ht.select(wl["target"])
    g = cases.load_golden("toy2")
    mix = _golden_mix(g)
    dims = np.asarray(wl["dims"])
    nchains, nsweeps = 16, 80
    tapes = np.stack([cases.tape(9000 + 31 * c + dof, 4 * cases.rj_tape_len(5, nsweeps)) for c in range(nchains)])
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, nchains, g["init"], n_trace=nchains)
    pop.set_modes(dof=dof, do_perm=perm)
    pop.set_tape(tapes)
    pop.init_chains()
    pop.sweeps(nsweeps)
    vis, st = pop.collect()
    tr = pop.trace()
    draws = 0
    for c in range(nchains):
        orc.tape(tapes[c])
        s0 = orc.chain_init(dims, g["init"], ptr)
        r = orc.rj_sweeps(mix, ptr, s0, nsweeps, do_perm=perm, dof=dof)
        assert not orc.tape_overrun()
        draws += orc.tape_used()
        assert np.array_equal(tr["k"][c], r["k"]), (c, "model sequence")
        _close(tr["lp"][c], r["lp"], "lp", 1e-11)
        _close(tr["theta"][c], r["theta"], "theta", 1e-11)
    assert st["draws"] + nchains == draws  # the oracle's count includes the chain-start uniform


def test_deferred_sync_state_transfers(amx):
    """amx_set_deferred_sync: set_state / get_state only enqueue; after amx_synchronize the (pinned) buffers hold
    what the blocking calls return, and two populations on two streams can be driven from one host thread."""
    import torch

    from automix_b200 import workloads as W

    wl = W.toy1()
    mix = W.ideal_proposal(wl)
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pops = [amx.RjPopulation(P, T, 4096, cases.default_init(wl, 5), seed=5 + q) for q in range(2)]
    for p in pops:
        p.init_chains()
        p.sweeps(20, burning=True)
    ref_states = [p.get_state() for p in pops]
    streams = [torch.cuda.Stream() for _ in pops]
    pins = []
    for st in ref_states:
        pin = {}
        for key in ("theta", "pk", "lp", "k", "nreinit", "pkllim"):
            t = torch.from_numpy(np.ascontiguousarray(st[key])).pin_memory()
            pin[key], pin["_t_" + key] = t.numpy(), t
        pins.append(pin)
    try:
        amx.set_deferred_sync(True)
        for rep in range(3):
            for p, pin, s, st in zip(pops, pins, streams, ref_states):
                amx.set_stream(s.cuda_stream)
                amx.synchronize()
                p.set_state_arrays(pin, st["sweep_i"] + 10 * rep)
                p.sweeps(10)
                p.get_state(out=pin)
        for s in streams:
            amx.set_stream(s.cuda_stream)
            amx.synchronize()
    finally:
        amx.set_deferred_sync(False)
        amx.set_stream(None)
    # the same 30 sweeps with blocking calls from the same start, on fresh populations with the same streams of uniforms
    for q, (st, pin) in enumerate(zip(ref_states, pins)):
        chk = amx.RjPopulation(P, T, 4096, cases.default_init(wl, 5), seed=5 + q)
        chk.init_chains()
        chk.sweeps(20, burning=True)
        assert np.array_equal(chk.get_state()["theta"], st["theta"])
        chk.sweeps(30)
        want = chk.get_state()
        for key in ("theta", "pk", "lp", "k"):
            assert np.array_equal(pin[key], want[key]), key


@pytest.mark.parametrize("name,nchains,pop_pk", [("c5_rj", 3000, False), ("c5_rj", 2500, True), ("toy2", 2000, False),
                                                 ("coalmine", 1500, False), ("c4_mixnorm", 600, False)])
def test_sorted_mode_is_bit_identical(amx, name, nchains, pop_pk):
    """amx_rj_set_sort: chains regrouped by (model, proposed model) before every launch.  Nothing per chain may change:
    states, traces, counters, uniforms consumed and visit counts equal the unsorted kernel's bit for bit, whatever
    the number of sweeps per sort; the per-group visit counts (Monte-Carlo error) still partition the total."""
    from automix_b200 import workloads as W

    wl = cases.workload(name)
    if wl["target"]["kind"] == "gaussmix":
        mix = W.ideal_proposal(wl)
    elif name == "coalmine":
        mix = _fitted_proposal(amx, wl)
    else:
        mix = _handmade_proposal(name, np.asarray(wl["dims"]), cases.default_init(wl, 5))
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    nsw = 47 if name != "c4_mixnorm" else 23
    init = cases.default_init(wl, 5)
    out = []
    for sort in (0, 1, 5):
        pop = amx.RjPopulation(P, T, nchains, init, seed=31, n_trace=5)
        pop.set_sort(sort)
        if pop_pk:
            pop.set_pk_mode(True, 7)
        pop.init_chains()
        pop.sweeps(13, burning=True)
        pop.sweeps(nsw)  # crosses block-move sweeps (every 10th) at different offsets inside a segment
        vis, st = pop.collect()
        p, se, ng = pop.visit_se()
        s = pop.get_state()
        tr = pop.trace()
        out.append((vis, st, s, tr, p, se, ng))
    v0, st0, s0, tr0 = out[0][:4]
    assert v0.sum() == nchains * (13 + nsw)
    for vis, st, s, tr, p, se, ng in out[1:]:
        assert np.array_equal(vis, v0)
        for f in ("acc_block", "try_block", "acc_single", "try_single", "acc_jump", "try_jump", "flops", "draws"):
            assert st[f] == st0[f], f
        for f in ("theta", "pk", "lp", "k", "nreinit", "pkllim"):
            assert np.array_equal(s[f], s0[f]), f
        for f in ("k", "lp", "theta", "pk"):
            assert np.array_equal(tr[f], tr0[f]), ("trace", f)
        assert np.allclose(p, v0 / v0.sum(), atol=1e-15) and ng >= 1 and np.all(np.isfinite(se) | (ng < 2))


def test_sorted_mode_on_tapes_against_oracle(amx, orc, ht):
    """The sort reads a chain's coming model draw by random access into its stream; on injected tapes that is a read
    ahead on the tape.  Sorted chains must still follow the oracle's model sequence exactly."""
    from automix_b200 import workloads as W

    wl = cases.workload("c5_rj")
    mix = W.ideal_proposal(wl)
    nch, nsw = 24, 60
    dmax = int(max(wl["dims"]))
    tapes = np.stack([cases.tape(900 + c, cases.rj_tape_len(dmax, nsw) + 8) for c in range(nch)])
    T, P = amx.Target(wl["target"]), amx.Proposal(mix)
    pop = amx.RjPopulation(P, T, nch, wl["init"], n_trace=nch)
    pop.set_sort(1)
    pop.set_tape(tapes)
    pop.init_chains()
    pop.sweeps(nsw)
    pop.collect()
    tr = pop.trace()
    ptr = ht.select(wl["target"])
    for c in range(nch):
        orc.tape(tapes[c])
        r = orc.rj_sweeps(mix, ptr, orc.chain_init(wl["dims"], wl["init"], ptr), nsw)
        assert np.array_equal(tr["k"][c], r["k"]), c
        _close(tr["lp"][c], r["lp"], "lp", DRIFT_RTOL)
