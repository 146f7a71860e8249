"""CPU baseline: the reference's own implementation of the hot path, timed on the host cores.
TEST / BENCH INFRASTRUCTURE (the cpu_baseline and `--impl reference` legs of bench.py).

The reference is single-threaded with file-scope RNG state (automix.c:1297-1298), so host
throughput is P independent processes, one per core, each with its own seed (BASELINE.md
section 3).  kind = "reference" when oracle/_ref (the unmodified reference compiled from its own
sources) is present, else "port" (our CPU restatement, oracle/amx_oracle.c).

No torch import here: the workers must start fast and must never touch CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _rj_worker(args):
    """One core: the reference's burn_samples + rjmcmc_samples on a given proposal mixture."""
    spec, mix, init, nburn, nsweeps, seed, kind = args
    import pyoracle as po

    ht = po.HostTargets()
    ptr = ht.select(spec)
    dims, ncomp = po.i32(mix["dims"]), po.i32(mix["ncomp"])
    wt, mean, tri, sig = po.f64(mix["wt"]), po.f64(mix["mean"]), po.f64(mix["tri"]), po.f64(mix["sig"])
    init = po.f64(init)
    nm = len(dims)
    if kind == "reference":
        lib = C.CDLL(os.path.join(HERE, "_ref", "libref_bench.so"))
        vis = np.zeros(nm, np.int64)
        rep, wall = C.c_double(0), C.c_double(0)
        cnt = np.zeros(6, np.uint64)
        lib.refbench_rj.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp, C.c_void_p, C.c_int, C.c_int,
                                    C.c_ulong, C.POINTER(C.c_long), _dp, _dp, C.POINTER(C.c_ulong)]
        rc = lib.refbench_rj(nm, _i(dims), _i(ncomp), _d(wt), _d(mean), _d(tri), _d(sig), _d(init), ptr, nburn,
                             nsweeps, seed, vis.ctypes.data_as(C.POINTER(C.c_long)), C.byref(rep), C.byref(wall),
                             cnt.ctypes.data_as(C.POINTER(C.c_ulong)))
        assert rc == 0
        return dict(sweeps=nsweeps, secs=wall.value, secs_reported=rep.value, visits=vis.tolist())
    orc = po.Checker("orc")
    sd = np.array([seed], dtype=np.uint64)
    orc.lib.orc_sdrni.argtypes = [C.POINTER(C.c_ulong)]
    orc.lib.orc_tape_set(None, 0)
    orc.lib.orc_sdrni(sd.ctypes.data_as(C.POINTER(C.c_ulong)))
    s0 = orc.chain_init(dims, init, ptr)
    a = orc.rj_sweeps(mix, ptr, s0, nburn, burning=True, trace=False)
    t0 = time.perf_counter()
    b = orc.rj_sweeps(mix, ptr, a["state"], nsweeps, trace=False)
    dt = time.perf_counter() - t0
    return dict(sweeps=nsweeps, secs=dt, secs_reported=dt, visits=b["visits"].tolist())


def _em_worker(args):
    x, Lmax, maxit, seed, kind = args
    import pyoracle as po

    x = po.f64(x)
    n, d = x.shape
    if kind == "reference":
        lib = C.CDLL(os.path.join(HERE, "_ref", "libref_bench.so"))
        it, L = C.c_int(0), C.c_int(0)
        wall = C.c_double(0)
        lib.refbench_em.argtypes = [C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_ulong, _ip, _ip, _dp]
        lib.refbench_em(d, n, _d(x), Lmax, maxit, seed, C.byref(it), C.byref(L), C.byref(wall))
        return dict(n=n, iters=it.value, L=L.value, secs=wall.value)
    orc = po.Checker("orc")
    sd = np.array([seed], dtype=np.uint64)
    orc.lib.orc_sdrni.argtypes = [C.POINTER(C.c_ulong)]
    orc.lib.orc_tape_set(None, 0)
    orc.lib.orc_sdrni(sd.ctypes.data_as(C.POINTER(C.c_ulong)))
    t0 = time.perf_counter()
    e = orc.fit_mixture(x, Lmax=Lmax, maxit=maxit)
    return dict(n=n, iters=e["iters"], L=e["L"], secs=time.perf_counter() - t0)


def kind_available() -> str:
    return "reference" if os.path.exists(os.path.join(HERE, "_ref", "libref_bench.so")) else "port"


def _pool_map(fn, jobs, timeout_s: float = 900.0):
    """Run one plain subprocess per job (never fork or spawn from a process that may hold a CUDA
    context; never hang: every worker has a hard timeout)."""
    tmp = tempfile.mkdtemp(prefix="amx_cpu_bench_")
    procs = []
    t0 = time.perf_counter()
    for i, job in enumerate(jobs):
        jf, rf = os.path.join(tmp, f"job{i}.pkl"), os.path.join(tmp, f"res{i}.pkl")
        with open(jf, "wb") as f:
            pickle.dump((fn.__name__, job), f)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
        procs.append((subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", jf, rf], env=env), rf))
    out = []
    for pr, rf in procs:
        try:
            pr.wait(timeout=max(1.0, timeout_s - (time.perf_counter() - t0)))
        except subprocess.TimeoutExpired:
            pr.kill()
            raise RuntimeError("CPU baseline worker timed out")
        if pr.returncode != 0:
            raise RuntimeError(f"CPU baseline worker failed with code {pr.returncode}")
        with open(rf, "rb") as f:
            out.append(pickle.load(f))
    return out, time.perf_counter() - t0


def rj_baseline(spec, mix, init, nburn: int, nsweeps: int, cores: int | None = None):
    """chain-sweeps/s of the reference on `cores` host cores (one independent process per core)."""
    import pyoracle as po

    po.build(ref=True)
    kind = kind_available()
    P = cores or usable_cores()
    jobs = [(spec, mix, init, nburn, nsweeps, 1000 + 17 * p, kind) for p in range(P)]
    res, wall = _pool_map(_rj_worker, jobs)
    slowest = max(r["secs"] for r in res)
    total = sum(r["sweeps"] for r in res)
    vis = np.sum([r["visits"] for r in res], axis=0)
    return dict(value=total / slowest, unit="chain-sweeps/s", cores=P, kind=kind,
                sample=f"{P} processes x {nsweeps} rjmcmc_samples sweeps (after {nburn} burn-in) of the reference on the "
                       f"same targets and the same fitted proposal; slowest process {slowest:.2f} s, pool wall {wall:.1f} s",
                per_core=total / slowest / P, model_probs=(vis / vis.sum()).round(4).tolist())


def em_baseline(x, Lmax: int, maxit: int, cores: int | None = None):
    """EM-fit samples/s (n * outer iterations / s) of the reference's fit_mixture_from_samples."""
    import pyoracle as po

    po.build(ref=True)
    kind = kind_available()
    P = cores or usable_cores()
    jobs = [(x, Lmax, maxit, 2000 + 13 * p, kind) for p in range(P)]
    res, wall = _pool_map(_em_worker, jobs)
    slowest = max(r["secs"] for r in res)
    total = sum(r["n"] * r["iters"] for r in res)
    return dict(value=total / slowest, unit="EM-fit samples/s", cores=P, kind=kind,
                sample=f"{P} processes x fit_mixture_from_samples on the first {len(x)} samples of the workload, "
                       f"Lmax={Lmax}, {res[0]['iters']} outer iteration(s); slowest process {slowest:.2f} s, pool wall {wall:.1f} s",
                per_core=total / slowest / P)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--worker":
        with open(sys.argv[2], "rb") as f:
            name, job = pickle.load(f)
        res = {"_rj_worker": _rj_worker, "_em_worker": _em_worker}[name](job)
        with open(sys.argv[3], "wb") as f:
            pickle.dump(res, f)
