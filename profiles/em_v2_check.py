"""EM second generation against the first (AMX_EM_V2=0) and timing at the benchmark size.
usage: python profiles/em_v2_check.py [small|bench]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from automix_b200 import _lib as amx, workloads as W

def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0

def fit(x, idx, Lmax, maxit, v2, **env):
    os.environ["AMX_EM_V2"] = "1" if v2 else "0"
    for k, v in env.items():
        os.environ[k] = str(v)
    r = amx.em_fit(x, idx, Lmax=Lmax, maxit=maxit)
    for k in env:
        os.environ.pop(k, None)
    return r

mode = sys.argv[1] if len(sys.argv) > 1 else "small"
if mode == "small":
    for (n, d, G, Lmax, maxit) in ((3000, 2, 3, 8, 40), (20000, 10, 6, 30, 30), (5000, 12, 4, 16, 20), (777, 4, 2, 5, 50), (200000, 10, 6, 30, 3)):
        x, _ = W.c5_em_samples(n=n, d=d, G=G, seed=7 + d)
        idx, _ = amx.em_draw_init(n, Lmax, W.splitmix_uniforms_fast(5, 4096))
        a = fit(x, idx, Lmax, maxit, False)
        for teams in (3,):
            b = fit(x, idx, Lmax, maxit, True, AMX_EM_TEAMS=teams)
            same = np.array_equal(a["trace_L"], b["trace_L"]) and np.array_equal(a["trace_ann"], b["trace_ann"])
            print(f"n={n} d={d} Lmax={Lmax} teams={teams}: iters {a['iters']}/{b['iters']} trace same={same} "
                  f"loglik rel {rel(b['trace_loglik'], a['trace_loglik']):.2e} mu rel {rel(b['mu'], a['mu']) if a['L']==b['L'] else -1:.2e} "
                  f"ms v1 {a['kernel_ms']:.2f} v2 {b['kernel_ms']:.2f} steps {a['comp_steps']}/{b['comp_steps']}", flush=True)
else:
    n, d, Lmax = 1_000_000, 10, 30
    x, _ = W.c5_em_samples(n=n, d=d, G=6, seed=2025)
    idx, _ = amx.em_draw_init(n, Lmax, W.splitmix_uniforms_fast(2025, 4096))
    maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    if len(sys.argv) > 3:  # profile mode: one second-generation fit only
        b = fit(x, idx, Lmax, maxit, True, AMX_EM_TEAMS=int(sys.argv[3]))
        print(f"v2: {b['kernel_ms']:.2f} ms, {b['comp_steps']} steps")
        sys.exit(0)
    os.environ["AMX_EM_DEBUG"] = "1"
    a = fit(x, idx, Lmax, maxit, False)
    print(f"v1: {a['kernel_ms']:.2f} ms, {a['comp_steps']} steps, {1e3*a['kernel_ms']/a['comp_steps']:.1f} us/step", flush=True)
    for teams in (3,):
        for ns in (8,):
            for rep in range(2):
                b = fit(x, idx, Lmax, maxit, True, AMX_EM_TEAMS=teams, AMX_EM_STAGES=ns)
            same = np.array_equal(a["trace_L"], b["trace_L"])
            print(f"v2 teams={teams} stages<={ns}: {b['kernel_ms']:.2f} ms, {b['comp_steps']} steps, {1e3*b['kernel_ms']/b['comp_steps']:.1f} us/step, "
                  f"trace same={same}, loglik rel {rel(b['trace_loglik'], a['trace_loglik']):.2e}", flush=True)
