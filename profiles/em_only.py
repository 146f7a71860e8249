"""EM fit alone (C5-EM shape) -- the command profiled by ncu for the K2 kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from automix_b200 import _lib as amx, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ndev = int(sys.argv[4]) if len(sys.argv) > 4 else 0
x, _ = W.c5_em_samples(n=n, d=d, seed=2025)
idx, _ = amx.em_draw_init(n, 30, W.splitmix_uniforms_fast(99, 4096))
for _ in range(2):
    r = amx.em_fit(x, idx, Lmax=30, maxit=maxit, devices=list(range(ndev)) if ndev else None)
print(f"gpus={ndev or 1} n={n} d={d} iters={r['iters']} steps={r['comp_steps']} ms={r['kernel_ms']:.2f} us/step={1e3*r['kernel_ms']/r['comp_steps']:.1f}")
