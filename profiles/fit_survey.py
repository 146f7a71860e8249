"""Stage 1 + 2 of the coal-mining pipeline on the device for several seeds: the fitted component counts, to set
beside the reference's own spread over seeds (oracle/_build/*.err of gen_golden_posterior.py: model 0: 3-4,
model 1: 2-4, model 2: 2-5, model 3: 2-5, model 4: 2-4, model 5: 2-3)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from automix_b200 import _lib as amx, workloads as W

wl = W.coalmine()
T = amx.Target(wl["target"])
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for s in range(nseeds):
    Ls, off = [], 0
    for k, d in enumerate(wl["dims"]):
        d = int(d)
        r = amx.rwm_adapt(T, k, 100000, 1, wl["init"][off:off + d], seed=1000 * s + k)
        off += d
        xs = r["samples"][0]
        idx, _ = amx.em_draw_init(len(xs), 30, W.splitmix_uniforms_fast(77 * s + k, 4096))
        e = amx.em_fit(xs, idx, Lmax=30, maxit=5000)
        Ls.append((e["L"], e["iters"]))
    print("seed", s, "fitted L", [a for a, _ in Ls], "iters", [b for _, b in Ls], flush=True)
