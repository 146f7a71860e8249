"""The N>1 host logic on CPU (gloo, world_size 2 and 3): shard plan, histogram sum, time max."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from automix_b200 import shard  # noqa: E402


def test_shard_ranges_partition_the_population():
    for total in (0, 1, 7, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            seen = 0
            nxt = 0
            for r in range(world):
                first, count = shard.shard_range(total, world, r)
                assert first == nxt and count >= 0
                nxt = first + count
                seen += count
            assert seen == total
    assert shard.weak_range(1 << 20, 3) == (3 << 20, 1 << 20)
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_histogram_equals_single_process(tmp_path, world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_worker

    total, nmodels, nsweeps = 1001, 5, 7
    out = tmp_path / "res.json"
    port = 29500 + world + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(out), str(total), str(nmodels), str(nsweeps)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(out.read_text())
    want = dist_worker.fake_visits(0, total, nmodels, nsweeps)
    assert res["world"] == world and res["owned"] == total
    assert np.array_equal(np.array(res["hist"]), want)
    assert abs(res["tmax"] - 0.001 * world) < 1e-12
