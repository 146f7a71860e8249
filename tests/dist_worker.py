"""Worker for tests/test_distributed_cpu.py: run under torch.distributed.run with the gloo backend."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from automix_b200 import shard  # noqa: E402


def fake_visits(first, count, nmodels, nsweeps):
    """Stand-in for a rank's sweep kernel: a deterministic function of the GLOBAL chain id only."""
    ids = np.arange(first, first + count, dtype=np.int64)
    hist = np.zeros(nmodels, np.int64)
    for s in range(nsweeps):
        k = (ids * 2654435761 + s * 40503) % nmodels
        hist += np.bincount(k, minlength=nmodels)
    return hist


def main():
    out_path, total, nmodels, nsweeps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    first, count = shard.shard_range(total, world, rank)
    hist = torch.from_numpy(fake_visits(first, count, nmodels, nsweeps))
    shard.allreduce_sum_(hist)
    t = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    shard.allreduce_max_(t)
    owned = torch.tensor([count], dtype=torch.int64)
    shard.allreduce_sum_(owned)
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"hist": hist.tolist(), "tmax": float(t[0]), "owned": int(owned[0]), "world": world}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
