"""World-size-2 gloo test of the N>1 path: batch sharding + the single all-gather of the Newton step (CPU tensors)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, ragged, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from robotoc_b200.shard import allgather_step, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(batch, world, rank)
    full = torch.arange(batch * 3 * 4, dtype=torch.float64).reshape(batch, 3, 4)  # stands for [batch, n_grid, d_stride]
    local = full[lo:hi].clone()
    got = allgather_step(local, global_batch=batch)
    ok = bool(torch.equal(got, full))
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([ok, lo, hi]))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_allgather_step_world2_gloo(tmp_path, batch):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, batch, batch % 2 == 1, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok, lo, hi = np.load(tmp_path / f"ok_{r}.npy")
        assert ok == 1
