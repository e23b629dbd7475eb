"""Multi-GPU sharding of the OCP batch.

OCP instances are independent (no coupling between batch elements anywhere in OCPSolver::updateSolution,
/root/reference/src/solver/ocp_solver.cpp:111-145), so the batch is split into contiguous blocks, one per rank / GPU;
there is no data-path collective inside an iteration.  The only exchange is ONE all-gather of the Newton step
(direction records) so that every rank sees the whole step (north_star: "a single NCCL all-gather of the converged
step").  torch.distributed is plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from typing import Tuple


def shard_range(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of a global batch owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("[shard_range] invalid argument: need 0 <= rank < world")
    if batch < 0:
        raise ValueError("[shard_range] invalid argument: batch must be non-negative")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_step(local, out=None, group=None):
    """All-gather the per-rank direction tensor [b_local, n_grid, d_stride] into [sum b, n_grid, d_stride].
    Equal shard sizes use all_gather_into_tensor (one NCCL call, in place into `out`); ragged shards are padded to the
    largest shard first."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    dist.all_gather(sizes, mine, group=group)
    sizes = [int(s.item()) for s in sizes]
    total = sum(sizes)
    if out is None:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    else:
        # ragged shards (batch % world != 0): pad to the largest shard, gather, then compact
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]].copy_(local)
        buf = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(buf.view(-1), pad.view(-1), group=group)
        off = 0
        for r, n in enumerate(sizes):
            out[off:off + n].copy_(buf[r * mx:r * mx + n])
            off += n
    return out
