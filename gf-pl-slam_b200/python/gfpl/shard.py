"""Multi-GPU plumbing of the batched replay: frame pairs are independent, so ranks own
disjoint contiguous blocks of pair ids and nothing is exchanged inside the path.  The only collective is
the gather of the fixed-size per-pair pose records (K7).  Backend-agnostic (nccl on GPUs, gloo in CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

from . import _structs as S


def partition(n_pairs, world, rank):
    """contiguous block [lo, hi) of pair ids owned by `rank` (block sizes differ by at most one)"""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pair_seed(base_seed, pair_id):
    """seed of one frame pair: independent of how the pairs are sharded"""
    return base_seed + pair_id


def gather_pose_records(local_bytes, world, out=None):
    """all_gather of equal-sized uint8 tensors holding gfpl_pose_result records; returns the [world*n] tensor"""
    if world == 1:
        return local_bytes
    if out is None:
        out = torch.empty(world * local_bytes.numel(), dtype=torch.uint8, device=local_bytes.device)
    dist.all_gather_into_tensor(out, local_bytes)
    return out


def gather_pose_records_ragged(local_records, world):
    """ranks with different pair counts (uneven partition): gather sizes, pad, gather, trim; numpy in/out"""
    n = torch.tensor([len(local_records)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    buf = np.zeros(m, S.pose_result_dt)
    buf[:len(local_records)] = local_records
    t = torch.from_numpy(buf.view(np.uint8).reshape(-1).copy())
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.numpy().view(S.pose_result_dt)[:sizes[r]] for r, o in enumerate(outs)]
    return np.concatenate(parts)


def max_over_ranks(value, world, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
