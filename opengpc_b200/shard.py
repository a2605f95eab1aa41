"""Multi-GPU sharding of the path: stereo pairs are independent units, so rank r of
`world` processes pairs r, r + world, r + 2*world, ... on its own resident context.  There is no
data-path collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the
barrier around the timed region, the max-over-ranks of the measured time and the sum of counts.
"""
import numpy as np


def shard_pairs(n_pairs, rank, world):
    """Indices of the pairs rank `rank` owns (round-robin, as SURVEY.md 8e prescribes)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_pairs, world, dtype=np.int64)


def owner_of(pair, world):
    return int(pair) % int(world)


def reduce_timing(ms_local, count_local, dist=None, device="cpu"):
    """(max over ranks of the elapsed ms, sum over ranks of the processed units)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms_local), int(count_local)
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    n = torch.tensor([int(count_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), int(n.item())


def gather_support_counts(counts_local, pair_ids_local, n_pairs, dist=None, device="cpu"):
    """Per-pair support counts of the whole job on every rank (results are gathered to the host,
    the only cross-rank exchange of the path): int64 [n_pairs]."""
    out = np.zeros(n_pairs, np.int64)
    out[np.asarray(pair_ids_local, np.int64)] = np.asarray(counts_local, np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return out
    import torch
    t = torch.from_numpy(out).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
