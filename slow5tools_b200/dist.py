"""Multi-GPU plumbing for the codec path.  Records are independent (one zlib stream per record, svb-zd chains
restart at every read: slow5.c:4046, slow5_press.c:1106,1162), so a batch is cut into contiguous, byte-balanced
ranges -- one per rank / GPU -- and there is NO collective on the data path; torch.distributed is used only for
the barrier and the max-over-ranks timing."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(sizes, world):
    """Cuts records [0, n) into `world` contiguous ranges of ~equal total size (bytes or samples), preserving
    order so that concatenating the per-rank outputs rebuilds the batch.  Returns world+1 boundaries."""
    sizes = np.asarray(sizes, dtype=np.float64)
    n = len(sizes)
    if n == 0:
        return [0] * (world + 1)
    csum = np.concatenate([[0.0], np.cumsum(sizes)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        # pick the boundary whose prefix is closest to the target, never going backwards
        if k > 0 and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return bounds


def max_over_ranks(value, device=None):
    """Largest `value` (float or list of floats) over all ranks; identity when not distributed."""
    vals = list(value) if isinstance(value, (list, tuple)) else [value]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vals if isinstance(value, (list, tuple)) else vals[0]
    t = torch.tensor(vals, dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = t.tolist()
    return out if isinstance(value, (list, tuple)) else out[0]


def sum_over_ranks(value, device=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])
