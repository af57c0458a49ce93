"""Multi-GPU plumbing: instances shard trivially across ranks (no collective on the step path).

`torch.distributed` is used only at the edges: a barrier + max-reduction of the measured time, and
an optional gather of final states / error statistics at the end of a run.
Works with the NCCL backend on GPUs and with gloo on CPU (used by the tests)."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous instance range [lo, hi) owned by `rank`: floor(g*N/G) .. floor((g+1)*N/G)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def gather_final_states(y_local, n_total, dist=None, device="cpu"):
    """All-gather the per-rank final states y_local [ny, n_local] into [ny, n_total] on every rank."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(y_local)
    world, rank = dist.get_world_size(), dist.get_rank()
    ny = y_local.shape[0]
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    nmax = max(sizes)
    buf = torch.zeros((ny, nmax), dtype=torch.float64, device=device)
    buf[:, :sizes[rank]] = torch.as_tensor(np.ascontiguousarray(y_local), device=device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([p[:, :s].cpu().numpy() for p, s in zip(parts, sizes)], axis=1)


def reduce_stats(max_err_norm, n_bad, elapsed_ms, dist=None, device="cpu"):
    """(max error norm, total non-finite instances, max elapsed time) over all ranks."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(max_err_norm), int(n_bad), float(elapsed_ms)
    mx = torch.tensor([max_err_norm, elapsed_ms], dtype=torch.float64, device=device)
    sm = torch.tensor([n_bad], dtype=torch.float64, device=device)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return float(mx[0]), int(sm[0]), float(mx[1])
