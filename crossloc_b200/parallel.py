"""Image sharding and the pose gather for multi-GPU evaluation.

Localization is embarrassingly parallel over images, so the only exchange is one all-gather of
[n_local, 18] fp32 rows (16 pose entries + translation and rotation error) per evaluated section; rank 0
then computes the accuracy buckets and medians exactly as the reference's printout does
(/root/reference/utils/evaluation.py:207-230).  Works with any torch.distributed backend: NCCL on the GPUs,
gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_total, rank, world):
    """Images of rank `rank`: i with i mod world == rank (round-robin keeps ranks balanced for any n)."""
    return list(range(rank, n_total, world))


def gather_rows(local_indices, local_rows, n_total, group=None):
    """All-gather per-image rows into global image order.  local_rows: [n_local, K] float32 tensor.

    Returns a [n_total, K] tensor (on local_rows' device) on every rank.
    """
    if not (dist.is_available() and dist.is_initialized()):
        out = torch.empty(n_total, local_rows.size(1), dtype=local_rows.dtype, device=local_rows.device)
        out[torch.as_tensor(local_indices, dtype=torch.long)] = local_rows
        return out
    world = dist.get_world_size(group)
    k = local_rows.size(1)
    n_max = (n_total + world - 1) // world
    # pad to equal length; column k carries the global index (-1 = padding)
    buf = torch.full((n_max, k + 1), -1.0, dtype=torch.float32, device=local_rows.device)
    n_local = len(local_indices)
    buf[:n_local, :k] = local_rows
    buf[:n_local, k] = torch.as_tensor(local_indices, dtype=torch.float32)
    gathered = torch.empty(world * n_max, k + 1, dtype=torch.float32, device=local_rows.device)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(gathered, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        gathered = torch.cat(parts, 0)
    idx = gathered[:, k].long()
    keep = idx >= 0
    out = torch.empty(n_total, k, dtype=torch.float32, device=local_rows.device)
    out[idx[keep]] = gathered[keep, :k]
    return out


def summarize(t_err, r_err):
    """Accuracy buckets and medians of utils/evaluation.py:212-230 (percentages, degrees, metres)."""
    t_err, r_err = np.asarray(t_err, dtype=np.float64), np.asarray(r_err, dtype=np.float64)
    n = max(1, len(t_err))
    pct = lambda t, r: float(np.sum(np.logical_and(t_err < t, r_err < r)) / n * 100)   # noqa: E731
    return {
        '30m10deg': pct(30.0, 10.0), '20m10deg': pct(20.0, 10.0), '10m7deg': pct(10.0, 7.0),
        '10m10deg': pct(10.0, 10.0), '5m5deg': pct(5.0, 5.0), '3m3deg': pct(3.0, 3.0),
        'median_r_deg': float(np.median(r_err)), 'median_t_m': float(np.median(t_err)),
        'mean_r_deg': float(np.mean(r_err)), 'std_r_deg': float(np.std(r_err)),
        'mean_t_m': float(np.mean(t_err)), 'std_t_m': float(np.std(t_err)), 'count': int(len(t_err)),
    }


def evaluate_sharded(localize_batch, gt_pose_of, n_total, batch, rank=0, world=1, device='cpu', group=None):
    """Localize images {i : i mod world == rank} in batches and gather [pose(16), t_err, r_err] rows.

    localize_batch(indices) -> float32 tensor [len(indices), 4, 4] of camera-to-world poses.
    gt_pose_of(i) -> 4x4 ground-truth pose.  Returns (rows [n_total, 18] tensor, summary dict).
    """
    from .synth import pose_errors
    mine = shard_indices(n_total, rank, world)
    rows = torch.empty(len(mine), 18, dtype=torch.float32)
    for s in range(0, len(mine), batch):
        idx = mine[s:s + batch]
        poses = localize_batch(idx).detach().to('cpu', torch.float32)
        for j, i in enumerate(idx):
            t, r = pose_errors(gt_pose_of(i), poses[j].numpy())
            rows[s + j, :16] = poses[j].reshape(16)
            rows[s + j, 16] = t
            rows[s + j, 17] = r
    allrows = gather_rows(mine, rows.to(device), n_total, group)
    cpu = allrows.cpu().numpy()
    return allrows, summarize(cpu[:, 16], cpu[:, 17])
