"""Utterance sharding across the GPUs of a box: contiguous ranges of utterances per rank,
balanced by frame count, no data-path collective — every rank runs the whole kernel chain on its range
and the results are gathered on the host.

The range arithmetic is pure Python; `gather_rows` uses whatever torch.distributed backend the caller
initialised (NCCL on the GPU box, gloo in the CPU tests) and moves only the small result arrays."""
import numpy as np


def frames_of(n_samples, frame_len, hop):
    """Windower::{hanning,rectangle}: a frame while bin <= remaining, advance by hop (ragged tail dropped)."""
    return 0 if n_samples < frame_len else (n_samples - frame_len) // hop + 1


def partition(frame_counts, world):
    """Split utterances 0..U-1 into `world` contiguous ranges with near-equal total frames.

    Returns [(start, end)] * world (end exclusive; ranges may be empty when U < world).  Greedy on the
    prefix sum: boundary r is the first utterance whose cumulative frame count reaches r/world of the total."""
    counts = np.asarray(frame_counts, dtype=np.int64)
    if counts.ndim != 1 or world < 1:
        raise ValueError("frame_counts must be 1-D and world >= 1")
    U = counts.size
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        # pick the closer of the two neighbouring boundaries, keep monotone
        if b > 0 and abs(csum[b - 1] - target) <= abs(csum[min(b, U)] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), U))
    bounds.append(U)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def my_range(frame_counts, world, rank):
    return partition(frame_counts, world)[rank]


def gather_rows(local_rows, dist=None, dst=0):
    """Host-side gather of per-rank result rows (numpy [n_local, ...]) in rank order onto rank `dst`.

    Returns the concatenated array on `dst`, None elsewhere.  With dist=None (single process) returns the input."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local_rows)
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(np.asarray(local_rows), bucket, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([b for b in bucket if b is not None and len(b)], axis=0) if any(len(b) for b in bucket) else bucket[0]
