"""Host-side multi-GPU logic on CPU: utterance partitioning and the rank-ordered host gather
(world_size 2, gloo).  The per-rank compute stand-in is the CPU oracle — the GPU path itself is
exercised by the -m gpu tests; what is checked here is that sharding + gathering reproduces the
unsharded result bit for bit (there is no data-path collective to get wrong)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
from voxbox_b200 import shard  # noqa: E402


def test_partition_properties():
    rng = np.random.default_rng(0)
    for U, world in ((360, 8), (7, 2), (3, 8), (0, 4), (1000, 3), (1, 1)):
        counts = rng.integers(1, 2000, size=U)
        parts = shard.partition(counts, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == U
        assert all(a <= b for a, b in parts) and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        if U >= 4 * world:
            loads = [counts[a:b].sum() for a, b in parts]
            assert max(loads) - min(loads) <= 2 * counts.max()
    # equal utterances split evenly
    assert shard.partition([998] * 360, 8) == [(45 * r, 45 * (r + 1)) for r in range(8)]
    assert shard.frames_of(160000, 400, 160) == 998 and shard.frames_of(399, 400, 160) == 0 and shard.frames_of(2049, 2048, 1024) == 1


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import oracle
    from voxbox_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lengths = [4000, 8000, 5600, 12000, 4800]  # ragged utterances
        N, hop, p = 400, 160, 12
        counts = [shard.frames_of(n, N, hop) for n in lengths]
        a, b = shard.my_range(counts, world, rank)
        rows = []
        for u in range(a, b):
            x = synth.utterance(u, 16000, seconds=lengths[u] / 16000.0)
            r, ac = oracle.batch_lpc(x, counts[u], N, hop, oracle.WIN_HANN_SYMMETRIC, p)
            rows.append(ac)
        local = np.concatenate(rows) if rows else np.zeros((0, p + 1))
        full = shard.gather_rows(local, dist, dst=0)
        if rank == 0:
            np.save(out_path, full)
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo_world2(tmp_path, oracle):
    import torch.multiprocessing as mp
    from voxbox_b200 import synth
    out = str(tmp_path / "gathered.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    lengths = [4000, 8000, 5600, 12000, 4800]
    ref = np.concatenate([oracle.batch_lpc(synth.utterance(u, 16000, seconds=n / 16000.0), shard.frames_of(n, 400, 160), 400, 160,
                                           oracle.WIN_HANN_SYMMETRIC, 12)[1] for u, n in enumerate(lengths)])
    assert got.shape == ref.shape and np.array_equal(got, ref)
