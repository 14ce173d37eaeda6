"""world_size-2 test of the multi-GPU plumbing on CPU (gloo): contiguous read sharding, the
one-off broadcast of the model's count buffers, and the rank-ordered gather of result records."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))


def _worker(rank, world, port, n_total, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(REPO / "pangea-plus_b200"))
    from pangea_b200 import RESULT_DTYPE
    from pangea_b200 import dist as pd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # model replication: rank 0 holds the counts, the others receive them
    rng = np.random.default_rng(5)
    counts = [rng.integers(0, 7, 4096).astype(np.int32), rng.integers(0, 99, 65536).astype(np.int32), np.array([9178], np.int64)]
    tensors = [torch.from_numpy(c.copy() if rank == 0 else np.zeros_like(c)) for c in counts]
    pd.broadcast_buffers(tensors, src=0)
    for t, c in zip(tensors, counts):
        assert np.array_equal(t.numpy(), c)
    # every rank "classifies" its shard: record i carries genus = i so order is checkable
    lo, hi = pd.shard_range(n_total, rank, world)
    rec = np.zeros(hi - lo, RESULT_DTYPE)
    rec["genus"] = np.arange(lo, hi)
    rec["votes"][:, 0] = 100
    local = torch.from_numpy(rec.view(np.uint8).copy())
    allrec = pd.gather_records(local, n_total, rank, world, dst=0)
    if rank == 0:
        got = allrec.numpy().view(RESULT_DTYPE)
        assert len(got) == n_total and np.array_equal(got["genus"], np.arange(n_total))
        np.save(os.path.join(out_dir, "ok.npy"), np.array([pd.records_checksum(got)]))
    else:
        assert allrec is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1001, 64])
def test_shard_broadcast_gather_world2(tmp_path, n_total):
    import torch.multiprocessing as mp

    port = 29600 + (os.getpid() + n_total) % 300
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok.npy").exists()


def test_shard_ranges_partition_the_reads():
    from pangea_b200 import dist as pd

    for n in (0, 1, 7, 1000, 1_000_000):
        for w in (1, 2, 4, 8):
            r = [pd.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
