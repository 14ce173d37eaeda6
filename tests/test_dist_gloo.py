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


def _train_worker(rank, world, port, out_dir):
    """sharded training on CPU: every rank counts its slice of the sequences with the ORACLE, the count buffers are
    sum-all-reduced (the exchange bench.py and rdp_classifier --gpus do over NCCL), and every rank must end up with
    the counts of one pass over the whole set."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(REPO / "pangea-plus_b200"))
    sys.path.insert(0, str(REPO / "tests"))
    import oracle_rdp as ora
    from pangea_b200 import dist as pd
    from pangea_b200 import synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tx = synth.synth_taxonomy(77, 40, length=400)
    genus, off = synth.hashed_member_plan(77, 300, tx["G"], length=400)
    lo, hi = pd.shard_range(300, rank, world)
    data = synth.hashed_members(77, tx["centroids"], genus, off, lo, hi - lo)
    om = ora.Model(data, off[lo:hi + 1] - off[lo], genus[lo:hi], tx["G"])
    m, nw, M, N = om.counts()
    bufs = [torch.from_numpy(np.ascontiguousarray(m).view(np.uint8).reshape(-1).copy()),
            torch.from_numpy(np.ascontiguousarray(nw).view(np.uint8).reshape(-1).copy()),
            torch.from_numpy(np.ascontiguousarray(M).view(np.uint8).reshape(-1).copy()),
            torch.from_numpy(np.array([N], np.int64).view(np.uint8).copy())]
    pd.allreduce_counts(bufs)
    full = ora.Model(synth.hashed_members(77, tx["centroids"], genus, off, 0, 300), off, genus, tx["G"])
    fm, fnw, fM, fN = full.counts()
    assert np.array_equal(bufs[0].numpy().view(np.int32).reshape(fm.shape), fm)
    assert np.array_equal(bufs[1].numpy().view(np.int32), fnw) and np.array_equal(bufs[2].numpy().view(np.int32), fM)
    assert int(bufs[3].numpy().view(np.int64)[0]) == fN == 300
    if rank == 0:
        np.save(os.path.join(out_dir, "train_ok.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_training_counts_world2(tmp_path):
    import torch.multiprocessing as mp

    port = 29950 + os.getpid() % 40
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "train_ok.npy").exists()
