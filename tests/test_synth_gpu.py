"""GPU tests of the device-side synthetic generators (include/pangea_b200_synth.h) against their numpy twins, and of
sharded / streamed training (pg_train_accumulate): counts accumulated slice by slice equal the counts of one pass,
and equal the oracle's, bit for bit (BASELINE configs[3])."""
import numpy as np
import pytest

import oracle_rdp as ora
import pangea_b200 as pg
from pangea_b200 import synth

pytestmark = pytest.mark.gpu
SEED = 0x3000000


def _dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_device_members_and_reads_equal_the_numpy_twins(ctx):
    import torch

    tx = synth.synth_taxonomy(SEED, 300, length=700)
    genus, off = synth.hashed_member_plan(SEED, 2000, tx["G"], length=700)
    first, count = 137, 1500                                 # a slice in the middle: any rank can make any slice
    want = synth.hashed_members(SEED, tx["centroids"], genus, off, first, count)
    d_out = torch.empty(want.size, dtype=torch.uint8, device="cuda")
    loc_off = _dev(off[first:first + count + 1] - off[first])
    ctx.synth_members(SEED, _dev(tx["centroids"]), 700, _dev(genus[first:first + count]), loc_off, first, count, d_out)
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy(), want)
    assert (want == ord("n")).mean() > 0.0003 and set(np.unique(want)) <= set(b"acgtn")
    for paired, rl, gap in ((True, 250, 189), (False, 250, 0), (True, 100, 7)):
        wd, wo, ws = synth.hashed_reads(0x250, want, off[first:first + count + 1] - off[first], genus[first:first + count], 1000, 3000,
                                        read_len=rl, gap=gap, paired=paired)
        span = int(wo[1])
        d_reads = torch.empty(3000 * span, dtype=torch.uint8, device="cuda")
        d_src = torch.empty(3000, dtype=torch.int32, device="cuda")
        ctx.synth_reads(0x250, d_out, loc_off, _dev(genus[first:first + count]), count, 1000, 3000, d_reads, d_src, read_len=rl,
                        gap=gap, paired=paired)
        ctx.sync()
        assert np.array_equal(d_reads.cpu().numpy(), wd), (paired, rl, gap)
        assert np.array_equal(d_src.cpu().numpy(), ws)
        if span > 700:                                       # no member is long enough: all-N records, source -1
            assert (ws == -1).all() and (wd == ord("N")).all()


def test_sharded_and_streamed_training_equals_one_pass(ctx):
    import torch

    tx = synth.synth_taxonomy(SEED + 1, 150, length=600)
    genus, off = synth.hashed_member_plan(SEED + 1, 1200, tx["G"], length=600)
    data = synth.hashed_members(SEED + 1, tx["centroids"], genus, off, 0, 1200)
    one = ctx.train(data, off, genus, tx["G"])
    m1, nw1, M1, N1 = one.counts()
    om = ora.Model(data, off, genus, tx["G"])
    rm, rnw, rM, rN = om.counts()
    assert np.array_equal(m1, rm) and np.array_equal(nw1, rnw) and np.array_equal(M1, rM) and N1 == rN == 1200
    # (a) streamed from the host in three uneven slices, offsets NOT rebased by the caller
    acc = ctx.model_create(tx["G"])
    for lo, hi in ((0, 100), (100, 777), (777, 1200)):
        ctx.train_accumulate(acc, data, off[lo:hi + 1], genus[lo:hi])
    acc.commit()
    m2, nw2, M2, N2 = acc.counts()
    assert np.array_equal(m2, m1) and np.array_equal(nw2, nw1) and np.array_equal(M2, M1) and N2 == 1200
    # (b) two "ranks": each generates its slice on the device and counts it into its own model; the integer sum of the
    # count buffers (what the NCCL all-reduce computes) is the one-pass model; tables derived from it are bit-identical
    parts = []
    cent = _dev(tx["centroids"])
    for lo, hi in ((0, 600), (600, 1200)):
        loc_off = _dev(off[lo:hi + 1] - off[lo])
        d_bytes = torch.empty(int(off[hi] - off[lo]), dtype=torch.uint8, device="cuda")
        d_genus = _dev(genus[lo:hi])
        ctx.synth_members(SEED + 1, cent, 600, d_genus, loc_off, lo, hi - lo, d_bytes)
        md = ctx.model_create(tx["G"])
        ctx.train_accumulate(md, d_bytes, loc_off, d_genus, device=True)
        parts.append(md)
    ctx.sync()
    from pangea_b200 import dist as pgdist

    bufs = [[torch.as_tensor(pgdist.DeviceBuffer(p, n), device="cuda") for p, n in md.buffers()] for md in parts]
    for i, (a, b) in enumerate(zip(*bufs)):
        dt = torch.int64 if i == 3 else torch.int32
        a.view(dt).add_(b.view(dt))
    torch.cuda.synchronize()
    parts[0].commit()
    m3, nw3, M3, N3 = parts[0].counts()
    assert np.array_equal(m3, m1) and np.array_equal(nw3, nw1) and np.array_equal(M3, M1) and N3 == 1200
    t1, t3 = one.tables(), parts[0].tables()
    for x, y in zip(t1, t3):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    # the model classifies like the one-pass model and like the oracle
    parts[0].set_lineage(tx["anc"])
    one.set_lineage(tx["anc"])
    rd, ro, rs = synth.hashed_reads(9, data, off, genus, 0, 400, read_len=200, gap=50)
    a, ba = ctx.classify(one, rd, ro, mode=1, want_boot=True)
    b, bb = ctx.classify(parts[0], rd, ro, mode=1, want_boot=True)
    assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb)
    ref = om.classify(rd, ro)
    assert np.array_equal(a["genus"], ref["genus"]) and np.array_equal(ba, ref["boot"])
    assert (a["genus"] == rs).mean() > 0.9
    om.free()
    for md in parts + [one, acc]:
        md.free()
