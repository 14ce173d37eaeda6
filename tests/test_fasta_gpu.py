"""FASTA ingest on the GPU (pg_fasta_ingest) against a plain reading of the same text, and the rdp_classifier
executable (which now ingests its query file on the GPU) against pg_classify on host-parsed records."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))

from pangea_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu
BIN = REPO / "pangea-plus_b200" / "bin"
WS = b" \t\n\r\f\v"


def plain_parse(text: bytes):
    """the host parser's reading (host/pg_host_common.c pg_fasta_read), restated"""
    ids, hdrs, seqs = [], [], []
    for line in text.split(b"\n"):
        body = line.rstrip(b"\r")
        if line[:1] == b">":
            hdrs.append(body[1:])
            k = 0
            while k < len(hdrs[-1]) and hdrs[-1][k:k + 1] not in [bytes([c]) for c in WS]:
                k += 1
            ids.append(hdrs[-1][:k])
            seqs.append(b"")
        elif ids:
            seqs[-1] += bytes(c for c in body if c not in WS)
    return ids, hdrs, seqs


@pytest.fixture(scope="module")
def ctx():
    import pangea_b200 as pg

    c = pg.Context(0)
    yield c
    c.close()


TRICKY = (b"junk before the first header\nACGT\n"
          b">r1 first record\nACGTACGT\nacgtnn\n"
          b">r2\r\nAC GT\tAC\r\nGGCC\r\n\r\n"
          b">\n"                                 # empty header, empty sequence
          b">r4\twith tab\n"
          b">r5 multi\nAAAA\nCCCC\nGGGG\nTTTT\n\n\nNNNN\n"
          b">r6 last without newline\nACGTTGCA")


def check(ctx, text):
    ids, hdrs, buf, off, reads = ctx.fasta_ingest(text)
    wi, wh, ws = plain_parse(text)
    assert ids == wi and hdrs == wh
    assert len(off) == len(ws) + 1
    got = [buf[off[i]:off[i + 1]].tobytes() for i in range(len(ws))]
    assert got == ws
    if reads is not None:
        assert len(reads) == len(ws)
        for i in (0, len(ws) // 2, len(ws) - 1):
            if 0 <= i < len(ws):
                ln, codes, mask = reads.unpack(i, len(ws[i]))
                assert ln == len(ws[i])
        reads.free()
    return ids, ws


def test_tricky_text(ctx):
    check(ctx, TRICKY)
    check(ctx, b"")
    check(ctx, b"no header at all\nACGT\n")
    check(ctx, b">only header")
    check(ctx, b">a\n>b\n>c\nACGT\n")


def test_large_multiline_and_illumina(ctx):
    tr = synth.synth16s(seed=5, seqs=300, genera=40, length=1400)
    rows = []
    for i in range(300):
        s = tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes()
        rows.append(b">S%04d Root;x\n" % i + b"\n".join(s[k:k + 60] for k in range(0, len(s), 60)) + b"\n")
    check(ctx, b"".join(rows))
    data, off, _ = synth.synth_reads(3, tr, 20000, paired=True)
    L = int(off[1])
    arr = data.reshape(-1, L)
    text = b"".join(b">r%07d:AB\n" % i + arr[i].tobytes() + b"\n" for i in range(arr.shape[0]))
    ids, ws = check(ctx, text)
    assert len(ids) == 20000 and ws[123] == arr[123].tobytes()


def test_ingested_reads_classify_like_host_parsed_reads(ctx):
    tr = synth.synth16s(seed=7, seqs=200, genera=50, length=900)
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    data, off, _ = synth.synth_reads(9, tr, 3000, paired=False)
    L = int(off[1])
    arr = data.reshape(-1, L)
    text = b"".join(b">q%05d\n" % i + arr[i].tobytes()[:130] + b"\r\n" + arr[i].tobytes()[130:] + b"\n" for i in range(arr.shape[0]))
    text += b">short\nACGT\n"
    ids, hdrs, buf, foff, reads = ctx.fasta_ingest(text)
    got = ctx.classify_packed_host(gm, reads, mode=1)
    from pangea_b200 import pack_sequences

    d2, o2 = pack_sequences([arr[i].tobytes() for i in range(arr.shape[0])] + [b"ACGT"])
    want = ctx.classify(gm, d2, o2, mode=1)
    assert got.tobytes() == want.tobytes()
    assert got["status"][-1] != 0                       # the short record is reported, not classified
    reads.free()
    gm.free()
