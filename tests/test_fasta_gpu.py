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


def test_cli_pieces_give_the_same_file(tmp_path):
    """rdp_classifier hands its query file to the GPU in record-aligned pieces; tiny pieces (PG_CLI_PIECE_BYTES)
    must give the file a single piece gives, multi-line records and CRLF included."""
    import os

    tr = synth.synth16s(seed=8, seqs=120, genera=30, length=700)
    names, anc = tr["node_names"], tr["anc"]
    with open(tmp_path / "train.fa", "w") as f:
        for i, g in enumerate(tr["genus"]):
            s = tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes().decode()
            f.write(f">T{i:04d}\t" + ";".join(names[n] for n in anc[g]) + "\n" + s + "\n")
    data, off, _ = synth.synth_reads(4, tr, 800, paired=False, read_len=230)
    arr = data.reshape(-1, int(off[1]))
    with open(tmp_path / "q.fa", "wb") as f:
        for i in range(arr.shape[0]):
            b = arr[i].tobytes()
            f.write(b">q%04d some description\r\n" % i + b[:100] + b"\r\n" + b[100:] + b"\n")
        f.write(b">tiny\nACGTACGT\n")
    exe = str(BIN / "rdp_classifier")
    subprocess.run([exe, "--train", "train.fa", "-t", "m.pgm"], cwd=tmp_path, check=True, capture_output=True)
    r1 = subprocess.run([exe, "-q", "q.fa", "-o", "one.txt", "-t", "m.pgm"], cwd=tmp_path, capture_output=True, text=True)
    r2 = subprocess.run([exe, "-q", "q.fa", "-o", "many.txt", "-t", "m.pgm"], cwd=tmp_path, capture_output=True, text=True,
                        env=dict(os.environ, PG_CLI_PIECE_BYTES="5000"))
    assert r1.returncode == 0 and r2.returncode == 0, (r1.stderr, r2.stderr)
    one = (tmp_path / "one.txt").read_text()
    assert one == (tmp_path / "many.txt").read_text() and r1.stdout == r2.stdout
    lines = one.split("\n")
    assert len([l for l in lines if l]) == 800 and lines[0].startswith("q0000\t")
    assert "recordID=tiny" in r1.stdout                      # the short record is reported on stdout, not classified
