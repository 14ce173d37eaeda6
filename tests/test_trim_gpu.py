"""GPU parity tests of the Trim join (pg_trim_join, bin/trim2) against the golden files made by the
reference's Trim/trim2.4.pl and against oracle/trim_ref.c on larger seeded inputs."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_pipeline as op
import pangea_b200 as pg
from pangea_b200 import synth_trim as stt
from test_trim_cpu import CASES, GOLD

pytestmark = pytest.mark.gpu
BIN = pg.PKG_DIR / "bin"


@pytest.mark.parametrize("name,a,b,gap,trunc", CASES)
def test_trim_api_matches_golden(ctx, name, a, b, gap, trunc):
    ab = (GOLD / a).read_bytes()
    bb = (GOLD / b).read_bytes() if b else None
    got = ctx.trim_join(ab, bb, paired=b is not None, gap=gap, truncate=trunc)
    assert got == (GOLD / f"{name}.expected.fasta").read_bytes()


def test_trim_cli_like_the_readme(tmp_path):
    """perl trim2.3.pl -a ../input_A.txt -b ../input_B.txt -g 100  (README.md:33)"""
    (tmp_path / "Trim").mkdir()
    (tmp_path / "input_A.txt").write_bytes((GOLD / "reads_A.qseq.txt").read_bytes())
    (tmp_path / "input_B.txt").write_bytes((GOLD / "reads_B.qseq.txt").read_bytes())
    r = subprocess.run([str(BIN / "trim2"), "-a", "../input_A.txt", "-b", "../input_B.txt", "-g", "100"], cwd=tmp_path / "Trim",
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout == "QSEQ file format found.\nTrimming complete.\n"
    out = tmp_path / "Trim" / "output_files" / "trim2" / "input_A.txt_runblast.fasta"
    assert out.read_bytes() == (GOLD / "qseq_g100.expected.fasta").read_bytes()
    assert (tmp_path / "singletons" / "input_A.txt_single.txt").exists()
    # -qc / -lc cannot change the cutoffs (the script's getopts string cannot parse them either)
    r = subprocess.run([str(BIN / "trim2"), "-a", "../input_A.txt", "-b", "../input_B.txt", "-g", "100", "-lc", "10"],
                       cwd=tmp_path / "Trim", capture_output=True, text=True, timeout=120)
    assert out.read_bytes() == (GOLD / "qseq_g100.expected.fasta").read_bytes()
    # Getopt::Std bundling: "-lc 80" is -l then -c 80 and the parse goes on, so a -g AFTER it still counts; "-qc 25"
    # is -q with the value "c" and leaves "25", a non-option word, which ends the parse (the -g after it is lost)
    out.unlink()
    r = subprocess.run([str(BIN / "trim2"), "-lc", "80", "-a", "../input_A.txt", "-b", "../input_B.txt", "-g", "100"],
                       cwd=tmp_path / "Trim", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and out.read_bytes() == (GOLD / "qseq_g100.expected.fasta").read_bytes()
    r = subprocess.run([str(BIN / "trim2"), "-a", "../input_A.txt", "-b", "../input_B.txt", "-qc", "25", "-g", "100"],
                       cwd=tmp_path / "Trim", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and out.read_bytes() == (GOLD / "qseq_default.expected.fasta").read_bytes()
    # FASTQ: records are echoed on stdout as well
    (tmp_path / "x.fastq").write_bytes((GOLD / "reads.fastq").read_bytes())
    r = subprocess.run([str(BIN / "trim2"), "-a", "x.fastq"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    want = (GOLD / "fastq_single.expected.fasta").read_text()
    assert r.stdout == want + "Trimming complete.\n"
    assert (tmp_path / "output_files" / "trim2" / "x.fastq_runblast.fasta").read_text() == want


def test_trim_vs_oracle_large_and_packed_reads(ctx, tmp_path):
    a, b = stt.make_qseq_pair(91, 20000)
    (tmp_path / "a.txt").write_text(a)
    (tmp_path / "b.txt").write_text(b)
    assert op.oracle_trim(tmp_path / "a.txt", tmp_path / "b.txt", 189, 11, tmp_path / "want.fa") == 0
    want = (tmp_path / "want.fa").read_bytes()
    got, reads = ctx.trim_join(a.encode(), b.encode(), paired=True, want_reads=True)
    assert got == want
    # the packed read store holds exactly the joined sequences (dropped pairs are zero-length records)
    seqs = [l for l in want.split(b"\n")[1::2]]
    kept = 0
    for i in range(len(reads)):
        ln, codes, mask = reads.unpack(i, 700)
        if ln == 0:
            continue
        s = seqs[kept]
        kept += 1
        assert ln == len(s)
        for p in (0, 1, ln // 2, ln - 1):
            valid = (mask[p // 32] >> (p % 32)) & 1
            assert valid == (1 if s[p:p + 1] in b"ACGT" else 0)
        if kept > 300:
            break
    assert kept > 0
    reads.free()
    fq = stt.make_fastq(92, 30001)
    (tmp_path / "x.fastq").write_text(fq)
    for paired, gap in ((False, 189), (True, 50)):
        assert op.oracle_trim(tmp_path / "x.fastq", tmp_path / "x.fastq" if paired else None, gap, 11, tmp_path / "w.fa") == 0
        assert ctx.trim_join(fq.encode(), None, paired=paired, gap=gap) == (tmp_path / "w.fa").read_bytes()
    assert ctx.trim_join(b"", None) == b""


def test_trimmed_reads_classify_like_the_text_route(ctx):
    """Trim join -> packed store -> classify == Trim join -> FASTA text -> classify"""
    from pangea_b200 import synth

    tr = synth.synth16s(seed=13, seqs=90, genera=25, length=700)
    rng = np.random.default_rng(4)
    A, B = [], []
    for i in range(300):
        m = int(rng.integers(0, len(tr["genus"])))
        s = tr["data"][tr["off"][m]:tr["off"][m + 1]]
        p = int(rng.integers(0, len(s) - 400))
        ra, rb = s[p:p + 150].tobytes().decode().upper(), s[p + 250:p + 400].tobytes().decode().upper()
        q = "h" * 150
        head = f"M\t1\t1\t1\t{i}\t{i}\t0"
        A.append(f"{head}\t1\t{ra}\t{q}\t1")
        B.append(f"{head}\t2\t{rb}\t{q}\t1")
    a, b = ("\n".join(A) + "\n").encode(), ("\n".join(B) + "\n").encode()
    model = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    model.set_lineage(tr["anc"])
    text, reads = ctx.trim_join(a, b, paired=True, gap=100, want_reads=True)
    import torch

    res_dev = torch.zeros(len(reads) * 64, dtype=torch.uint8, device="cuda")
    ctx.classify_packed(model, reads, res_dev, mode=1)
    ctx.sync()
    packed = np.frombuffer(res_dev.cpu().numpy().tobytes(), dtype=pg.RESULT_DTYPE)
    seqs = text.split(b"\n")[1::2]
    data, off = pg.pack_sequences([s for s in seqs if s])
    via_text = ctx.classify(model, data, off, mode=1)
    assert len(packed) == 300 and len(via_text) == 300
    assert packed.tobytes() == via_text.tobytes()
    reads.free()
    model.free()
