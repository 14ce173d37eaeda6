"""CPU tests of the Trim-join oracle (oracle/trim_ref.c) against golden files made by the reference's
own Trim/trim2.4.pl (tests/golden/make_golden.py) and, where the reference tree and perl exist,
against the live script on fresh seeds."""
from pathlib import Path

import pytest

import oracle_pipeline as op
from pangea_b200 import synth_trim as stt

GOLD = Path(__file__).parent / "golden" / "trim"
CASES = [("qseq_g100", "reads_A.qseq.txt", "reads_B.qseq.txt", 100, 11),
         ("qseq_default", "reads_A.qseq.txt", "reads_B.qseq.txt", 189, 11),
         ("qseq_g7_t5", "reads_A.qseq.txt", "reads_B.qseq.txt", 7, 5),
         ("fastq_single", "reads.fastq", None, 189, 11),
         ("fastq_paired_g25", "reads.fastq", "reads.fastq", 25, 11)]


@pytest.mark.parametrize("name,a,b,gap,trunc", CASES)
def test_trim_oracle_matches_golden(name, a, b, gap, trunc, tmp_path):
    out = tmp_path / "o.fa"
    assert op.oracle_trim(GOLD / a, GOLD / b if b else None, gap, trunc, out) == 0
    assert out.read_bytes() == (GOLD / f"{name}.expected.fasta").read_bytes()


def test_golden_documents_the_script_quirks():
    q = (GOLD / "qseq_g100.expected.fasta").read_text().split("\n")
    assert q[0].startswith(">HWI-ST41:7:") and q[0].endswith(":1:AB")          # header = columns 1-8 joined by ':'
    assert "N" * 100 in q[1] and "." not in q[1]                                 # gap of Ns; '.' no-calls became N
    n_in = len((GOLD / "reads_A.qseq.txt").read_text().split("\n")) - 1
    assert 0 < len(q) // 2 < n_in                                                # pairs with a short mate are dropped
    p = (GOLD / "fastq_paired_g25.expected.fasta").read_text().split("\n")
    assert all(l.endswith("\t") or l.endswith("0") for l in p[1::2] if l)        # mate 2 keeps the TAB of "SEQ\t"
    s = (GOLD / "fastq_single.expected.fasta").read_text().split("\n")
    assert "0" in s[1::2]                                                        # a too-short FASTQ read prints as 0


@pytest.mark.skipif(not op.have_reference(), reason="reference tree / perl not available here")
@pytest.mark.parametrize("seed", [7, 8])
def test_trim_oracle_against_the_live_script(seed, tmp_path):
    a, b = stt.make_qseq_pair(seed, 80)
    (tmp_path / "a.txt").write_text(a)
    (tmp_path / "b.txt").write_text(b)
    for gap, trunc in ((None, None), (33, None), (5, 3)):
        op.real_trim(tmp_path / "a.txt", tmp_path / "b.txt", gap, tmp_path / "real.fa", trunc)
        assert op.oracle_trim(tmp_path / "a.txt", tmp_path / "b.txt", 189 if gap is None else gap, 11 if trunc is None else trunc,
                              tmp_path / "mine.fa") == 0
        assert (tmp_path / "mine.fa").read_bytes() == (tmp_path / "real.fa").read_bytes()
    (tmp_path / "x.fastq").write_text(stt.make_fastq(seed, 61))                 # odd count: the last pair has no mate
    for paired, gap in ((False, None), (True, 12)):
        op.real_trim(tmp_path / "x.fastq", tmp_path / "x.fastq" if paired else None, gap, tmp_path / "real.fa")
        assert op.oracle_trim(tmp_path / "x.fastq", tmp_path / "x.fastq" if paired else None, 189 if gap is None else gap, 11,
                              tmp_path / "mine.fa") == 0
        assert (tmp_path / "mine.fa").read_bytes() == (tmp_path / "real.fa").read_bytes()
