"""First hit per read on the GPU (pg_first_hits, bin/get_uniq) against the C restatement (itself pinned against the
live Scripts/get_uniq.pl in tests/test_uniq_cpu.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))

import oracle_pipeline as op  # noqa: E402
from test_uniq_cpu import EDGE, hits_text  # noqa: E402

pytestmark = pytest.mark.gpu
BIN = REPO / "pangea-plus_b200" / "bin"


@pytest.fixture(scope="module")
def ctx():
    import pangea_b200 as pg

    c = pg.Context(0)
    yield c
    c.close()


def test_edge_lines_and_known_answer(ctx):
    for text in (EDGE, EDGE + b"\n", b"", b"a", b"a\n", b"\n\n\n", b"x\ty\nx\tz\n"):
        assert ctx.first_hits(text) == op.oracle_first_hits(text)
    assert ctx.first_hits(EDGE)[1] == [0, 2, 3, 4, 5, 7, 9, 11]


def test_large_input(ctx):
    text = hits_text(7, 200000)
    got, lines = ctx.first_hits(text)
    ref, rlines = op.oracle_first_hits(text)
    assert got == ref and lines == rlines
    # size-independent properties: idempotent; one line per distinct first column
    assert ctx.first_hits(got)[0] == got
    firsts = [l.split(b"\t", 1)[0] for l in got.split(b"\n") if l]
    assert len(firsts) == len(set(firsts)) == len(lines)


def test_cli(tmp_path):
    (tmp_path / "hits.txt").write_bytes(EDGE)
    r = subprocess.run([str(BIN / "get_uniq"), "-f", "hits.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "\nLoading input file...\n"
    assert (tmp_path / "hits.txt.unique").read_bytes() == op.oracle_first_hits(EDGE)[0]
    r = subprocess.run([str(BIN / "get_uniq")], cwd=tmp_path, capture_output=True, text=True)
    assert r.stdout.startswith("Usage: perl taxcollector_ncbi-0.01.pl")
    r = subprocess.run([str(BIN / "get_uniq"), "-f", "missing.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert "Error: Unable to open database file missing.txt." in r.stdout
