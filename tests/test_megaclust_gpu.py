"""Megaclust on the GPU (pg_megaclust, bin/megaclust2) against the golden files made by the reference's own
Perl, and against the C restatement on inputs the Perl would need minutes for."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))

import oracle_pipeline as op  # noqa: E402
from pangea_b200 import synth_mega  # noqa: E402
from test_megaclust_cpu import CASES, fmt  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = REPO / "tests" / "golden" / "megaclust"
BIN = REPO / "pangea-plus_b200" / "bin"


@pytest.fixture(scope="module")
def ctx():
    import pangea_b200 as pg

    c = pg.Context(0)
    yield c
    c.close()


def test_golden_from_the_real_script(ctx):
    text = (GOLD / "input.txt").read_bytes()
    want = json.loads((GOLD / "expected.json").read_text())
    for name, case in CASES.items():
        res, ex, by = ctx.megaclust(text, **case["kw"])
        assert [l.decode() for l in fmt(res)] == want[name]["lines"], name
        assert want[name]["stdout"] == f"Run complete:\n{ex} hits examined\n{by} hits beyond thresholds and therefore not counted.\n"
        assert res == op.oracle_megaclust(text, **case["kw"])[0]           # same order too: first appearance


def test_large_input_against_the_oracle(ctx):
    text = synth_mega.make_consensus_text(77, 120000, otus=5000)
    for kw in (dict(), dict(sim=80.0, bits=100.0), dict(every=True, sim=85.0)):
        got = ctx.megaclust(text, **kw)
        ref = op.oracle_megaclust(text, **kw)
        assert got == ref
    # size-independent properties: -c counts are never below the distinct-pair counts; totals add up
    pairs, ex, by = ctx.megaclust(text, sim=80.0)
    every, ex2, by2 = ctx.megaclust(text, sim=80.0, every=True)
    assert (ex, by) == (ex2, by2) and [s for s, _ in pairs] == [s for s, _ in every]
    assert all(a[1] <= b[1] for a, b in zip(pairs, every)) and sum(c for _, c in every) == ex - by


def test_degenerate_inputs(ctx):
    assert ctx.megaclust(b"") == ([], 0, 0)
    assert ctx.megaclust(b"# only\n#comments\n") == ([], 0, 0)
    assert ctx.megaclust(b"\n\n") == ([], 2, 2)
    line = b"q\ts\t99\t1\t1\t1\t1\t1\t1\t1\t1e-30\t300"
    assert ctx.megaclust(line) == ([(b"s", 1)], 1, 0)                       # no final newline
    assert ctx.megaclust(line + b"\n" + line + b"\n") == ([(b"s", 1)], 2, 0)
    assert ctx.megaclust(line + b"\n" + line + b"\n", every=True) == ([(b"s", 2)], 2, 0)


def test_cli_is_a_drop_in(tmp_path):
    text = (GOLD / "input.txt").read_bytes()
    want = json.loads((GOLD / "expected.json").read_text())
    (tmp_path / "in.txt").write_bytes(text)
    for name, case in CASES.items():
        r = subprocess.run([str(BIN / "megaclust2"), "-i", "in.txt", "-o", f"{name}.txt", *case["args"]], cwd=tmp_path,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == want[name]["stdout"]
        lines = (tmp_path / f"{name}.txt").read_text().split("\n")
        assert lines[0] == "OTU,times_hit" and sorted(l for l in lines[1:] if l) == want[name]["lines"]
    r = subprocess.run([str(BIN / "megaclust2"), "-i", "in.txt", "-o", "d.txt", "-d", ";"], cwd=tmp_path, capture_output=True, text=True)
    assert (tmp_path / "d.txt").read_text().startswith("OTU;times_hit\n")
    r = subprocess.run([str(BIN / "megaclust2"), "-i", "in.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.stdout.startswith("Must specify both an input and output filename\nUsage:\n")
    r = subprocess.run([str(BIN / "megaclust2"), "-i", "in.txt", "-o", "x.txt", "-s", "101"], cwd=tmp_path, capture_output=True, text=True)
    assert r.stdout.startswith("similarity threshold must be between 0 and 100\nUsage:\n")
    r = subprocess.run([str(BIN / "megaclust2"), "-i", "missing.txt", "-o", "x.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 2 and r.stderr.startswith("couldn't open infile")
