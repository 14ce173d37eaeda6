"""Megaclust: the C restatement against the live Perl script and against the committed
golden files; the Perl-number reader against Perl itself; the megaclustable CLI (plain C) against golden tables."""
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

GOLD = REPO / "tests" / "golden" / "megaclust"
BIN = REPO / "pangea-plus_b200" / "bin"

CASES = {
    "default": dict(args=[], kw={}),
    "pipeline": dict(args=["-b", "100", "-s", "80", "-e", "1e-20"], kw=dict(sim=80.0, bits=100.0)),     # README.md:176
    "every": dict(args=["-c", "1", "-s", "90"], kw=dict(sim=90.0, every=True)),
    "loose": dict(args=["-s", "0.0", "-e", "1", "-b", "0.0"], kw=dict(sim=0.0, ev=1.0, bits=0.0)),
}


def fmt(result, delim=b","):
    return sorted(s + delim + str(c).encode() for s, c in result)


def test_oracle_matches_golden():
    text = (GOLD / "input.txt").read_bytes()
    want = json.loads((GOLD / "expected.json").read_text())
    for name, case in CASES.items():
        res, ex, by = op.oracle_megaclust(text, **case["kw"])
        assert [l.decode() for l in fmt(res)] == want[name]["lines"], name
        assert want[name]["stdout"] == f"Run complete:\n{ex} hits examined\n{by} hits beyond thresholds and therefore not counted.\n"
        assert len({s for s, _ in res}) == len(res)               # one line per subject


@pytest.mark.skipif(not op.have_megaclust_reference(), reason="reference tree or perl absent")
def test_oracle_matches_live_script():
    for seed in (1, 2, 3):
        text = synth_mega.make_consensus_text(seed, 400, otus=60)
        for name, case in CASES.items():
            lines, header, stdout = op.real_megaclust(text, case["args"])
            res, ex, by = op.oracle_megaclust(text, **case["kw"])
            assert header == b"OTU,times_hit"
            assert fmt(res) == lines, (seed, name)
            assert stdout.decode() == f"Run complete:\n{ex} hits examined\n{by} hits beyond thresholds and therefore not counted.\n"


@pytest.mark.skipif(not op.have_megaclust_reference(), reason="reference tree or perl absent")
def test_number_reader_matches_perl():
    samples = ["98.63", "1e-20", "1E-21x", " 937", "+200.0", ".5e-20", "5.", ".", "", "abc", "0x10", "1_000", "inf", "-Infinity",
               "3e", "3e+", "12e3junk", "  -4.25e-3 ", "100", "0.0", "7e-81", "123456789012345678901234", "0.000000000000000000012"]
    script = "for (@ARGV) { my $v = $_ + 0; print(($v != $v) ? 'nan' : sprintf('%.17g', $v), \"\\n\") }"
    got = subprocess.run(["perl", "-e", script, "--", *samples], capture_output=True, text=True).stdout.split("\n")
    for s, g in zip(samples, got):
        v = op.oracle_number(s.encode())
        mine = "nan" if v != v else "%.17g" % v
        assert mine.lower() == g.lower(), (s, mine, g)


def test_megaclustable_cli_matches_golden(tmp_path):
    want = json.loads((GOLD / "tables.json").read_text())
    for name, spec in want.items():
        for fn, body in spec["files"].items():
            (tmp_path / fn).write_text(body)
        out = tmp_path / f"{name}.txt"
        r = subprocess.run([str(BIN / "megaclustable"), "-m", *spec["order"], "-t", spec["level"], "-o", out.name],
                           cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert out.read_text() == spec["table"], name
    r = subprocess.run([str(BIN / "megaclustable"), "-m", "x"], capture_output=True, text=True)
    assert r.stdout == "Please enter the correct parameters.\n"
