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


@pytest.mark.skipif(not op.have_megaclust_reference(), reason="reference tree or perl absent")
def test_product_number_reader_matches_perl(tmp_path):
    """csrc/pg_perlnum.h is what the megaclust2 executable (thresholds from argv) and the CUDA kernel (fields of the
    input) both use; its host build must read numbers like Perl does: to the last bit on ordinary decimals (up to
    19 digits, |exponent| <= 22: one exact table entry, one IEEE operation), within 4 ulp beyond that (the header
    says so: decimal strings that close to a threshold are outside the defined behaviour)."""
    import ctypes

    src = tmp_path / "pn.c"
    src.write_text('#include "pg_perlnum.h"\ndouble pn(const char *s, int n) { return pg_perl_number(s, n); }\n')
    so = tmp_path / "pn.so"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(REPO / "pangea-plus_b200" / "csrc"),
                    "-o", str(so), str(src)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.pn.restype = ctypes.c_double
    lib.pn.argtypes = [ctypes.c_char_p, ctypes.c_int]
    samples = ["98.63", "1e-20", "1E-21x", " 937", "+200.0", ".5e-20", "5.", ".", "", "abc", "0x10", "1_000", "inf", "-Infinity",
               "3e", "3e+", "12e3junk", "  -4.25e-3 ", "100", "0.0", "7e-81", "95", "94.99", "80", "1e-5", "2e-45", "250",
               "99.50", "100.00", "0.001", "1e-180", "3.5e-7"]
    script = "for (@ARGV) { my $v = $_ + 0; print(($v != $v) ? 'nan' : sprintf('%.17g', $v), \"\\n\") }"
    got = subprocess.run(["perl", "-e", script, "--", *samples], capture_output=True, text=True).stdout.split("\n")
    far = {"7e-81", "1e-180", "2e-45"}                      # scaled by more than 10^22: several roundings
    for s, g in zip(samples, got):
        v = lib.pn(s.encode(), len(s))
        mine = "nan" if v != v else "%.17g" % v
        if s in far:
            assert abs(v - float(g)) <= 4 * abs(float(g)) * 2.0 ** -52, (s, mine, g)
        else:
            assert mine.lower() == g.lower(), (s, mine, g)
