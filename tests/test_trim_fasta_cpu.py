"""The FASTA paths of trim2 (-j join, -q quality trim; Trim/trim2.4.pl join_fasta :300-383, parse_fasta :384-466) are
host-side line cursors in bin/trim2 (no GPU work: two files read in lockstep).  They are pinned to golden files made by
the reference script (tests/golden/trim_fasta/, made by make_golden() below) and, where the reference tree and perl
exist, to the live script on seeded inputs with every awkward line shape the script has an accident for."""
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

import oracle_pipeline as op
import pangea_b200 as pg

GOLD = Path(__file__).parent / "golden" / "trim_fasta"
BIN = pg.PKG_DIR / "bin" / "trim2"
REF = Path("/root/reference/Trim/trim2.4.pl")


def make_fasta_pair(seed, nrec, trailing_newline=True, crlf=False):
    rng = np.random.default_rng(seed)

    def rec(tag, i):
        hdr = f">{tag}{i:03d}" + (" some description" if rng.random() < 0.6 else "") + (" has > inside" if rng.random() < 0.1 else "")
        nl = int(rng.integers(1, 4))
        lines = ["".join(rng.choice(list("ACGTN"), int(rng.integers(1, 70)))) for _ in range(nl)]
        if rng.random() < 0.1:
            lines.insert(int(rng.integers(0, len(lines) + 1)), "")           # an empty line inside the record
        return [hdr] + lines

    out = []
    for tag, n in (("a", nrec), ("b", nrec + int(rng.integers(-1, 2)))):       # the files may differ by a record
        ls = []
        for i in range(max(n, 1)):
            ls += rec(tag, i)
        text = ("\r\n" if crlf else "\n").join(ls) + (("\r\n" if crlf else "\n") if trailing_newline else "")
        out.append(text)
    return out


def make_qual(seed, fasta_text):
    rng = np.random.default_rng(seed)
    out = []
    for l in fasta_text.split("\n"):
        if l.startswith(">"):
            out.append(l)
        elif l == "" and rng.random() < 0.5:
            out.append("")
        else:
            q = [str(int(rng.integers(0, 41))) for _ in range(len(l))]
            if q and rng.random() < 0.2:
                q[int(rng.integers(0, len(q)))] = ""                           # two blanks in a row: an empty field
            if q and rng.random() < 0.1:
                q[0] = "-3"
            out.append(" ".join(q))
    return "\n".join(out)


def run_mine(args, cwd):
    r = subprocess.run([str(BIN)] + args, cwd=cwd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout


def run_ref(args, cwd):
    r = subprocess.run(["perl", str(REF)] + args, cwd=cwd, capture_output=True, text=True, timeout=120)
    return r.stdout


CASES = [("join_g5", ["-j", "-g", "5"]), ("join_nogap", ["-j"]), ("join_g0", ["-j", "-g", "0"]), ("qual", None)]


def make_golden():
    """regenerates tests/golden/trim_fasta/ from the reference script (run in the build container)"""
    GOLD.mkdir(parents=True, exist_ok=True)
    a, b = make_fasta_pair(11, 12)
    (GOLD / "a.fa").write_text(a)
    (GOLD / "b.fa").write_text(b)
    (GOLD / "a.qual").write_text(make_qual(12, a))
    for name, args in CASES:
        with tempfile.TemporaryDirectory() as wd:
            if args is None:
                out = run_ref(["-a", str(GOLD / "a.fa"), "-q", str(GOLD / "a.qual")], wd)
                (GOLD / f"{name}.expected.stdout").write_text(out.replace(str(GOLD) + "/", ""))
            else:
                run_ref(["-a", str(GOLD / "a.fa"), "-b", str(GOLD / "b.fa")] + args, wd)
                (GOLD / f"{name}.expected.fasta").write_bytes((Path(wd) / "output_files/trim2/a.fa_runblast.fasta").read_bytes())


@pytest.mark.parametrize("name,args", CASES)
def test_fasta_paths_match_golden(name, args, tmp_path):
    if args is None:
        out = run_mine(["-a", str(GOLD / "a.fa"), "-q", str(GOLD / "a.qual")], tmp_path)
        assert out.replace(str(GOLD) + "/", "") == (GOLD / f"{name}.expected.stdout").read_text()
        assert (tmp_path / "output_files/trim2/a.fa_runblast.fasta").read_bytes() == b""      # the script never writes it
    else:
        out = run_mine(["-a", str(GOLD / "a.fa"), "-b", str(GOLD / "b.fa")] + args, tmp_path)
        assert out == ""                                                                      # join_fasta exits: no "Trimming complete."
        assert (tmp_path / "output_files/trim2/a.fa_runblast.fasta").read_bytes() == (GOLD / f"{name}.expected.fasta").read_bytes()


def test_fasta_path_messages(tmp_path):
    out = run_mine(["-a", str(GOLD / "a.fa"), "-j"], tmp_path)
    assert out == "Error. Input is -j for joining ends, but you did not provided both sequence a and b with -a and -b options.\n\n"
    out = run_mine(["-a", str(GOLD / "a.fa")], tmp_path)
    assert out == "Error: Please, specify the FASTA quality file with -q option.\n"
    out = run_mine(["-a", str(GOLD / "a.fa"), "-q", "nosuch.qual"], tmp_path)
    assert out == "nosuch.qual\nError: Unable to open nosuch.qual required for FASTA file triming.\n"


@pytest.mark.skipif(not (REF.exists() and op.have_reference()), reason="reference tree / perl not available here")
@pytest.mark.parametrize("seed", range(6))
def test_fasta_paths_against_the_live_script(seed, tmp_path):
    a, b = make_fasta_pair(100 + seed, 3 + 4 * seed, trailing_newline=seed % 2 == 0, crlf=seed == 5)
    (tmp_path / "a.fa").write_text(a)
    (tmp_path / "b.fa").write_text(b)
    (tmp_path / "a.qual").write_text(make_qual(200 + seed, a) + ("\n" if seed % 3 else ""))
    for args in (["-j", "-g", "7"], ["-j"], ["-b", "IGNORED", "-j"]):
        mine, ref = tmp_path / "mine", tmp_path / "ref"
        for d in (mine, ref):
            d.mkdir(exist_ok=True)
        base = ["-a", str(tmp_path / "a.fa")] + (["-b", str(tmp_path / "b.fa")] if "IGNORED" not in args else [])
        extra = [x for x in args if x not in ("-b", "IGNORED")]
        om, orf = run_mine(base + extra, mine), run_ref(base + extra, ref)
        assert om == orf
        assert (mine / "output_files/trim2/a.fa_runblast.fasta").read_bytes() == (ref / "output_files/trim2/a.fa_runblast.fasta").read_bytes()
    mine, ref = tmp_path / "mq", tmp_path / "rq"
    for d in (mine, ref):
        d.mkdir(exist_ok=True)
    args = ["-a", str(tmp_path / "a.fa"), "-q", str(tmp_path / "a.qual")]
    assert run_mine(args, mine) == run_ref(args, ref)


if __name__ == "__main__":
    make_golden()
