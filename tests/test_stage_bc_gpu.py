"""GPU parity tests for Stage B (lineage) and Stage C (consensus), through the C ABI and through
the drop-in executables.  Expected outputs: the golden files produced by the reference's own
Perl/C (tests/golden/make_golden.py), the reference's tax_class binary (oracle/_ref, compiled
from ncbitc.c), and the C restatements in oracle/ on larger seeded inputs."""
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_pipeline as op
import pangea_b200 as pg
from pangea_b200 import synth_tax as st

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
BIN = pg.PKG_DIR / "bin"


def build_bins(dump_dir: Path, dst: Path) -> Path:
    dst.mkdir(parents=True, exist_ok=True)
    for f in ("nodes.dmp", "names.dmp", "gi_taxid_nucl.dmp"):
        shutil.copy(Path(dump_dir) / f, dst / f)
    pg.tax_build(dst)                                      # our `tax_class -c`
    return dst


def run(cmd, cwd=None, check=True):
    r = subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True, timeout=600)
    if check:
        assert r.returncode == 0, f"{cmd}: rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return r


# ------------------------------------------------------------------ Stage B

def test_tax_class_cli_transcript(tmp_path):
    """`tax_class -s/-g/-t/-n` print exactly what the reference binary printed (B4)."""
    d = build_bins(GOLD / "tax_mini", tmp_path / "t")
    want = (GOLD / "tax_mini" / "tax_class_transcript.expected.txt").read_text()
    got = []
    for block in want.split("$ tax_class ")[1:]:
        head = block.split("\n", 1)[0]
        args = head.split("  [exit")[0].split()
        r = run([BIN / "tax_class"] + args, cwd=d, check=False)
        got.append("$ tax_class " + " ".join(args) + f"  [exit {r.returncode}]\n" + r.stdout)
    assert "".join(got) == want


@pytest.mark.skipif(not op.ref_tax_class().exists(), reason="oracle/_ref/tax_class not built")
def test_tax_class_cli_vs_reference_binary(tmp_path):
    """same answers as the reference's own tax_class on a seeded taxonomy, reading the SAME .bin files"""
    tx = st.make_taxonomy(31, 200, 4000)
    st.write_dumps(tx, str(tmp_path / "d"))
    op.ref_build_bins(tmp_path / "d", tmp_path / "ref")   # .bin written by the reference
    rng = np.random.default_rng(3)
    gis = [g for g, _ in tx["gi"]]
    taxids = [t for t, *_ in tx["nodes"]]
    gap = next(g for g in range(2, max(gis)) if g not in set(gis))
    for flag, pool, extra in (("-s", gis, [gap]), ("-g", gis, [gap]), ("-t", taxids, []), ("-n", taxids, [999999])):
        ids = [int(v) for v in rng.choice(pool, 12, replace=False)] + extra     # extra: unmapped gi / unknown taxid
        want = ""
        for v in ids:
            a = run([op.ref_tax_class(), flag, str(v)], cwd=tmp_path / "ref", check=False)
            assert a.returncode == 0
            want += a.stdout
        # our tool answers several requests from one load of the tables, in order
        args = [x for v in ids for x in (flag, str(v))]
        b = run([BIN / "tax_class"] + args, cwd=tmp_path / "ref", check=False)
        assert (b.returncode, b.stdout) == (0, want), flag


@pytest.mark.parametrize("case", ["tax_mini", "tax_synth"])
def test_taxcollector_cli_matches_golden(case, tmp_path):
    """drop-in for `perl NCBI-taxcollector-0.01.pl -f -o`, started like the README says: from the
    directory that holds Tax_class/"""
    build_bins(GOLD / case, tmp_path / "Tax_class")
    out = tmp_path / "out.txt"
    run([BIN / "taxcollector", "-f", GOLD / case / "hits.txt", "-o", out], cwd=tmp_path)
    assert out.read_bytes() == (GOLD / case / "hits_class.expected.txt").read_bytes()


def test_lineage_api_vs_oracle_large(ctx, tmp_path):
    tx = st.make_taxonomy(77, 3000, 200000)
    st.write_dumps(tx, str(tmp_path / "d"))
    d = build_bins(tmp_path / "d", tmp_path / "b")
    lines, ids, per = st.make_blast_hits(78, tx, 4000, max_hits=12, unmapped=0.05)
    hits = tmp_path / "hits.txt"
    hits.write_text("\n".join(lines) + "\n")
    want_file = tmp_path / "want.txt"
    assert op.oracle_taxcollector(d, hits, want_file) == 0
    want = [l.split("\t")[1].encode() for l in want_file.read_text().split("\n") if l]
    gi = np.array([int(l.split("|")[1]) for l in lines], np.int32)
    t = ctx.tax_load(d)
    got = t.lineage(gi)
    assert got == want
    leaf = t.leaf(gi)
    m = dict(tx["gi"])
    assert leaf.tolist() == [m.get(int(g), 0) for g in gi]
    assert t.lineage(np.zeros(0, np.int32)) == []                           # empty batch
    edge = t.lineage(np.array([0, -5, 2_000_000_000, -16, -17], np.int32))    # out of range: never a taxid
    assert edge == [b"Unidentified(GI:0);", b"Unidentified(GI:-5);", b"Unidentified(GI:2000000000);",
                    b"Unidentified(GI:-15);Unidentified(GI:-16);", b"Unidentified(GI:-19);"]   # the script's 6->5 / 7->9 rules
    t.free()
    # the CLI on the same input
    (tmp_path / "Tax_class").mkdir()
    for f in d.iterdir():
        shutil.copy(f, tmp_path / "Tax_class" / f.name)
    out = tmp_path / "cli.txt"
    run([BIN / "taxcollector", "-f", hits, "-o", out], cwd=tmp_path)
    assert out.read_bytes() == want_file.read_bytes()


def test_taxcollector_cli_stops_at_empty_line(tmp_path):
    build_bins(GOLD / "tax_mini", tmp_path / "Tax_class")
    src = (GOLD / "tax_mini" / "hits.txt").read_text().split("\n")
    (tmp_path / "h.txt").write_text("\n".join(src[:3]) + "\n\n" + "\n".join(src[3:]) + "\n")
    run([BIN / "taxcollector", "-f", tmp_path / "h.txt", "-o", tmp_path / "o.txt"], cwd=tmp_path)
    want = (GOLD / "tax_mini" / "hits_class.expected.txt").read_text().split("\n")[:3]
    assert (tmp_path / "o.txt").read_text() == "\n".join(want) + "\n"


# ------------------------------------------------------------------ Stage C

@pytest.mark.parametrize("case,blast", [("tax_synth", "hits_class.expected.txt"), ("consensus_probes", "blast_class.txt")])
def test_consensus_cli_matches_golden(case, blast, tmp_path):
    out = tmp_path / "out.txt"
    r = run([BIN / "consensus", "-b", GOLD / case / blast, "-r", GOLD / case / "rdp.txt", "-o", out])
    assert out.read_bytes() == (GOLD / case / "consensus.expected.txt").read_bytes()
    assert "Loading input files..." in r.stdout and "Done!" in r.stdout
    if case == "consensus_probes":
        assert "not found: X9\t R8" in r.stdout                             # C8: BLAST id absent from the RDP file


def test_consensus_cli_accepts_and_ignores_soap_file(tmp_path):
    soap = tmp_path / "soap.txt"
    soap.write_text("whatever\n")
    out = tmp_path / "out.txt"
    run([BIN / "consensus", "-b", GOLD / "tax_synth" / "hits_class.expected.txt", "-r", GOLD / "tax_synth" / "rdp.txt",
         "-s", soap, "-o", out])
    assert out.read_bytes() == (GOLD / "tax_synth" / "consensus.expected.txt").read_bytes()
    r = run([BIN / "consensus", "-b", GOLD / "tax_synth" / "hits_class.expected.txt"])          # C1: usage, exit 0
    assert r.stdout.startswith("Usage: perl Consensus-1.0.pl")


def test_consensus_vs_oracle_large(tmp_path):
    tx = st.make_taxonomy(55, 2000, 100000)
    st.write_dumps(tx, str(tmp_path / "d"))
    d = build_bins(tmp_path / "d", tmp_path / "b")
    lines, ids, per = st.make_blast_hits(56, tx, 5000, max_hits=20)
    hits = tmp_path / "hits.txt"
    hits.write_text("\n".join(lines) + "\n")
    cls = tmp_path / "class.txt"
    assert op.oracle_taxcollector(d, hits, cls) == 0
    ids, by = op.group_lineages(cls)
    rdp = tmp_path / "rdp.txt"
    rdp.write_text("\n".join(st.make_rdp_lines(57, ids, by, agree=0.6)) + "\n")
    want = tmp_path / "want.txt"
    assert op.oracle_consensus(cls, rdp, want) == 0
    out = tmp_path / "out.txt"
    run([BIN / "consensus", "-b", cls, "-r", rdp, "-o", out, "--quiet"])
    assert out.read_bytes() == want.read_bytes()
    # BLAST ids missing from the RDP file are skipped, RDP ids at the front without BLAST lines are passed over
    rl = rdp.read_text().split("\n")
    rdp2 = tmp_path / "rdp2.txt"
    rdp2.write_text("\n".join(["NOSUCH" + "\t" * 5 + "Bacteria\tdomain\t1.0"] + rl[:200] + rl[260:-1]) + "\n")
    assert op.oracle_consensus(cls, rdp2, want) == 0
    run([BIN / "consensus", "-b", cls, "-r", rdp2, "-o", out, "--quiet"])
    assert out.read_bytes() == want.read_bytes()


def test_consensus_soap_votes_is_the_script_on_the_merged_file(tmp_path):
    """--soap-votes (opt-in): SOAP hits vote by the rule 'what the reference script prints when every read's SOAP lines
    are inserted after the read's last BLAST line'.  The hits of a synthetic class file are split into a BLAST file
    and a SOAP file (some reads keep hits only in one of them, SOAP lists a few reads BLAST does not have, in a
    different place); the tool with the flag must print what the script's restatement prints for the merged file, and
    without the flag what it prints for the BLAST file alone."""
    tx = st.make_taxonomy(65, 1500, 80000)
    st.write_dumps(tx, str(tmp_path / "d"))
    d = build_bins(tmp_path / "d", tmp_path / "b")
    lines, ids, per = st.make_blast_hits(66, tx, 3000, max_hits=12)
    hits = tmp_path / "hits.txt"
    hits.write_text("\n".join(lines) + "\n")
    cls = tmp_path / "class.txt"
    assert op.oracle_taxcollector(d, hits, cls) == 0
    ids, by = op.group_lineages(cls)
    rdp = tmp_path / "rdp.txt"
    rdp.write_text("\n".join(st.make_rdp_lines(67, ids, by, agree=0.6)) + "\n")
    rng = np.random.default_rng(68)
    blast, soap_runs, merged = [], [], []
    cl = [l for l in cls.read_text().split("\n") if l]
    runs = {}
    for l in cl:
        runs.setdefault(l.split("\t")[0], []).append(l)
    for k, (rid, ls) in enumerate(runs.items()):
        to_soap = [l for l in ls if rng.random() < 0.4]
        to_blast = [l for l in ls if l not in to_soap]
        if not to_blast:                                  # every read keeps a BLAST line (reads without one do not vote)
            to_blast, to_soap = ls[:1], ls[1:]
        blast += to_blast
        if to_soap:
            soap_runs.append(to_soap)
        merged += to_blast + to_soap
    half = len(soap_runs) // 2                              # the SOAP file lists the reads in another order, and one read BLAST lacks
    soap = [l for r in soap_runs[half:] for l in r] + ["ZZ_only_soap\t[0]Bacteria;\t99.00\t250"] + [l for r in soap_runs[:half] for l in r]
    for name, ls in (("blast.txt", blast), ("soap.txt", soap), ("merged.txt", merged)):
        (tmp_path / name).write_text("\n".join(ls) + "\n")
    want = tmp_path / "want.txt"
    out = tmp_path / "out.txt"
    assert op.oracle_consensus(tmp_path / "merged.txt", rdp, want) == 0
    run([BIN / "consensus", "-b", tmp_path / "blast.txt", "-r", rdp, "-s", tmp_path / "soap.txt", "-o", out, "--quiet", "--soap-votes"])
    assert out.read_bytes() == want.read_bytes()
    assert op.oracle_consensus(tmp_path / "blast.txt", rdp, want) == 0
    run([BIN / "consensus", "-b", tmp_path / "blast.txt", "-r", rdp, "-s", tmp_path / "soap.txt", "-o", out, "--quiet"])
    assert out.read_bytes() == want.read_bytes()            # without the flag the SOAP file is opened and ignored (C7)


def test_consensus_stops_where_the_reference_loops_forever(tmp_path):
    b = tmp_path / "b.txt"
    r = tmp_path / "r.txt"
    b.write_text("A\t[0]Bacteria;\t90.0\n")
    r.write_text("A" + "\t" * 5 + "Bacteria\tdomain\t1.0\nB" + "\t" * 5 + "Bacteria\tdomain\t1.0\n")
    res = run([BIN / "consensus", "-b", b, "-r", r, "-o", tmp_path / "o.txt"], check=False)
    assert res.returncode == 2 and "forever" in res.stderr
    assert (tmp_path / "o.txt").read_text() == "A\t[0]Bacteria;\t90.0\n#Matches found: 1\n"


def test_consensus_api_directly(ctx):
    lin = [b"[0]Bacteria;[1]Firmicutes;", b"[0]Bacteria;", b"Unidentified(GI:9);", b""]
    pid = [b"91.0", b"99.5", b"70.00", b""]
    rdp = [b"Bacteria\tdomain\t1.0\t\"Firmicutes\"\tphylum\t0.8", b"Bacteria\tdomain\t1.0", b""]
    win, nm = ctx.consensus(np.array([0, 2, 3, 4]), lin, pid, rdp)
    assert win.tolist() == [0, 2, -1] and nm.tolist() == [2, 0, 0]          # -1: $tempresult is not touched
    win, nm = ctx.consensus(np.array([0, 0]), [], [], [b"x\tdomain\t1"])     # a read without hits
    assert win.tolist() == [-1] and nm.tolist() == [0]


# ------------------------------------------------------------------ Stage A executable + the README pipeline

def write_training_fasta(path, tr):
    names, anc = tr["node_names"], tr["anc"]
    with open(path, "w") as f:
        for i, g in enumerate(tr["genus"]):
            s = tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes().decode()
            f.write(f">T{i:06d}\t" + ";".join(names[n] for n in anc[g]) + "\n")
            for k in range(0, len(s), 80):
                f.write(s[k:k + 80] + "\n")


def test_rdp_classifier_cli_and_pipeline(tmp_path):
    import oracle_rdp as ora
    from pangea_b200 import synth

    tr = synth.synth16s(seed=61, seqs=150, genera=30, length=700)
    write_training_fasta(tmp_path / "train.fa", tr)
    model = tmp_path / "m.pgm"
    run([BIN / "rdp_classifier", "--train", tmp_path / "train.fa", "-t", model])
    data, off, src = synth.synth_reads(62, tr, 80, paired=True)
    reads = [data[off[i]:off[i + 1]].tobytes() for i in range(80)] + [b"ACGT" * 5]
    with open(tmp_path / "q.fa", "w") as f:
        for i, r in enumerate(reads):
            f.write(f">r{i:07d}:AB some description\n{r.decode()}\n")
    outs = {}
    for fmt in ("allrank", "fixrank", "pangea"):
        res = run([BIN / "rdp_classifier", "-q", tmp_path / "q.fa", "-o", tmp_path / f"{fmt}.txt", "-t", model, "-f", fmt])
        outs[fmt] = (tmp_path / f"{fmt}.txt").read_text().split("\n")[:-1]
        assert "ShortSequenceException" in res.stdout and "r0000080:AB" in res.stdout
        assert len(outs[fmt]) == 80                                          # the short read is left out of the file
    # expected lines from the oracle; the CLI numbers genera by first appearance in the training file
    first = {}
    for g in tr["genus"]:
        first.setdefault(int(g), len(first))
    order = sorted(first, key=first.get)
    genus_cli = np.array([first[int(g)] for g in tr["genus"]], np.int32)
    om = ora.Model(tr["data"], tr["off"], genus_cli, tr["G"])
    d2, o2 = pg.pack_sequences(reads[:80])
    ref = om.classify(d2, o2)
    anc_cli = tr["anc"][order]
    votes = om.votes(ref, anc_cli)
    ranks = ["rootrank", "domain", "phylum", "class", "order", "family", "genus"]

    def conf(v):
        return "1.0" if v == 100 else "0.0" if v == 0 else (f"0.{v // 10}" if v % 10 == 0 else f"0.{v:02d}")

    for i in range(80):
        lineage = anc_cli[ref[i]["genus"]]
        trip = [(tr["node_names"][n], ranks[d], conf(votes[i][d])) for d, n in enumerate(lineage)]
        want_all = f"r{i:07d}:AB\t" + ("-" if ref[i]["reversed"] else "") + "".join(f"\t{a}\t{b}\t{c}" for a, b, c in trip)
        assert outs["allrank"][i] == want_all
        assert outs["pangea"][i] == f"r{i:07d}:AB" + "\t" * 5 + "\t".join(f"{a}\t{b}\t{c}" for a, b, c in trip[1:])
        assert outs["fixrank"][i] == f"r{i:07d}:AB\t" + ("-" if ref[i]["reversed"] else "") + "".join(f"\t{a}\t{b}\t{c}" for a, b, c in trip[1:])
    om.free()
    # README pipeline: the 5-TAB file feeds the consensus stage; BLAST lineages reuse the RDP names
    blast = []
    for i in range(80):
        names = [tr["node_names"][n] for n in anc_cli[ref[i]["genus"]]][1:]
        blast.append(f"r{i:07d}:AB\t" + "".join(f"[{k}]{n};" for k, n in enumerate(names)) + "[6]" + names[-1] + "_sp.;\t99.00\t250")
        blast.append(f"r{i:07d}:AB\t[0]{names[0]};[5]uncultured;[6]uncultured;\t88.10\t250")
    (tmp_path / "blast.txt").write_text("\n".join(blast) + "\n")
    run([BIN / "consensus", "-b", tmp_path / "blast.txt", "-r", tmp_path / "pangea.txt", "-o", tmp_path / "cons.txt"])
    got = (tmp_path / "cons.txt").read_text().split("\n")[:-1]
    assert len(got) == 160 and all(l.startswith("#Matches found: ") for l in got[1::2])
    want = tmp_path / "cons_want.txt"
    assert op.oracle_consensus(tmp_path / "blast.txt", tmp_path / "pangea.txt", want) == 0
    assert (tmp_path / "cons.txt").read_bytes() == want.read_bytes()
    # the stock allrank layout is NOT parsed by Consensus-1.1: empty output there too
    run([BIN / "consensus", "-b", tmp_path / "blast.txt", "-r", tmp_path / "allrank.txt", "-o", tmp_path / "cons2.txt"], check=False)
    assert op.oracle_consensus(tmp_path / "blast.txt", tmp_path / "allrank.txt", want) in (0, 1)
    assert (tmp_path / "cons2.txt").read_bytes() == want.read_bytes() == b""


def test_rdp_classifier_cli_genus_token_mode(tmp_path):
    """headers like the reference's rdp_download files: genus = 2nd token, flat taxonomy"""
    fa = GOLD / "rdp_373_subset.fa"
    model = tmp_path / "m.pgm"
    run([BIN / "rdp_classifier", "--train", fa, "-t", model, "--genus-token", "2"])
    run([BIN / "rdp_classifier", "-q", fa, "-o", tmp_path / "o.txt", "-t", model])
    from pangea_b200 import synth

    ids, hdr, seqs = synth.read_fasta(fa)
    lines = (tmp_path / "o.txt").read_text().split("\n")[:-1]
    assert len(lines) == len(ids)
    hit = sum(1 for l, h in zip(lines, hdr) if l.split("\t")[5] == h.split()[1])
    assert hit >= 0.9 * len(ids)                                             # self-classification
    assert all(l.split("\t")[2:5] == ["Root", "rootrank", "1.0"] for l in lines)


def test_rdp_classifier_cli_multi_gpu_equals_single(tmp_path):
    """rdp_classifier --gpus N: the model counts are replicated from device 0 (ncclBroadcast between distinct GPUs, peer
    copies when the box has one GPU and it is listed twice), the query file is cut into record-aligned pieces that the
    device contexts take in turn, and the lines come out in input order: the file must be byte-identical to the
    one-GPU file whatever the piece size, and the records must be the oracle's."""
    import torch

    import oracle_rdp as ora
    from pangea_b200 import synth

    tr = synth.synth16s(seed=71, seqs=400, genera=140, length=800)
    write_training_fasta(tmp_path / "train.fa", tr)
    model = tmp_path / "m.pgm"
    run([BIN / "rdp_classifier", "--train", tmp_path / "train.fa", "-t", model])
    data, off, src = synth.synth_reads(72, tr, 5000, paired=True)
    with open(tmp_path / "q.fa", "w") as f:
        for i in range(5000):
            f.write(f">r{i:07d}:AB\n{data[off[i]:off[i + 1]].tobytes().decode()}\n")
            if i % 997 == 0:
                f.write(f">short{i}\nACGTACGT\n")
    one = tmp_path / "one.txt"
    r1 = run([BIN / "rdp_classifier", "-q", tmp_path / "q.fa", "-o", one, "-t", model, "--contexts-per-gpu", "1"])
    ndev = torch.cuda.device_count()
    devs = "0,1" if ndev >= 2 else "0,0"
    env = dict(os.environ, PG_CLI_PIECE_BYTES="300000")          # ~430 reads per piece: a dozen pieces in flight
    for extra in (["--gpus", "2", "--devices", devs], ["--gpus", "2", "--devices", devs, "--contexts-per-gpu", "1", "--format-threads", "1"],
                  ["--gpus", "1"]):
        out = tmp_path / "multi.txt"
        r2 = subprocess.run([str(BIN / "rdp_classifier"), "-q", str(tmp_path / "q.fa"), "-o", str(out), "-t", str(model)] + extra,
                            capture_output=True, text=True, env=env, timeout=600)
        assert r2.returncode == 0, r2.stderr
        if out.read_bytes() != one.read_bytes():             # say where, not just that
            a, b = one.read_text().split("\n"), out.read_text().split("\n")
            bad = [i for i in range(min(len(a), len(b))) if a[i] != b[i]]
            raise AssertionError(f"{extra}: {len(a)} vs {len(b)} lines, {len(bad)} differ, first {bad[:5]}: "
                                 + " | ".join(f"{a[i][:150]} <> {b[i][:150]}" for i in bad[:2]) + f" stderr: {r2.stderr[-300:]}")
        assert r2.stdout == r1.stdout and r2.stdout.count("ShortSequenceException") == 6
    # the lines carry the oracle's assignments
    first = {}
    for g in tr["genus"]:
        first.setdefault(int(g), len(first))
    genus_cli = np.array([first[int(g)] for g in tr["genus"]], np.int32)
    om = ora.Model(tr["data"], tr["off"], genus_cli, tr["G"])
    ref = om.classify(data[: off[300]], off[:301])
    order = sorted(first, key=first.get)
    lines = one.read_text().split("\n")[:-1]
    assert len(lines) == 5000
    for i in range(300):
        cells = lines[i].split("\t")
        assert cells[0] == f"r{i:07d}:AB" and cells[1] == ("-" if ref[i]["reversed"] else "")
        assert cells[-3] == tr["node_names"][tr["anc"][order[ref[i]["genus"]]][-1]]
    om.free()


def test_rdp_classifier_cli_stock_trainset_files(tmp_path):
    """--export-rdp writes the five files of an RDP trainset directory (layouts per appendix B,
    unverified against the jar); classifying from rRNAClassifier.properties must give, byte for byte, the lines the
    .pgm model gives -- in every format -- and -f db carries trainset number, taxid and confidence per rank."""
    from pangea_b200 import synth

    tr = synth.synth16s(seed=81, seqs=260, genera=70, length=700)
    write_training_fasta(tmp_path / "train.fa", tr)
    model = tmp_path / "m.pgm"
    run([BIN / "rdp_classifier", "--train", tmp_path / "train.fa", "-t", model])
    ts = tmp_path / "trainset"
    ts.mkdir()
    r = run([BIN / "rdp_classifier", "--export-rdp", ts, "-t", model])
    assert "exported 70 genera" in r.stdout
    for fn in ("rRNAClassifier.properties", "bergeyTrainingTree.xml", "logWordPrior.txt", "wordConditionalProbIndexArr.txt",
               "genus_wordConditionalProbList.txt"):
        assert (ts / fn).stat().st_size > 0
    tree = (ts / "bergeyTrainingTree.xml").read_text().split("\n")
    assert tree[0].startswith("<trainsetNo>") and "<file>bergeyTrainingTree</file>" in tree[0]
    assert tree[1].startswith('<TreeNode name="Root" taxid="0" rank="rootrank" parentTaxid="-1" leaveCount="260" genusIndex="-1">')
    assert sum('genusIndex="-1"' not in l for l in tree[1:] if l) == 70
    idx_lines = (ts / "wordConditionalProbIndexArr.txt").read_text().split("\n")
    assert len([l for l in idx_lines if l]) == 1 + 65537
    data, off, src = synth.synth_reads(82, tr, 600, paired=True)
    with open(tmp_path / "q.fa", "w") as f:
        for i in range(600):
            f.write(f">q{i:05d}\n{data[off[i]:off[i + 1]].tobytes().decode()}\n")
        f.write(">tiny\nACGT\n")
    for fmt in ("allrank", "fixrank", "pangea", "db"):
        a, b = tmp_path / f"a_{fmt}.txt", tmp_path / f"b_{fmt}.txt"
        ra = run([BIN / "rdp_classifier", "-q", tmp_path / "q.fa", "-o", a, "-t", model, "-f", fmt])
        rb = run([BIN / "rdp_classifier", "-q", tmp_path / "q.fa", "-o", b, "-t", ts / "rRNAClassifier.properties", "-f", fmt])
        assert a.read_bytes() == b.read_bytes() and a.stat().st_size > 0, fmt
        assert ra.stdout == rb.stdout and "ShortSequenceException" in ra.stdout
    db = (tmp_path / "a_db.txt").read_text().split("\n")[:-1]
    allr = (tmp_path / "a_allrank.txt").read_text().split("\n")[:-1]
    assert len(db) == 7 * 600                                # Root .. genus, one line each
    cells = allr[0].split("\t")
    for k in range(7):
        qid, tset, taxid, conf = db[k].split("\t")
        assert qid == "q00000" and tset == "0" and conf == cells[4 + 3 * k]
    # two devices from trainset files: every device reads the files itself (tables carry no counts to broadcast)
    import torch

    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    c = tmp_path / "c.txt"
    run([BIN / "rdp_classifier", "-q", tmp_path / "q.fa", "-o", c, "-t", ts / "rRNAClassifier.properties", "--gpus", "2", "--devices", devs])
    assert c.read_bytes() == (tmp_path / "a_allrank.txt").read_bytes()
    # -g picks the default model from the environment when -t is absent
    env = dict(os.environ, PANGEA_RDP_MODEL_FUNGALLSU=str(model))
    r = subprocess.run([str(BIN / "rdp_classifier"), "-q", str(tmp_path / "q.fa"), "-o", str(c), "-g", "fungallsu"], capture_output=True,
                       text=True, env=env, timeout=300)
    assert r.returncode == 0 and c.read_bytes() == (tmp_path / "a_allrank.txt").read_bytes()
    r = subprocess.run([str(BIN / "rdp_classifier"), "-q", str(tmp_path / "q.fa"), "-o", str(c), "-g", "18s"], capture_output=True, text=True)
    assert r.returncode != 0
