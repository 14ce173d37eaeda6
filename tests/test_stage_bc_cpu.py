"""CPU tests of the Stage B / Stage C oracles.

The golden files under tests/golden/ were produced by the reference's own code
(tests/golden/make_golden.py: tax_class compiled from Tax_class/ncbitc.c, the Perl scripts
NCBI-taxcollector-0.01.pl and Consensus_BLAST_SOAP_RDP-1.1.pl).  The C restatements in
oracle/ must reproduce them byte for byte; where the reference tree and perl are present
(the build container) they are additionally checked live on fresh seeds."""
import shutil
import subprocess
from pathlib import Path

import pytest

import oracle_pipeline as op
from pangea_b200 import synth_tax as st

GOLD = Path(__file__).parent / "golden"
needs_ref_binary = pytest.mark.skipif(not op.ref_tax_class().exists(), reason="oracle/_ref/tax_class not built")
needs_reference = pytest.mark.skipif(not op.have_reference(), reason="reference tree / perl not available here")


@needs_ref_binary
@pytest.mark.parametrize("case", ["tax_mini", "tax_synth"])
def test_taxcollector_oracle_matches_golden(case, tmp_path):
    op.ref_build_bins(GOLD / case, tmp_path / "bins")
    out = tmp_path / "out.txt"
    assert op.oracle_taxcollector(tmp_path / "bins", GOLD / case / "hits.txt", out) == 0
    assert out.read_bytes() == (GOLD / case / "hits_class.expected.txt").read_bytes()


@pytest.mark.parametrize("case,blast", [("tax_synth", "hits_class.expected.txt"), ("consensus_probes", "blast_class.txt")])
def test_consensus_oracle_matches_golden(case, blast, tmp_path):
    out = tmp_path / "out.txt"
    assert op.oracle_consensus(GOLD / case / blast, GOLD / case / "rdp.txt", out) == 0
    assert out.read_bytes() == (GOLD / case / "consensus.expected.txt").read_bytes()


def test_golden_documents_the_reference_quirks():
    mini = (GOLD / "tax_mini" / "hits_class.expected.txt").read_text().split("\n")
    assert mini[0].split("\t")[1] == "[0]Eukaryota;[9]Metazoa;[1]Chordata;[2]Mammalia;[4]Bovidae;[5]Bos;[6]Bos_taurus;"
    assert mini[1].split("\t")[1] == "[0]Bacteria;[5]uncultured_bacterium;[6]uncultured_bacterium;"       # genus back-fill
    assert mini[2].split("\t")[1] == "[0]Bacteria;[5]Strain_7_sp._X77;[6]Strain_7_sp._X77;"               # '6' branch wins over 7->9
    assert mini[2].split("\t")[-1] == "937"                                                                # " 937" collapses
    assert mini[3].split("\t")[1] == "[0]Unclassified;[6]viral_thing_5;"                                   # a '5' suppresses the back-fill
    assert mini[4].split("\t")[1] == "Unidentified(GI:9);"
    assert mini[6].split("\t")[1] == "Unidentified(GI:15);Unidentified(GI:16);"                            # the '6' rule hits the GI text too
    probes = (GOLD / "consensus_probes" / "consensus.expected.txt").read_text().split("\n")
    assert probes[1] == "#Matches found: 6"                       # quotes / digits sanitised on the RDP side only
    assert probes[5] == "#Matches found: 3"                       # unknown rank tokens match each other (undef eq undef)
    assert probes[7] == "#Matches found: 1"                       # a blank in a BLAST name shifts the pairing
    assert probes[10].startswith("R6\t[0]Bacteria;[5]x;[6]x;")    # stale $maxblastcount: string-order tie rules
    assert probes[14].startswith("R8\t[0]Bacteria;[1]Firmicutes;\t91.0")   # "9.5" lt "91.0", "100" lt "91.0" as strings


@needs_reference
@pytest.mark.parametrize("seed", [101, 202])
def test_oracles_against_the_live_reference(seed, tmp_path):
    tx = st.make_taxonomy(seed, 150, 3000)
    st.write_dumps(tx, str(tmp_path / "dumps"))
    lines, ids, per = st.make_blast_hits(seed + 1, tx, 25, max_hits=6)
    hits = tmp_path / "hits.txt"
    hits.write_text("\n".join(lines) + "\n")
    real, mine = tmp_path / "real_class.txt", tmp_path / "mine_class.txt"
    op.real_taxcollector(tmp_path / "dumps", hits, real)
    op.ref_build_bins(tmp_path / "dumps", tmp_path / "bins")
    assert op.oracle_taxcollector(tmp_path / "bins", hits, mine) == 0
    assert mine.read_bytes() == real.read_bytes()
    ids, by = op.group_lineages(real)
    rdp = tmp_path / "rdp.txt"
    rdp.write_text("\n".join(st.make_rdp_lines(seed + 2, ids, by)) + "\n")
    creal, cmine = tmp_path / "real_cons.txt", tmp_path / "mine_cons.txt"
    op.real_consensus(real, rdp, creal)
    assert op.oracle_consensus(real, rdp, cmine) == 0
    assert cmine.read_bytes() == creal.read_bytes()


@needs_reference
def test_consensus_oracle_skips_blast_ids_missing_from_rdp(tmp_path):
    """C8: BLAST ids absent from the RDP file are skipped ("not found:"); a leading RDP id with no
    BLAST line is passed over while $found is still undef."""
    blast = tmp_path / "b.txt"
    rdp = tmp_path / "r.txt"
    blast.write_text("A\t[0]Bacteria;\t90.0\nZ\t[0]Archaea;\t91.0\nB\t[0]Bacteria;[1]Firmicutes;\t92.0\nB\t[0]Bacteria;\t99.0\n")
    rdp.write_text("\n".join(i + "\t" * 5 + "Bacteria\tdomain\t1.0\tFirmicutes\tphylum\t0.5" for i in ["Q0", "A", "B"]) + "\n")
    real, mine = tmp_path / "real.txt", tmp_path / "mine.txt"
    op.real_consensus(blast, rdp, real)
    assert op.oracle_consensus(blast, rdp, mine) == 0
    assert mine.read_bytes() == real.read_bytes()
    assert real.read_text().count("#Matches found") == 2


def test_consensus_oracle_terminates_where_the_reference_loops(tmp_path):
    """documented deviation: an RDP id without BLAST lines makes the reference print 'not found:'
    forever; the restatement stops and says so."""
    blast = tmp_path / "b.txt"
    rdp = tmp_path / "r.txt"
    blast.write_text("A\t[0]Bacteria;\t90.0\n")
    rdp.write_text("A" + "\t" * 5 + "Bacteria\tdomain\t1.0\nB" + "\t" * 5 + "Bacteria\tdomain\t1.0\n")
    assert op.oracle_consensus(blast, rdp, tmp_path / "o.txt") == 1


def _masked_bins_equal(ref_dir: Path, my_dir: Path):
    """.bin files must be interchangeable.  The reference leaves bytes after a string's NUL (and
    struct padding) uninitialised / stale (ncbitc.c:511-557); those bytes are don't-care."""
    import numpy as np

    assert (ref_dir / "gi_taxid_nucl.dmp.bin").read_bytes() == (my_dir / "gi_taxid_nucl.dmp.bin").read_bytes()
    a = np.frombuffer((ref_dir / "nodes.dmp.bin").read_bytes(), np.uint8).reshape(-1, 28).copy()
    b = np.frombuffer((my_dir / "nodes.dmp.bin").read_bytes(), np.uint8).reshape(-1, 28).copy()
    assert a.shape == b.shape
    for x in (a, b):
        x[:, [15, 19, 27]] = 0                                   # struct padding
        e = x[:, 9:12]
        e[e[:, 0] == 0] = 0                                      # embl code: bytes after the first NUL
        e[(e[:, 0] != 0) & (e[:, 1] == 0), 2] = 0
    assert (a == b).all()
    ra, rb = (ref_dir / "names.dmp.bin").read_bytes(), (my_dir / "names.dmp.bin").read_bytes()
    assert len(ra) == len(rb) and ra[:4] == rb[:4]
    a = np.frombuffer(ra[4:], np.uint8).reshape(-1, 196).copy()
    b = np.frombuffer(rb[4:], np.uint8).reshape(-1, 196).copy()
    assert (a[:, :4] == b[:, :4]).all()
    for lo in (4, 68, 132):
        for x, y in zip(a[:, lo:lo + 64], b[:, lo:lo + 64]):
            sx, sy = bytes(x).split(b"\0")[0], bytes(y).split(b"\0")[0]
            assert sx == sy


@needs_ref_binary
@pytest.mark.parametrize("case", ["tax_mini", "tax_synth"])
def test_tax_build_writes_the_reference_bin_layout(case, tmp_path):
    """pg_tax_build == `tax_class -c` (B1-B3): pure file conversion, runs without a GPU."""
    import pangea_b200 as pg

    op.ref_build_bins(GOLD / case, tmp_path / "ref")
    mine = tmp_path / "mine"
    mine.mkdir()
    for f in ("nodes.dmp", "names.dmp", "gi_taxid_nucl.dmp"):
        shutil.copy(GOLD / case / f, mine / f)
    pg.tax_build(mine)
    _masked_bins_equal(tmp_path / "ref", mine)
    # and the reference's own reader accepts our files
    out = subprocess.run([str(op.ref_tax_class()), "-n", "2" if case == "tax_mini" else "1"], cwd=mine, capture_output=True, text=True)
    assert "scientific name" in out.stdout
