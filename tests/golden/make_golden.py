#!/usr/bin/env python
"""Regenerates the Stage B / Stage C golden fixtures by running the REFERENCE ITSELF
(/root/reference: tax_class compiled from Tax_class/ncbitc.c by oracle/Makefile `ref`, the
Perl scripts NCBI-taxcollector-0.01.pl and Consensus_BLAST_SOAP_RDP-1.1.pl) on seeded
synthetic inputs and on the probe cases of SURVEY.md appendix A.  Run in the build
container (needs perl and /root/reference); the outputs are committed.

    python tests/golden/make_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
from pangea_b200 import synth_tax as st  # noqa: E402


def run_reference_taxcollector(dump_dir: Path, hits: Path, out: Path):
    """the README recipe: cd $PANGEAWD; tax_class -c inside Tax_class/; perl NCBI-taxcollector... -f -o"""
    with tempfile.TemporaryDirectory() as wd:
        wd = Path(wd)
        tc = wd / "Tax_class"
        tc.mkdir()
        for f in ("nodes.dmp", "names.dmp", "gi_taxid_nucl.dmp"):
            shutil.copy(dump_dir / f, tc / f)
        shutil.copy(REPO / "oracle" / "_ref" / "tax_class", tc / "tax_class")
        subprocess.run(["./tax_class", "-c"], cwd=tc, check=True)
        subprocess.run(["perl", str(REF / "Tax_class" / "NCBI-taxcollector-0.01.pl"), "-f", str(hits), "-o", str(out)],
                       cwd=wd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run_reference_consensus(blast: Path, rdp: Path, out: Path, timeout=120):
    subprocess.run(["perl", str(REF / "Consensus" / "Consensus_BLAST_SOAP_RDP-1.1.pl"), "-b", str(blast), "-r", str(rdp),
                    "-o", str(out)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)


def group_lineages(class_file: Path):
    ids, by = [], []
    for l in class_file.read_text().split("\n"):
        if not l:
            continue
        f = l.split("\t")
        if not ids or ids[-1] != f[0]:
            ids.append(f[0])
            by.append([])
        by[-1].append(f[1] if len(f) > 1 else "")
    return ids, by


# the mini taxonomy with exactly the probed nodes
MINI_NODES = [(1, 1, "no rank", ""), (2, 131567, "superkingdom", ""), (433, 204441, "family", ""),
              (1224, 2, "phylum", ""), (9903, 9895, "genus", ""), (9895, 40674, "family", ""), (9913, 9903, "species", "BT"),
              (28211, 1224, "class", ""), (33208, 2759, "kingdom", ""), (2759, 131567, "superkingdom", ""),
              (7711, 33208, "phylum", ""), (40674, 7711, "class", ""), (77133, 2, "species", ""),
              (125216, 433, "genus", ""), (131567, 1, "no rank", ""), (204441, 28211, "order", ""),
              (887062, 125216, "species", "RS"), (900001, 2, "no rank", ""), (900002, 900001, "species", ""),
              (900003, 1, "no rank", ""), (900004, 900003, "species", "")]
MINI_NAMES = [(1, "all", "", "synonym"), (1, "root", "", "scientific name"), (2, "Bacteria", "Bacteria <prokaryote>", "scientific name"),
              (2, "eubacteria", "", "genbank common name"), (433, "Methylocystaceae", "", "scientific name"),
              (1224, "Proteobacteria", "", "scientific name"), (2759, "Eukaryota", "", "scientific name"),
              (7711, "Chordata", "", "scientific name"), (9895, "Bovidae", "", "scientific name"), (9903, "Bos", "", "scientific name"),
              (9913, "Bos bovis", "", "synonym"), (9913, "Bos taurus", "", "scientific name"), (9913, "cow", "", "common name"),
              (28211, "Alphaproteobacteria", "", "scientific name"), (33208, "Metazoa", "", "scientific name"),
              (40674, "Mammalia", "", "scientific name"), (77133, "uncultured bacterium", "", "scientific name"),
              (125216, "Methylocystis", "", "scientific name"), (131567, "cellular organisms", "", "scientific name"),
              (204441, "Rhizobiales", "", "scientific name"), (887062, "Methylocystis sp. SC2", "", "scientific name"),
              (900001, "unclassified Bacteria", "", "scientific name"), (900002, "Strain 7 sp. X77", "", "scientific name"),
              (900003, "other sequences", "", "scientific name"), (900004, "viral thing 5", "", "scientific name"),
              (900005, "never printed last record", "", "scientific name")]
MINI_GI = [(2, 9913), (5, 887062), (7, 77133), (11, 900002), (12, 900004), (16, 0), (20, 2)]
MINI_HITS = ["Q1\tgi|2|gb|A|\t99.0\t100\t0\t0\t1\t100\t1\t100\t0.0\t200", "Q2\tgi|7|gb|A|\t99.0\t100\t0\t0\t1\t100\t1\t100\t0.0\t200",
             "Q3\tgi|11|gb|A|\t98.5\t100\t0\t0\t1\t100\t1\t100\t0.0\t 937", "Q4\tgi|12|gb|A|\t97.0\t100\t0\t0\t1\t100\t1\t100\t0.0\t200",
             "S0000002\tgi|9|gb|A|\t70.00\t100\t0\t0\t1\t100\t1\t100\t1e-5\t50", "Q5\tgi|5|gb|A|\t91.25\t100\t0\t0\t1\t100\t1\t100\t0.0\t200",
             "Q6\tgi|16|gb|A|\t88.00\t100\t0\t0\t1\t100\t1\t100\t0.0\t200", "Q7\tgi|20|gb|A|\t88.00\t100\t0\t0\t1\t100\t1\t100\t0.0\t200"]

# consensus probe cases (BLAST-class lines / RDP lines)
T = "\t"
A3_BLAST = [
    "R1\t[0]Bacteria;[1]Firmicutes;[2]Bacilli;[3]Bacillales;[4]Bacillaceae;[5]Bacillus;[6]Bacillus_sp._8A18S6;\t92.61\t100",
    "R2\t[0]Bacteria;[1]Actinobacteria;[2]Actinobacteria;[3]Actinomycetales;[4]Micrococcaceae;[5]Arthrobacter;[6]Arthrobacter_sp.;\t81.87\t100",
    "R2\t[0]Bacteria;[5]uncultured_bacterium;[6]uncultured_bacterium;\t78.63\t100",
    "R2\t[0]Bacteria;[1]Actinobacteria;[2]Actinobacteria;[3]Actinomycetales;[4]Micrococcaceae;[5]Kocuria;[6]Kocuria_rosea;\t100.00\t100",
    "R3\t[0]Eukaryota;[9]Metazoa;[1]Chordata;\t99.0\t10",
    "R4\t[0]Bacteria;[1]Candidatus Foo;[2]Bar;\t99.0\t10",
    "R5\t[0]Bacteria;[1]Actinobacteria;[2]Actinobacteria;\t99.0\t10",
    "R6\t[0]Archaea;[1]A;[2]B;[3]C;[4]D;[5]E;[6]F;\t99.00\t10",
    "R6\t[0]Bacteria;[5]x;[6]x;\t80.00\t10",
    "R6\t[0]Bacteria;[1]A;[2]B;[3]C;[4]D;[5]E;[6]F;\t70.00\t10",
    "R7\tUnidentified(GI:9);\t70.00\t10",
    "X9\t[0]Bacteria;\t50.0\t10",
    "R8\t[0]Bacteria;[1]Firmicutes;\t91.0\t10",
    "R8\t[0]Bacteria;[1]Firmicutes;\t9.5\t10",
    "R8\t[0]Bacteria;[1]Firmicutes;\t100\t10",
]
A3_RDP = [
    "R1" + T * 5 + T.join(["Bacteria", "domain", "1.0", '"Firmicutes"', "phylum", "1.0", '"Bacilli"', "class", "1.0", "Bacillales", "order", "1.0",
                           "Bacillaceae 1", "family", "0.98", "Bacillus", "genus", "0.9"]),
    "R2" + T * 5 + T.join(["Bacteria", "domain", "1.0", '"Actinobacteria"', "phylum", "1.0"]),
    "R3" + T * 5 + T.join(["Eukaryota", "domain", "1.0", "Metazoa", "subkingdomX", "1.0", "Chordata", "phylum", "0.5"]),
    "R4" + T * 5 + T.join(["Bacteria", "domain", "1.0", "Foo", "phylum", "1.0", "Bar", "class", "1.0"]),
    "R5" + T * 5 + T.join(["Bacteria", "domain", "1.0", '"Actinobacteria"', "phylum", "1.0", "Actinobacteria", "class", "1.0", "Actinobacteridae", "subclass", "1.0"]),
    "R6" + T * 5 + T.join(["Bacteria", "domain", "1.0"]),
    "R7" + T * 5 + T.join(["Bacteria", "domain", "1.0"]),
    "R8" + T * 5 + T.join(["Bacteria", "domain", "1.0", "Firmicutes", "phylum", "0.7"]),
]


def main():
    subprocess.run(["make", "-C", str(REPO / "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    # ---- mini taxonomy (appendix A.1/A.2)
    d = HERE / "tax_mini"
    d.mkdir(exist_ok=True)
    st.write_dumps(dict(nodes=sorted(MINI_NODES), names=MINI_NAMES, gi=MINI_GI), str(d))
    (d / "hits.txt").write_text("\n".join(MINI_HITS) + "\n")
    run_reference_taxcollector(d, d / "hits.txt", d / "hits_class.expected.txt")
    # tax_class -s / -t / -n transcripts from the reference binary
    with tempfile.TemporaryDirectory() as wd:
        for f in ("nodes.dmp", "names.dmp", "gi_taxid_nucl.dmp"):
            shutil.copy(d / f, Path(wd) / f)
        tcb = str(REPO / "oracle" / "_ref" / "tax_class")
        subprocess.run([tcb, "-c"], cwd=wd, check=True)
        lines = []
        for args in (["-s", "5"], ["-s", "2"], ["-s", "1"], ["-s", "16"], ["-g", "5"], ["-g", "16"], ["-t", "3"], ["-t", "9913"],
                     ["-t", "131567"], ["-n", "9913"], ["-n", "7"], ["-n", "2"], ["-n", "887062"], ["-s", "7"], ["-s", "20"]):
            r = subprocess.run([tcb] + args, cwd=wd, capture_output=True, text=True)
            lines.append("$ tax_class " + " ".join(args) + f"  [exit {r.returncode}]\n" + r.stdout)
        (d / "tax_class_transcript.expected.txt").write_text("".join(lines))
    # ---- seeded synthetic taxonomy + BLAST hits + RDP lines
    d = HERE / "tax_synth"
    d.mkdir(exist_ok=True)
    tx = st.make_taxonomy(7, 300, 5000)
    st.write_dumps(tx, str(d))
    lines, ids, per = st.make_blast_hits(8, tx, 60)
    (d / "hits.txt").write_text("\n".join(lines) + "\n")
    run_reference_taxcollector(d, d / "hits.txt", d / "hits_class.expected.txt")
    ids, by = group_lineages(d / "hits_class.expected.txt")
    (d / "rdp.txt").write_text("\n".join(st.make_rdp_lines(9, ids, by)) + "\n")
    run_reference_consensus(d / "hits_class.expected.txt", d / "rdp.txt", d / "consensus.expected.txt")
    # ---- appendix A.3 consensus probes (X9 is a BLAST id absent from the RDP file)
    d = HERE / "consensus_probes"
    d.mkdir(exist_ok=True)
    (d / "blast_class.txt").write_text("\n".join(A3_BLAST) + "\n")
    (d / "rdp.txt").write_text("\n".join(A3_RDP) + "\n")
    run_reference_consensus(d / "blast_class.txt", d / "rdp.txt", d / "consensus.expected.txt")
    # ---- Trim join: the real Trim/trim2.4.pl on seeded QSEQ pairs and FASTQ
    sys.path.insert(0, str(REPO / "tests"))
    import oracle_pipeline as op
    from pangea_b200 import synth_trim as stt

    d = HERE / "trim"
    d.mkdir(exist_ok=True)
    a, b = stt.make_qseq_pair(41, 40)
    (d / "reads_A.qseq.txt").write_text(a)
    (d / "reads_B.qseq.txt").write_text(b)
    op.real_trim(d / "reads_A.qseq.txt", d / "reads_B.qseq.txt", 100, d / "qseq_g100.expected.fasta")
    op.real_trim(d / "reads_A.qseq.txt", d / "reads_B.qseq.txt", None, d / "qseq_default.expected.fasta")
    op.real_trim(d / "reads_A.qseq.txt", d / "reads_B.qseq.txt", 7, d / "qseq_g7_t5.expected.fasta", truncate=5)
    (d / "reads.fastq").write_text(stt.make_fastq(42, 30))
    op.real_trim(d / "reads.fastq", None, None, d / "fastq_single.expected.fasta")
    op.real_trim(d / "reads.fastq", d / "reads.fastq", 25, d / "fastq_paired_g25.expected.fasta")
    print("golden fixtures regenerated under", HERE)


if __name__ == "__main__":
    main()
