"""Golden files of the Megaclust stage, made by the reference's own Perl (run in the build container):
   tests/golden/megaclust/input.txt       seeded consensus-like text incl. the edge lines
   tests/golden/megaclust/expected.json   per option set: sorted output lines + stdout of megaclust2.pl
   tests/golden/megaclust/tables.json     megaclustable.pl inputs and the tables it wrote"""
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))
import oracle_pipeline as op  # noqa: E402
from pangea_b200 import synth_mega  # noqa: E402
from test_megaclust_cpu import CASES  # noqa: E402

out = REPO / "tests" / "golden" / "megaclust"
out.mkdir(parents=True, exist_ok=True)
text = synth_mega.make_consensus_text(0x3E6A, 1500, otus=120)
(out / "input.txt").write_bytes(text)
exp = {}
for name, case in CASES.items():
    lines, header, stdout = op.real_megaclust(text, case["args"])
    exp[name] = {"lines": [l.decode() for l in lines], "stdout": stdout.decode()}
(out / "expected.json").write_text(json.dumps(exp, indent=1))

# megaclustable: two megaclust outputs (made by the live megaclust2.pl, so in Perl's hash order) + hand-made quirks
t2 = synth_mega.make_consensus_text(0x3E6B, 900, otus=120, edge=False)
f1 = b"OTU,times_hit\n" + b"\n".join(op.real_megaclust(text, ["-s", "80", "-b", "100"])[0]) + b"\n"
f2 = b"OTU,times_hit\n" + b"\n".join(op.real_megaclust(t2, ["-s", "80", "-b", "100"])[0]) + b"\n"
quirk = b"OTU,times_hit\n[0]Bacteria;[1]X;,3\n[0]Bacteria,4\n[0]Archaea;[1]Y;,abc\n[0]NoCount\nplain line,9\n[1]Z;[0]Late;,2.5\n[0]Bacteria;[2]W;,1e2\n"
tables = {}
for name, files, level in (("domain", {"a.txt": f1, "b.txt": f2}, "0"), ("phylum", {"a.txt": f1, "b.txt": f2}, "1"),
                           ("quirks", {"q.txt": quirk, "a.txt": f1}, "0"), ("species", {"a.txt": f1, "q.txt": quirk}, "6")):
    tab = op.real_megaclustable(files, level)
    tables[name] = {"files": {k: v.decode() for k, v in files.items()}, "order": list(files), "level": level,
                    "table": tab.decode()}
(out / "tables.json").write_text(json.dumps(tables, indent=1))
print("written", out)
