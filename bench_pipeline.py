#!/usr/bin/env python
"""bench_pipeline.py -- Stage B (lineage) and Stage C (consensus) throughput on one B200
(BASELINE.json configs[4] in miniature: synthetic BLAST/SOAP/RDP outputs, seeded).

Not the headline metric (bench.py is); this reports hit-lines/s for gi -> lineage and reads/s for
the consensus vote, through the C ABI with HOST buffers (text in, text/indices out), beside the
CPU restatements in oracle/ on the same input and all host cores they can use (both are serial,
like the reference).  Where the reference tree and perl exist (the build container) the real
Perl/C reference is timed on a small slice as well."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--species", type=int, default=20000)
    ap.add_argument("--max-gi", type=int, default=5_000_000)
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--mega-reads", type=int, default=400000, help="reads of the Megaclust input (about 3 lines each)")
    args = ap.parse_args()

    import oracle_pipeline as op
    import pangea_b200 as pg
    from pangea_b200 import synth_tax as st

    tmp = Path(tempfile.mkdtemp(prefix="pgpipe"))
    tx = st.make_taxonomy(0x7A70, args.species, args.max_gi)
    st.write_dumps(tx, str(tmp / "tax"))
    pg.tax_build(tmp / "tax")
    lines, ids, per = st.make_blast_hits(0x7A71, tx, args.reads)
    hits = tmp / "hits.txt"
    hits.write_text("\n".join(lines) + "\n")
    gi = np.array([int(l.split("|", 2)[1]) for l in lines], np.int32)
    nh = len(lines)

    ctx = pg.Context(0)
    t0 = time.perf_counter()
    tax = ctx.tax_load(tmp / "tax")
    load_s = time.perf_counter() - t0
    tax.lineage(gi[:1000])
    # timed: the C ABI call (host gi in, host strings + offsets out, buffers allocated once)
    lbuf, loff = tax.lineage_raw(gi)
    best = 1e9
    for _ in range(args.repeat):
        t0 = time.perf_counter()
        tax.lineage_raw(gi, lbuf, loff)
        best = min(best, time.perf_counter() - t0)
    lineage_rate = nh / best
    lin = tax.lineage(gi)

    t0 = time.perf_counter()
    cls = tmp / "class.txt"
    assert op.oracle_taxcollector(tmp / "tax", hits, cls) == 0
    cpu_lineage_s = time.perf_counter() - t0
    want = [l.split("\t")[1].encode() for l in cls.read_text().split("\n") if l]
    assert lin == want, "GPU lineages differ from the oracle"

    # ---- consensus
    ids2, by = op.group_lineages(cls)
    rdp_lines = st.make_rdp_lines(0x7A72, ids2, by)
    rdp = tmp / "rdp.txt"
    rdp.write_text("\n".join(rdp_lines) + "\n")
    hit_off = np.zeros(len(ids2) + 1, np.int64)
    hit_off[1:] = np.cumsum([len(b) for b in by])
    cl = [l.split("\t") for l in cls.read_text().split("\n") if l]
    lin_b = [f[1].encode() for f in cl]
    pid_b = [f[2].encode() for f in cl]
    rdp_b = [l.split("\t" * 5, 1)[1].encode() for l in rdp_lines]
    ctx.consensus(hit_off[:101], lin_b[:int(hit_off[100])], pid_b[:int(hit_off[100])], rdp_b[:100])
    packed = (pg.pack_sequences(lin_b), pg.pack_sequences(pid_b), pg.pack_sequences(rdp_b))
    best_c = 1e9
    for _ in range(args.repeat):
        t0 = time.perf_counter()
        win, nm = ctx.consensus_packed(hit_off, *packed)      # the C ABI call: host text buffers in, indices out
        best_c = min(best_c, time.perf_counter() - t0)
    t0 = time.perf_counter()
    out = tmp / "cons.txt"
    assert op.oracle_consensus(cls, rdp, out) == 0
    cpu_cons_s = time.perf_counter() - t0
    got = out.read_text().split("\n")
    ref_nm = [int(l.split(": ")[1]) for l in got if l.startswith("#Matches")]
    assert ref_nm == nm.tolist(), "GPU match counts differ from the oracle"
    ref_win = [l for l in got if l and not l.startswith("#")]
    clines = [l for l in cls.read_text().split("\n") if l]
    assert ref_win == [clines[w] for w in win], "GPU winners differ from the oracle"

    line = {
        "metric": "lineage hit-lines/s and consensus reads/s (Stage B/C, one B200, host text buffers in and out)",
        "hit_lines": nh, "reads": len(ids2), "taxids": len(tx["nodes"]), "max_gi": args.max_gi,
        "lineage": {"value": lineage_rate, "unit": "hit-lines/s", "tax_load_s": load_s,
                    "cpu_port": {"value": nh / cpu_lineage_s, "unit": "hit-lines/s", "cores": 1,
                                 "what": "oracle/taxcollector_ref.c incl. file I/O (the reference forks tax_class per lookup: ~1e2 hit-lines/s, BASELINE.md)"}},
        "consensus": {"value": len(ids2) / best_c, "unit": "reads/s",
                      "cpu_port": {"value": len(ids2) / cpu_cons_s, "unit": "reads/s", "cores": 1,
                                   "what": "oracle/consensus_ref.c incl. file I/O (the Perl script: ~1.4e3 reads/s, BASELINE.md)"}},
        "gpu_launches": ctx.launch_count(),
        "parity": "lineage strings, winners and match counts identical to the oracle on the whole input",
    }
    # ---- Megaclust: thresholds + OTU counting over consensus-like text
    from pangea_b200 import synth_mega

    mtext = synth_mega.make_consensus_text(0x7A73, args.mega_reads, otus=5000)
    mlines = mtext.count(b"\n")
    ctx.megaclust(mtext[: 1 << 16])
    mbufs = (np.zeros(mlines + 2, np.int64), np.zeros(mlines + 2, np.int32), np.zeros(mlines + 2, np.int64))
    best_m = 1e9
    for _ in range(args.repeat):
        t0 = time.perf_counter()
        ctx.megaclust_raw(mtext, sim=80.0, bits=100.0, bufs=mbufs)           # the C ABI call: host text in, OTU arrays out
        best_m = min(best_m, time.perf_counter() - t0)
    mres = ctx.megaclust(mtext, sim=80.0, bits=100.0)
    t0 = time.perf_counter()
    mref = op.oracle_megaclust(mtext, sim=80.0, bits=100.0)
    cpu_m = time.perf_counter() - t0
    assert mres == mref, "GPU OTU counts differ from the oracle"
    line["megaclust"] = {"value": mlines / best_m, "unit": "lines/s", "lines": mlines, "otus": len(mres[0]),
                         "cpu_port": {"value": mlines / cpu_m, "unit": "lines/s", "cores": 1, "what": "oracle/megaclust_ref.c"}}
    if op.have_megaclust_reference():
        t0 = time.perf_counter()
        op.real_megaclust(mtext[: mtext.find(b"\n", len(mtext) // 8) + 1], ["-s", "80", "-b", "100"])
        line["megaclust"]["reference_perl"] = {"value": (mlines / 8) / (time.perf_counter() - t0), "unit": "lines/s",
                                               "sample": "first eighth of the input, megaclust2.pl"}
    if op.have_reference():
        n_ref = min(300, nh)
        h2 = tmp / "hits_small.txt"
        h2.write_text("\n".join(lines[:n_ref]) + "\n")
        t0 = time.perf_counter()
        op.real_taxcollector(tmp / "tax", h2, tmp / "real_class.txt")
        line["lineage"]["reference_perl_c"] = {"value": n_ref / (time.perf_counter() - t0), "unit": "hit-lines/s",
                                                "sample": f"first {n_ref} hit lines, incl. tax_class -c"}
    print(json.dumps(line), flush=True)
    tax.free()
    ctx.close()


if __name__ == "__main__":
    main()
