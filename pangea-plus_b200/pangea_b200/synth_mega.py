"""Seeded synthetic input of the Megaclust stage: consensus-output-like text --
BLAST tabular lines whose subject is a lineage string, interleaved with `#Matches found:` comment lines,
plus the edge shapes of megaclust2.pl's split and numeric rules."""
from __future__ import annotations

import numpy as np


def make_consensus_text(seed: int, reads: int, otus: int = 300, edge: bool = True) -> bytes:
    rng = np.random.default_rng(seed)
    ranks = ["Bacteria", "Archaea", "Eukaryota"]
    lineages = []
    for i in range(otus):
        d = ranks[int(rng.integers(0, 3))]
        depth = int(rng.integers(2, 8))
        parts = [f"[0]{d}"] + [f"[{k}]Taxon{k}_{int(rng.integers(0, 5 + 20 * k))}" for k in range(1, depth)]
        lineages.append(";".join(parts) + ";")
    w = 1.0 / np.arange(1, otus + 1)
    w /= w.sum()
    out = []
    for r in range(reads):
        nhit = 1 + int(rng.integers(0, 3))
        q = f"read{r % max(reads * 3 // 4, 1):07d}"            # some queries come back: distinct-pair counting matters
        for _ in range(nhit):
            lin = lineages[int(rng.choice(otus, p=w))]
            pid = rng.uniform(70, 100)
            ev = 10.0 ** -rng.uniform(5, 80)
            bits = int(rng.integers(50, 900))
            sep = "\t\t" if rng.random() < 0.1 else "\t"       # taxcollector output carries doubled TABs now and then
            out.append(f"{q}{sep}{lin}\t{pid:.2f}\t250\t3\t0\t1\t250\t1\t250\t{ev:.0e}\t{bits}")
        out.append(f"#Matches found: {int(rng.integers(0, 7))}")
    if edge:
        out += [
            "",                                                 # blank line: examined, beyond
            "q1 \tsubjB\t96\t1\t1\t1\t1\t1\t1\t1\t1e-20\t200",  # blank before TAB; thresholds met exactly
            "q4\t\tsubjC\t95\t1\t1\t1\t1\t1\t1\t1\t0.0\t 937",  # doubled TAB, blank-padded bitscore
            "q5\tsubjD\t94.99\t1\t1\t1\t1\t1\t1\t1\t0.0\t500",
            "q6\tsubjE\t95abc\t1\t1\t1\t1\t1\t1\t1\t1E-21x\t2e2",
            "q7\tsubjF\t 96\t1\t1\t1\t1\t1\t1\t1\t.5e-20\t+200.0",
            "q8\tsubjG\tinf\t1\t1\t1\t1\t1\t1\t1\tnan\t200",
            "q9\tsubjH",                                        # too few fields: every number reads 0
            "# a comment",
            "q1\tsubjB\t99\t1\t1\t1\t1\t1\t1\t1\t1e-30\t300",   # same query, same subject again
            "q1\tsubjB \t99\t1\t1\t1\t1\t1\t1\t1\t1e-30\t300",  # "subjB" then " \t" delimiter: same subject
            "q10\tsubj,with,commas\t100\t1\t1\t1\t1\t1\t1\t1\t0\t1000",
        ]
    return ("\n".join(out) + "\n").encode()
