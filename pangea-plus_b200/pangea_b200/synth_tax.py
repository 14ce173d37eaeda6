"""Seeded synthetic NCBI-style taxonomy dumps, BLAST hit tables and RDP lines for Stages B and C
(SURVEY.md 8(d) config 5).  Everything stays inside the reference's defined behaviour:
names <= 60 bytes, lines < 511 bytes, every chain ends under node 1, sorted dumps."""
from __future__ import annotations

import os

import numpy as np

MAIN = ["superkingdom", "kingdom", "phylum", "class", "order", "family", "genus", "species"]
EXTRA_AFTER = {
    "phylum": "subphylum", "class": "subclass", "order": "suborder", "family": "subfamily",
    "genus": "subgenus", "species": "subspecies", "superkingdom": "no rank", "kingdom": "subkingdom",
}
SYLL = ["ba", "ci", "lus", "cocc", "us", "myc", "es", "bact", "er", "ium", "spir", "a", "thermo", "phil",
        "aceae", "ales", "ia", "ota", "vibrio", "rhizo", "pseudo", "monas", "strepto", "lacto"]


def _name(rng, cap=True):
    s = "".join(rng.choice(SYLL) for _ in range(int(rng.integers(2, 5))))
    return s.capitalize() if cap else s


def make_taxonomy(seed: int, nspecies: int, max_gi: int):
    """-> dict(nodes=[(taxid,parent,rank,embl)], names=[(taxid,name,unique,class)], gi=[(gi,taxid)])"""
    rng = np.random.default_rng(seed)
    nodes = {1: (1, "no rank", "")}
    names = {1: [("root", "", "scientific name"), ("all", "", "synonym")]}
    children: dict[tuple[int, str], list[int]] = {}
    next_id = [2]

    def new_node(parent, rank, name=None, embl=""):
        next_id[0] += int(rng.integers(1, 4))            # gaps => zero records in nodes.dmp.bin
        t = next_id[0]
        nodes[t] = (parent, rank, embl)
        nm = name if name is not None else _name(rng)
        recs = []
        if rng.random() < 0.3:
            recs.append((_name(rng) + " " + _name(rng, False), "", "synonym"))
        recs.append((nm, "", "scientific name"))
        if rng.random() < 0.2:
            recs.append((_name(rng, False), "", "common name"))
        if rng.random() < 0.05:
            recs.append((nm + " <" + _name(rng, False) + ">", nm + " <x>", "authority"))
        names[t] = recs
        children.setdefault((parent, rank), []).append(t)
        return t

    cell = new_node(1, "no rank", "cellular organisms")
    other = new_node(1, "no rank", "other sequences")
    sks = [new_node(cell, "superkingdom", n) for n in ("Bacteria", "Archaea", "Eukaryota")]
    leaves = []
    for _ in range(nspecies):
        r = rng.random()
        if r < 0.04:                                       # species under a no-rank whose parent is 1
            grp = new_node(other, "no rank", "artificial sequences") if rng.random() < 0.3 else other
            sp = new_node(grp, "species", "synthetic construct " + str(int(rng.integers(1, 99))))
            leaves.append(sp)
            continue
        cur = sks[int(rng.integers(0, 3))]
        if r < 0.12:                                       # species (almost) directly under the superkingdom
            k = (cur, "no rank")
            grp = children[k][0] if k in children and rng.random() < 0.8 else new_node(cur, "no rank", "unclassified " + _name(rng))
            nm = "uncultured " + _name(rng, False) if rng.random() < 0.6 else "Strain " + str(int(rng.integers(1, 80))) + " sp. X" + str(int(rng.integers(10, 99)))
            leaves.append(new_node(grp, "species", nm))
            continue
        genus_name = None
        for rank in MAIN[1:]:
            if rank == "kingdom" and rng.random() < 0.8:
                continue
            if rank in ("order", "family", "class") and rng.random() < 0.12:
                continue                                   # missing rank
            k = (cur, rank)
            if rank != "species" and k in children and rng.random() < 0.75:
                cur = children[k][int(rng.integers(0, len(children[k])))]
                if rank == "genus":
                    genus_name = names[cur][[c for _, _, c in names[cur]].index("scientific name")][0]
            else:
                if rank == "species":
                    gn = genus_name or _name(rng)
                    style = rng.random()
                    nm = gn + " " + _name(rng, False)
                    if style < 0.15:
                        nm = gn + " sp. " + str(int(rng.integers(1, 999))) + "A" + str(int(rng.integers(1, 9)))
                    cur = new_node(cur, rank, nm, embl="".join(rng.choice(list("ABCDEFGH")) for _ in range(2)) if rng.random() < 0.5 else "")
                else:
                    cur = new_node(cur, rank)
                    if rank == "genus":
                        genus_name = names[cur][[c for _, _, c in names[cur]].index("scientific name")][0]
            ex = EXTRA_AFTER.get(rank)
            if ex and rank != "species" and rng.random() < 0.15:
                k2 = (cur, ex)
                cur = children[k2][0] if k2 in children and rng.random() < 0.7 else new_node(cur, ex)
        leaf = cur
        if rng.random() < 0.1:
            leaf = new_node(cur, "no rank" if rng.random() < 0.5 else "subspecies", None)
        leaves.append(leaf)
    # a handful of nodes without any name record
    for t in list(nodes)[10::97]:
        if t not in sks and t != 1:
            names.pop(t, None)
    leaves = np.array(leaves)
    ngi = min(max_gi // 2, max(10, nspecies * 3))
    gis = np.sort(rng.choice(np.arange(1, max_gi + 1), size=ngi, replace=False))
    tax = leaves[rng.integers(0, len(leaves), ngi)]
    # a few gis point at internal nodes
    internal = np.array([t for t in nodes if t > 1])
    m = rng.random(ngi) < 0.03
    tax = np.where(m, internal[rng.integers(0, len(internal), ngi)], tax)
    node_list = sorted((t,) + nodes[t] for t in nodes)
    name_list = [(t,) + rec for t in sorted(names) for rec in names[t]]
    return dict(nodes=node_list, names=name_list, gi=list(zip(gis.tolist(), tax.tolist())), max_gi=int(max_gi))


def write_dumps(tx: dict, d: str) -> None:
    """writes nodes.dmp, names.dmp, gi_taxid_nucl.dmp into directory d (NCBI dump syntax)."""
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "nodes.dmp"), "w") as f:
        for t, parent, rank, embl in tx["nodes"]:
            f.write(f"{t}\t|\t{parent}\t|\t{rank}\t|\t{embl}\t|\t8\t|\t0\t|\t1\t|\t0\t|\t0\t|\t0\t|\t0\t|\t0\t|\t\t|\n")
    with open(os.path.join(d, "names.dmp"), "w") as f:
        for t, name, uniq, cls in tx["names"]:
            f.write(f"{t}\t|\t{name}\t|\t{uniq}\t|\t{cls}\t|\n")
    with open(os.path.join(d, "gi_taxid_nucl.dmp"), "w") as f:
        for g, t in tx["gi"]:
            f.write(f"{g}\t{t}\n")


def make_blast_hits(seed: int, tx: dict, nreads: int, max_hits: int = 20, unmapped: float = 0.02):
    """-> (lines list[str], read ids list[str], per-read hit gi lists).  -outfmt 6 lines:
    qid, gi|N|gb|ACC|, pident, length, mismatch, gapopen, qstart, qend, sstart, send, evalue, bitscore"""
    rng = np.random.default_rng(seed)
    gis = np.array([g for g, _ in tx["gi"]])
    lines, ids, per_read = [], [], []
    for r in range(nreads):
        qid = f"S{r:08d}"
        ids.append(qid)
        h = int(min(rng.geometric(0.2), max_hits))
        hit_gis = []
        for _ in range(h):
            if rng.random() < unmapped:
                g = int(rng.integers(1, tx["max_gi"] + 1))          # most likely a gap -> taxid 0
            else:
                g = int(gis[rng.integers(0, len(gis))])
            hit_gis.append(g)
            pid = f"{rng.uniform(75, 100):.2f}"
            bits = int(rng.integers(50, 1000))
            bit_s = (" " + str(bits)) if rng.random() < 0.1 else str(bits)     # BLAST pads short bitscores
            lines.append(f"{qid}\tgi|{g}|gb|AB{int(rng.integers(100000, 999999))}.1|\t{pid}\t{int(rng.integers(100, 250))}\t"
                         f"{int(rng.integers(0, 20))}\t{int(rng.integers(0, 3))}\t1\t{int(rng.integers(100, 250))}\t"
                         f"{int(rng.integers(1, 1000))}\t{int(rng.integers(1000, 2000))}\t{rng.uniform(0, 1e-20):.0e}\t{bit_s}")
        per_read.append(hit_gis)
    return lines, ids, per_read


RDP_RANKS = ["domain", "phylum", "class", "order", "family", "genus"]


def make_rdp_lines(seed: int, ids, class_lines_by_read, agree: float = 0.7):
    """RDP lines in the 5-TAB layout Consensus-1.1 parses: id, 5 TABs, then name/rank/conf triples.
    Names are drawn from the read's own BLAST lineages (agreement per rank), 10 % quoted,
    5 % with trailing digits, sometimes an extra 'subclass' triple."""
    rng = np.random.default_rng(seed)
    out = []
    for rid, lineages in zip(ids, class_lines_by_read):
        toks = {}
        for lin in lineages:
            for part in lin.split(";"):
                if part.startswith("[") and "]" in part:
                    k, v = part[1:].split("]", 1)
                    toks.setdefault(k, []).append(v)
        triples = []
        for i, rank in enumerate(RDP_RANKS):
            pool = toks.get(str(i), [])
            if pool and rng.random() < agree:
                nm = pool[int(rng.integers(0, len(pool)))]
            else:
                nm = _name(rng)
            r = rng.random()
            if r < 0.10:
                nm = '"' + nm + '"'
            elif r < 0.15:
                nm = nm + " " + str(int(rng.integers(1, 9)))
            triples += [nm, rank, f"{rng.integers(0, 101) / 100:.2f}".rstrip("0").rstrip(".") if rng.random() < 0.8 else "1.0"]
            if rank == "class" and rng.random() < 0.1:
                triples += [_name(rng), "subclass", "0.5"]
        out.append(rid + "\t\t\t\t\t" + "\t".join(triples))
    return out
