"""Seeded synthetic inputs for the tests and bench.py.

The reference's training file rdp_download_9178seqs.fa is absent from the
reference tree (.MISSING_LARGE_BLOBS), so configs 1-3 use synth16s(): a
6-level random taxonomy whose genus centroids diverge level by level from one
1500-base root, members = centroid + 1 % substitutions + 0.1 % 'n'.
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"acgt", dtype=np.uint8)
RANKS = ["domain", "phylum", "class", "order", "family", "genus"]
_COMP = np.arange(256, dtype=np.uint8)
for a, b in zip(b"acgtACGTuU", b"tgcaTGCAaA"):
    _COMP[a] = b


def _mutate(rng, seq, rate_var, var_mask, rate_cons=0.005):
    p = np.where(var_mask, rate_var, rate_cons)
    hit = rng.random(seq.size) < p
    out = seq.copy()
    out[hit] = BASES[rng.integers(0, 4, int(hit.sum()))]
    return out


def synth16s(seed: int, seqs: int, genera: int, length: int = 1500):
    """-> dict(data uint8, off int64, genus int32[seqs], G, anc int32[G,7], node_names, node_ranks)

    anc[g] = node ids root-first: Root, domain, phylum, class, order, family, genus.
    """
    rng = np.random.default_rng(seed)
    genera = int(genera)
    sizes = [min(2, genera), min(29, genera), min(max(genera // 15, 2), genera), min(max(genera // 6, 2), genera),
             min(max(genera // 3, 2), genera), genera]
    rates = [0.06, 0.05, 0.04, 0.03, 0.03, 0.02]
    var_mask = rng.random(length) < 0.60
    root = BASES[rng.integers(0, 4, length)]
    node_names, node_ranks, node_parent = ["Root"], ["rootrank"], [-1]
    prev_ids, prev_seqs = [0], [root]
    lineage_of = {0: [0]}
    for lvl, (cnt, rate) in enumerate(zip(sizes, rates)):
        ids, sq = [], []
        # every parent gets one child first, the rest pick a parent at random
        parents = list(range(len(prev_ids))) if cnt >= len(prev_ids) else list(rng.choice(len(prev_ids), cnt, replace=False))
        parents = parents[:cnt] + list(rng.integers(0, len(prev_ids), max(0, cnt - len(parents))))
        for j, pi in enumerate(parents):
            nid = len(node_names)
            node_names.append(f"{RANKS[lvl].capitalize()}{lvl}x{j:05d}")
            node_ranks.append(RANKS[lvl])
            node_parent.append(prev_ids[pi])
            lineage_of[nid] = lineage_of[prev_ids[pi]] + [nid]
            ids.append(nid)
            sq.append(_mutate(rng, prev_seqs[pi], rate, var_mask))
        prev_ids, prev_seqs = ids, sq
    G = len(prev_ids)
    anc = np.array([lineage_of[n] for n in prev_ids], dtype=np.int32)
    # Zipf(s=1) genus sizes, at least one member each
    w = 1.0 / np.arange(1, G + 1)
    extra = max(seqs - G, 0)
    cnts = np.ones(G, dtype=np.int64) + np.floor(extra * w / w.sum()).astype(np.int64)
    short = max(seqs, G) - int(cnts.sum())
    cnts[: max(short, 0)] += 1
    genus = np.repeat(np.arange(G, dtype=np.int32), cnts)
    rng.shuffle(genus)
    lens = rng.integers(int(length * 0.9), length + 1, genus.size)
    off = np.zeros(genus.size + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    data = np.empty(int(off[-1]), dtype=np.uint8)
    for i, g in enumerate(genus):
        s = _mutate(rng, prev_seqs[g], 0.01, var_mask, 0.01)
        nmask = rng.random(length) < 0.001
        s[nmask] = ord("n")
        data[off[i]:off[i + 1]] = s[: lens[i]]
    return dict(data=data, off=off, genus=genus, G=G, anc=anc, node_names=node_names, node_ranks=node_ranks,
                node_parent=np.array(node_parent, dtype=np.int32))


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def synth_reads(seed: int, train: dict, nreads: int, paired: bool = True, read_len: int = 250, gap: int = 189,
                err: float = 0.005, chunk: int = 50000):
    """Illumina-like reads drawn from the training members.

    paired: record = mateA + 'N'*gap + mateB (the Trim join, Trim/trim2.4.pl:228-244), else one mate.
    Half of the records are reverse-complemented whole.  Fixed record length => off is arithmetic.
    -> (data uint8 [nreads*L], off int64, source_genus int32)
    """
    rng = np.random.default_rng(seed)
    tdata, toff, tgenus = train["data"], train["off"], train["genus"]
    span = 2 * read_len + gap if paired else read_len
    L = span
    tlen = np.diff(toff)
    ok = np.nonzero(tlen >= span)[0]
    out = np.empty((nreads, L), dtype=np.uint8)
    src = np.empty(nreads, dtype=np.int32)
    ar = np.arange(span, dtype=np.int64)
    for c0 in range(0, nreads, chunk):
        cn = min(chunk, nreads - c0)
        m = ok[rng.integers(0, ok.size, cn)]
        start = toff[m] + rng.integers(0, tlen[m] - span + 1)
        win = tdata[start[:, None] + ar[None, :]]
        hit = rng.random(win.shape) < err
        win[hit] = BASES[rng.integers(0, 4, int(hit.sum()))]
        if paired:
            win[:, read_len:read_len + gap] = ord("N")
        flip = rng.random(cn) < 0.5
        win[flip] = _COMP[win[flip][:, ::-1]]
        out[c0:c0 + cn] = win
        src[c0:c0 + cn] = tgenus[m]
    off = np.arange(nreads + 1, dtype=np.int64) * L
    return out.reshape(-1), off, src


def read_fasta(path):
    """Minimal FASTA reader for tests -> (ids, headers, seqs as bytes)."""
    ids, hdr, seqs, cur = [], [], [], []
    with open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if ids:
                    seqs.append(b"".join(cur))
                cur = []
                h = line[1:].decode()
                hdr.append(h)
                ids.append(h.split()[0] if h.split() else "")
            elif ids:
                cur.append(bytes(line).replace(b" ", b""))
    if ids:
        seqs.append(b"".join(cur))
    return ids, hdr, seqs
