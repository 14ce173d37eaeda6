"""CPU emulation of the certified path's block lower bounds (design tool, not product, not a test).

Builds a synth16s model with the CPU oracle, restates the deficit table / lineage-ordered block layout /
block minima of pg_certified.cu in numpy, and counts for a sample of reads how many (task, block) pairs
each candidate bound leaves open.  Used in round 2 to size the coarse first-level bound for 10 000-genus
models before spending GPU time on it.

usage: python scripts/bound_emulator.py [genera] [seqs] [reads]
"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))
import oracle_rdp as ora  # noqa: E402
from pangea_b200 import synth  # noqa: E402


def layout_from_lineage(anc):
    """python restatement of pg_model_layout_from_lineage (pg_certified.cu)"""
    G, depth = anc.shape
    order = sorted(range(G), key=lambda g: tuple(anc[g]))
    units = []
    stack = [(0, G, 0)]
    while stack:
        a, b, lv = stack.pop()
        if b - a <= 64 or lv >= depth:
            for s0 in range(a, b, 64):
                units.append((s0, min(s0 + 64, b)))
            continue
        kids = []
        i = a
        while i < b:
            j = i + 1
            while j < b and anc[order[j]][lv] == anc[order[i]][lv]:
                j += 1
            kids.append((i, j, lv + 1))
            i = j
        stack.extend(reversed(kids))
    units.sort()
    pos, fill = [], 0
    for a, b in units:
        if fill + (b - a) > 64:
            pos.extend([-1] * ((-len(pos)) % 64))
            fill = 0
        pos.extend(order[a:b])
        fill += b - a
    pos.extend([-1] * ((-len(pos)) % 64))
    return np.array(pos, np.int32)


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 6 * G
    nreads = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    tr = synth.synth16s(0x3000000 if G >= 5000 else 0x9178, seqs, G)
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    _, _, logP = om.tables()
    print("model built", logP.shape, flush=True)
    pos = layout_from_lineage(tr["anc"])
    nblk = len(pos) // 64
    print("blocks", nblk, "padding", int((pos < 0).sum()), flush=True)
    rowmax = logP.max(axis=1)
    vmax = float(np.abs(logP).max())

    data, off, src = synth.synth_reads(0x250, tr, nreads, paired=True)
    ref = om.classify(data, off)
    k_stats = []
    for i in range(nreads):
        seq = data[off[i]:off[i + 1]]
        if ref["reversed"][i]:
            seq = synth.revcomp(seq)
        w = ora.words(seq.tobytes())
        n = len(w)
        k = n // 8
        # deficits of this read's words only
        rows = logP[w]                                        # [n, G]
        q = np.floor((rowmax[w, None].astype(np.float64) - rows.astype(np.float64)) * 128.0).astype(np.int64)
        q = np.minimum(q, 4095)
        qp = np.full((n, nblk * 64), 4095, np.int64)
        ok = pos >= 0
        qp[:, ok] = q[:, pos[ok]]
        bm = qp.reshape(n, nblk, 64).min(axis=2)              # [n, nblk]
        draws = ora.jrandom_stream(1, n, 100 * k).reshape(100, k)
        u = 2.0 ** -24
        m = (k - 1) * u
        margin = k + int(np.ceil(2 * (m / (1 - m) * k * vmax * 1.0001) * 128)) + 1
        exact = qp[draws].sum(axis=1)                         # [100, npos]
        champ = exact.min(axis=1)                             # [100]
        thr = champ + margin
        lb16 = bm[draws].sum(axis=1)                          # [100, nblk]
        open16 = lb16 <= thr[:, None]
        best_blk = int(np.argmin(qp.sum(axis=0)) // 64)
        out = {"n": n, "open16": int(open16.sum()), "thr": float(thr.mean()), "lbfar": float(np.median(lb16))}
        for shift, cap in ((7, 15), (6, 15), (8, 15), (7, 7), (6, 7)):
            c = np.minimum(bm >> shift, cap)
            lbc = c[draws].sum(axis=1) << shift
            oc = lbc <= thr[:, None]
            grp_open = oc.reshape(100, -1)[:, : (nblk // 28) * 28].reshape(100, nblk // 28, 28).any(axis=2) if nblk >= 28 else oc.any(axis=1, keepdims=True)
            # groups of 28 blocks with anything open for any task
            out[f"open_s{shift}c{cap}"] = int(oc.sum())
            out[f"blk_s{shift}c{cap}"] = int(oc.any(axis=0).sum())        # blocks open for at least one task
            out[f"grp_s{shift}c{cap}"] = int(grp_open.any(axis=0).sum())
            # first-16-draw version of the coarse bound (one trip)
            lb1 = c[draws[:, :16]].sum(axis=1) << shift
            out[f"open1_s{shift}c{cap}"] = int((lb1 <= thr[:, None]).sum())
            lb2 = c[draws[:, :32]].sum(axis=1) << shift
            out[f"open2_s{shift}c{cap}"] = int((lb2 <= thr[:, None]).sum())
        out["blk16"] = int(open16.any(axis=0).sum())
        k_stats.append(out)
        if i < 8:
            print(i, out, "best", best_blk, flush=True)
    keys = [k_ for k_ in k_stats[0] if k_ != "n"]
    print("---- means over", nreads, "reads; pairs per read =", 100 * nblk)
    for k_ in keys:
        v = np.array([s[k_] for s in k_stats], float)
        print(f"{k_:18s} mean {v.mean():10.1f}  median {np.median(v):10.1f}  p90 {np.percentile(v, 90):10.1f}  max {v.max():10.1f}")


if __name__ == "__main__":
    main()
