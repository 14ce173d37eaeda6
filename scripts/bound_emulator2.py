"""CPU emulation, second design: level-1 bound over the read's most informative positions only.
(design tool for round 2; see bound_emulator.py)"""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200")); sys.path.insert(0, str(REPO / "tests")); sys.path.insert(0, str(REPO / "scripts"))
import oracle_rdp as ora
from pangea_b200 import synth
from bound_emulator import layout_from_lineage

G = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
seqs = int(sys.argv[2]) if len(sys.argv) > 2 else 6 * G
nreads = int(sys.argv[3]) if len(sys.argv) > 3 else 64
paired = (sys.argv[4] != "single") if len(sys.argv) > 4 else True
tr = synth.synth16s(0x3000000 if G >= 5000 else 0x9178, seqs, G)
om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
_, _, logP = om.tables()
_, nw, _, _ = om.counts()
pos = layout_from_lineage(tr["anc"]); nblk = len(pos) // 64; ok = pos >= 0
rowmax = logP.max(axis=1); vmax = float(np.abs(logP).max())
# full bm table, chunked
bm_all = np.empty((65536, nblk), np.uint16)
for w0 in range(0, 65536, 2048):
    rows = logP[w0:w0 + 2048]
    q = np.minimum(np.floor((rowmax[w0:w0 + 2048, None].astype(np.float64) - rows) * 128.0), 4095).astype(np.uint16)
    qp = np.full((rows.shape[0], nblk * 64), 4095, np.uint16); qp[:, ok] = q[:, pos[ok]]
    bm_all[w0:w0 + 2048] = qp.reshape(-1, nblk, 64).min(axis=2)
score_mean = bm_all.mean(axis=1)                       # ideal-ish per-word informativeness
score_med = np.median(bm_all, axis=1)
print("bm table done; words with median bm > 500:", int((score_med > 500).sum()), flush=True)

data, off, src = synth.synth_reads(0x250, tr, nreads, paired=paired)
ref = om.classify(data, off)
stats = []
for i in range(nreads):
    seq = data[off[i]:off[i + 1]]
    if ref["reversed"][i]: seq = synth.revcomp(seq)
    w = ora.words(seq.tobytes()); n = len(w); k = n // 8
    q = np.minimum(np.floor((rowmax[w, None].astype(np.float64) - logP[w]) * 128.0), 4095).astype(np.int64)
    qp = np.full((n, nblk * 64), 4095, np.int64); qp[:, ok] = q[:, pos[ok]]
    bm = bm_all[w].astype(np.int64)
    draws = ora.jrandom_stream(1, n, 100 * k).reshape(100, k)
    u = 2.0 ** -24; m_ = (k - 1) * u
    margin = k + int(np.ceil(2 * (m_ / (1 - m_) * k * vmax * 1.0001) * 128)) + 1
    exact = qp[draws].sum(axis=1); champ = exact.min(axis=1); thr = champ + margin
    best_blk = int(np.argmin(qp.sum(axis=0)) // 64)
    lb16 = bm[draws].sum(axis=1); open16 = lb16 <= thr[:, None]; open16[:, best_blk] = False
    out = {"n": n, "open16": int(open16.sum()), "blk16": int(open16.any(axis=0).sum())}
    for name, sc in (("mean", score_mean[w]), ("med", score_med[w])):
        rank = np.argsort(-sc, kind="stable")
        for m in (48, 64, 96, 128):
            keep = np.zeros(n, bool); keep[rank[:m]] = True
            bmk = bm * keep[:, None]
            lb = bmk[draws].sum(axis=1)
            o = lb <= thr[:, None]; o[:, best_blk] = False
            out[f"{name}{m}_open"] = int(o.sum()); out[f"{name}{m}_blk"] = int(o.any(axis=0).sum())
            out[f"{name}{m}_draws"] = float(keep[draws].sum(axis=1).mean())
    stats.append(out)
    if i < 4: print(i, out, flush=True)
print("---- means over", nreads, "reads; pairs per read =", 100 * nblk, "blocks", nblk)
for k_ in stats[0]:
    v = np.array([s[k_] for s in stats], float)
    print(f"{k_:18s} mean {v.mean():10.1f}  median {np.median(v):10.1f}  p90 {np.percentile(v, 90):10.1f}  max {v.max():10.1f}")
