"""Reads of many different lengths (400..600 bases, ~200 distinct word counts): plan 4 against plan 3 (cert_plan=3),
records compared.  usage: python scripts/probe/varied.py [reads]"""
import sys, time
import numpy as np
sys.path.insert(0, "pangea-plus_b200"); sys.path.insert(0, "tests")
import pangea_b200 as pg
from pangea_b200 import synth, pack_sequences

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
tr = synth.synth16s(0x9178, 9178, 1219)
ctx = pg.Context(0)
gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"]); gm.set_lineage(tr["anc"])
rng = np.random.default_rng(3)
seq_i = rng.integers(0, len(tr["off"]) - 1, n); lens = rng.integers(400, 601, n)
reads = []
for i, ln in zip(seq_i, lens):
    s = tr["data"][tr["off"][i]:tr["off"][i + 1]]
    p = int(rng.integers(0, len(s) - ln)); reads.append(s[p:p + ln].tobytes())
data, off = pack_sequences(reads)
out = {}
for name, kw in (("plan 4", {}), ("plan 3", dict(cert_plan=3))):
    ctx.classify(gm, data, off, mode=1, **kw)                      # warm: sample lists, count images
    t0 = time.perf_counter(); res = ctx.classify(gm, data, off, mode=1, **kw); dt = time.perf_counter() - t0
    out[name] = res
    print(f"{name}: {n / dt / 1e6:.2f} M reads/s end to end (host buffers), {ctx.classify_stats()}")
assert out["plan 4"].tobytes() == out["plan 3"].tobytes()
print("records identical")
