"""Seeded synthetic Illumina QSEQ / FASTQ files for the Trim join."""
from __future__ import annotations

import numpy as np


def _qual(rng, n, off, good=0.9):
    base = np.clip(38 - np.arange(n) * (0.08 + 0.2 * rng.random()) + rng.normal(0, 4, n), 2, 40).astype(int)
    base[rng.random(n) < 0.03] = 2
    if rng.random() > good:
        base[:] = np.clip(rng.normal(12, 6, n), 2, 40).astype(int)      # a bad read: trimmed below the length cutoff
    return bytes((base + off).tolist())


def _seq(rng, n, dot=True):
    s = rng.choice(np.frombuffer(b"ACGT", np.uint8), n)
    s[rng.random(n) < 0.01] = ord(".") if dot else ord("N")
    return s.tobytes()


def make_qseq_pair(seed: int, nreads: int, lo: int = 90, hi: int = 170):
    """-> (text of read-1 file, text of read-2 file): 11 TAB-separated QSEQ columns
    machine, run, lane, tile, x, y, index, read number, sequence ('.' = no call), quality (+64), filter"""
    rng = np.random.default_rng(seed)
    A, B = [], []
    for i in range(nreads):
        n1, n2 = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
        head = f"HWI-ST{seed}\t7\t{1 + i % 8}\t{1100 + i % 32}\t{int(rng.integers(1000, 20000))}\t{int(rng.integers(1000, 200000))}\t0"
        flt = "1" if rng.random() < 0.9 else "0"
        A.append(f"{head}\t1\t{_seq(rng, n1).decode()}\t{_qual(rng, n1, 64).decode()}\t{flt}")
        B.append(f"{head}\t2\t{_seq(rng, n2).decode()}\t{_qual(rng, n2, 64).decode()}\t{flt}")
    return "\n".join(A) + "\n", "\n".join(B) + "\n"


def make_fastq(seed: int, nreads: int, lo: int = 90, hi: int = 170):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(nreads):
        n = int(rng.integers(lo, hi))
        out.append(f"@R{seed}_{i} 1:N:0\n{_seq(rng, n, dot=False).decode()}\n+\n{_qual(rng, n, 33).decode()}")
    return "\n".join(out) + "\n"
