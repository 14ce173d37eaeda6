"""CPU tests of the Stage A oracle (oracle/rdp_ref.c).

The RDP 2.5 jar is not vendored by the reference and cannot run here, and the
reference stores no RDP output, so parity with the jar is UNPINNED.  What these
tests pin: the java.util.Random known-answer streams, hand
checked 8-mer ids (A1/A3), a toy model evaluated independently in numpy
(A5-A8), and the structural invariants of A9.
"""
import numpy as np
import pytest

import oracle_rdp as ora
from pangea_b200 import pack_sequences, synth

# ---------------------------------------------------------------- A8: java.util.Random

JAVA_VECTORS = {
    (1, 243): [78, 28, 127, 213, 89, 229, 113, 226, 217, 37, 232, 232],
    (1, 486): [321, 28, 127, 213, 332, 472, 356, 226, 460, 280, 475, 475],
    (1, 1418): [235, 290, 37, 503, 1158, 736, 1176, 682, 744, 580, 217, 893],
    (1, 64): [46, 6, 26, 26, 13, 2, 21, 42, 61, 45, 0, 9],
    (1, 5): [0, 3, 2, 3, 4, 4, 4, 1, 3, 3, 4, 3],
    (42, 10): [0, 3, 8, 4, 0],
}


@pytest.mark.parametrize("key", list(JAVA_VECTORS))
def test_java_random_known_answers(key):
    seed, n = key
    want = JAVA_VECTORS[key]
    assert ora.jrandom_stream(seed, n, len(want)).tolist() == want


def test_java_random_nextint_plain():
    out = np.zeros(3, np.int32)
    ora.lib().rdp_jrandom_ints(1, 3, out.ctypes.data)
    assert out.tolist() == [-1155869325, 431529176, 1761283695]


class PyJavaRandom:
    """Independent restatement of java.util.Random for cross-checking the C one."""
    M = (1 << 48) - 1

    def __init__(self, seed):
        self.s = (seed ^ 0x5DEECE66D) & self.M

    def next(self, bits):
        self.s = (self.s * 0x5DEECE66D + 0xB) & self.M
        v = self.s >> (48 - bits)
        return v - (1 << 32) if v >= (1 << 31) else v

    def next_int(self, n):
        if n & (-n) == n:
            return (n * self.next(31)) >> 31
        while True:
            bits = self.next(31)
            val = bits % n
            if bits - val + (n - 1) < (1 << 31):
                return val


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 243, 1000, 4096, 6999, (1 << 30) + 1])
def test_java_random_vs_python(n):
    r = PyJavaRandom(1)
    want = [r.next_int(n) for _ in range(500)]
    assert ora.jrandom_stream(1, n, 500).tolist() == want


# ---------------------------------------------------------------- A1 / A3: words

def test_word_ids_by_hand():
    # A=0 T/U=1 G=2 C=3, oldest base most significant
    assert ora.words(b"ACGTACGT").tolist() == [0x3939]
    assert ora.words(b"AAAAAAAAC").tolist() == [0, 3]
    assert ora.words(b"acguACGU").tolist() == [0x3939]
    # an N (or any IUPAC code) kills every 8-mer touching it
    assert ora.words(b"ACGTNACGTACGT").tolist() == [0x3939]
    assert ora.words(b"ACGTACGNACGTACG").tolist() == []
    assert ora.words(b"ACGTAC").tolist() == []
    # duplicates are kept, order is sequence order
    assert ora.words(b"AAAAAAAAAA").tolist() == [0, 0, 0]


def test_revcomp_word_by_hand():
    L = ora.lib()
    assert L.rdp_revcomp_word(0x3939) == 0x3939            # ACGTACGT is its own reverse complement
    assert L.rdp_revcomp_word(3) == 0x9555                 # AAAAAAAC -> GTTTTTTT
    for w in [0, 1, 77, 0x1234, 0xFFFF, 0xABCD]:
        assert L.rdp_revcomp_word(L.rdp_revcomp_word(w)) == w


def test_words_of_reverse_complement():
    rng = np.random.default_rng(5)
    seq = synth.BASES[rng.integers(0, 4, 300)].copy()
    seq[[17, 140]] = ord("n")
    fwd = ora.words(seq.tobytes())
    rc = ora.words(synth.revcomp(seq).tobytes())
    L = ora.lib()
    assert rc.tolist() == [L.rdp_revcomp_word(int(w)) for w in fwd[::-1]]


# ---------------------------------------------------------------- A5-A8: toy model, independent numpy evaluation

def py_train(seqs, genus, G):
    m = np.zeros((65536, G), np.int32)
    nw = np.zeros(65536, np.int32)
    M = np.zeros(G, np.int32)
    for s, g in zip(seqs, genus):
        for w in set(ora.words(s).tolist()):
            m[w, g] += 1
            nw[w] += 1
        M[g] += 1
    N = len(seqs)
    f32 = np.float32
    Pw = (nw.astype(f32) + f32(0.5)) / (f32(N) + f32(1.0))
    logPrior = np.log(Pw.astype(np.float64)).astype(f32)
    logLeave = np.log((M.astype(f32) + f32(1.0)).astype(np.float64)).astype(f32)
    q = (m.astype(f32) + Pw[:, None]) / (M.astype(f32) + f32(1.0))[None, :]
    logP = np.where(m > 0, np.log(q.astype(np.float64)).astype(f32), logPrior[:, None] - logLeave[None, :]).astype(f32)
    return m, nw, M, N, logPrior, logLeave, logP


def py_classify(tables, seq, min_boot=0):
    logPrior, logLeave, logP = tables
    f32 = np.float32
    if len(seq) < 50:
        return None
    w = ora.words(seq)
    L = ora.lib()
    fwd = rev = f32(0)
    for x in w:
        fwd = f32(fwd + logPrior[x])
        rev = f32(rev + logPrior[L.rdp_revcomp_word(int(x))])
    reversed_ = bool(rev > fwd)
    if reversed_:
        w = np.array([L.rdp_revcomp_word(int(x)) for x in w[::-1]], np.int32)
    n = len(w)
    G = logP.shape[1]
    acc = np.zeros(G, f32)
    for x in w:
        acc = (acc + logP[x]).astype(f32)      # elementwise fp32 add, one rounding per step
    genus = int(np.argmax(acc)) if n else 0    # np.argmax = first maximum
    score = acc[genus] if n else f32(0)
    k = max(n // 8, min_boot)
    rng = PyJavaRandom(1)
    boot = []
    for _ in range(100):
        acc = np.zeros(G, f32)
        for _ in range(k):
            acc = (acc + logP[w[rng.next_int(n)]]).astype(f32)
        boot.append(int(np.argmax(acc)))
    return genus, n, score, reversed_, boot


@pytest.fixture(scope="module")
def toy():
    tr = synth.synth16s(seed=11, seqs=40, genera=6, length=400)
    seqs = [tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes() for i in range(len(tr["genus"]))]
    model = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    yield tr, seqs, model
    model.free()


def test_toy_counts_and_tables(toy):
    tr, seqs, model = toy
    m, nw, M, N, logPrior, logLeave, logP = py_train(seqs, tr["genus"], tr["G"])
    om, onw, oM, oN = model.counts()
    assert oN == N and np.array_equal(oM, M) and np.array_equal(onw, nw) and np.array_equal(om, m)
    olp, oll, ot = model.tables()
    assert np.array_equal(olp.view(np.uint32), logPrior.view(np.uint32))
    assert np.array_equal(oll.view(np.uint32), logLeave.view(np.uint32))
    assert np.array_equal(ot.view(np.uint32), logP.view(np.uint32))
    # hand check of one formula instance in double
    w = int(np.nonzero(nw)[0][0])
    assert abs(float(logPrior[w]) - np.log((nw[w] + 0.5) / (N + 1.0))) < 1e-6


def test_toy_classify_matches_numpy(toy):
    tr, seqs, model = toy
    tables = model.tables()
    rng = np.random.default_rng(3)
    reads = []
    for i in range(12):
        s = np.frombuffer(seqs[i], np.uint8)[: 120 + 17 * i].copy()
        if i % 3 == 0:
            s = synth.revcomp(s)
        if i % 4 == 1:
            s[30] = ord("N")
        reads.append(s.tobytes())
    reads.append(b"ACGT" * 5)                      # short: < 50 bases
    reads.append(b"N" * 60)                        # long enough, zero words
    data, off = pack_sequences(reads)
    res = model.classify(data, off, threads=2)
    for i, r in enumerate(reads):
        want = py_classify(tables, r)
        if want is None:
            assert res[i]["status"] == 1 and res[i]["genus"] == -1
            continue
        genus, n, score, reversed_, boot = want
        assert res[i]["status"] == 0
        assert (res[i]["genus"], res[i]["n_words"], bool(res[i]["reversed"])) == (genus, n, reversed_)
        assert np.float32(res[i]["score"]).view(np.uint32) == np.float32(score).view(np.uint32)
        assert res[i]["boot"].tolist() == boot


def test_votes_invariants(toy):
    tr, seqs, model = toy
    data, off = pack_sequences(seqs)
    res = model.classify(data, off, threads=4)
    votes = model.votes(res, tr["anc"])
    assert (votes[:, 0] == 100).all()                                  # conf(root) = 1.0
    assert (np.diff(votes, axis=1) <= 0).all()                         # monotone non-increasing root -> genus
    # self-classification of training sequences returns their own genus
    assert (res["genus"] == tr["genus"]).mean() >= 0.95


def test_min_boot_words_parameter(toy):
    tr, seqs, model = toy
    data, off = pack_sequences([seqs[0][:60]])                         # 53 words -> k = 6; with min 5 still 6
    a = model.classify(data, off, 0)
    b = model.classify(data, off, 5)
    assert a["boot"].tolist() == b["boot"].tolist()
    data, off = pack_sequences([seqs[0][:50]])                         # 43 words -> k = 5 either way
    c = model.classify(data, off, 8)                                   # k forced to 8: another stream
    assert c["n_words"][0] == 43


@pytest.mark.parametrize("n,min_boot", [(486, 0), (243, 0), (37, 5), (9, 5), (640, 0)])
def test_bootstrap_is_a_count_matrix_product(n, min_boot):
    """What the tensor-core kernel (csrc/pg_mma.cu) rests on, checked on the CPU with the oracle's java.util.Random:
    every read re-seeds the generator with 1, so replicate t of a read with n words draws the same word POSITIONS in
    every read, and the 100 replicate sums of any integer column x are the product C_n x with C_n[t][j] = how often
    replicate t draws position j.  Split into two byte columns the product stays exact: 256 * (C hi) + (C lo)."""
    k = max(n // 8, min_boot)
    draws = ora.jrandom_stream(1, n, 100 * k).reshape(100, k)            # the oracle's stream: replicate by replicate
    C = np.zeros((101, n), np.int64)
    C[0] = 1                                                             # task 0: the full sum
    for t in range(100):
        np.add.at(C[1 + t], draws[t], 1)
    assert C[1:].sum(axis=1).tolist() == [k] * 100 and C.max() <= 255    # counts fit the u8 operand
    rng = np.random.default_rng(n)
    q = rng.integers(0, 4096, (n, 7))                                    # 12-bit deficits of 7 table positions
    by_draw = np.stack([q.sum(axis=0)] + [q[draws[t]].sum(axis=0) for t in range(100)])
    assert np.array_equal(C @ q, by_draw)
    lo, hi = q & 255, q >> 8
    assert np.array_equal(256 * (C @ hi) + (C @ lo), by_draw)
    # the byte bounds: rounded down and capped, four times their sum never exceeds the sum of the 16-bit minima
    b8 = np.minimum(q >> 2, 255)
    assert np.all(4 * (C @ b8) <= by_draw)
