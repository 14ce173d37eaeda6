"""GPU parity tests for Stage A: the CUDA path (through the C ABI) against the
CPU oracle on the same seeded inputs.  Integer / index results must be
bit-exact; the summed log-posterior is compared bit-for-bit as well (the
north-star tolerance is 1e-5 relative; strict mode does better)."""
import numpy as np
import pytest

import oracle_rdp as ora
import pangea_b200 as pg
from pangea_b200 import pack_sequences, synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def small(ctx):
    tr = synth.synth16s(seed=21, seqs=300, genera=70, length=900)      # G not a multiple of 32
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    yield tr, om, gm
    om.free()
    gm.free()


def test_counts_bit_exact(small):
    tr, om, gm = small
    m, nw, M, N = gm.counts()
    rm, rnw, rM, rN = om.counts()
    assert N == rN == len(tr["genus"])
    assert np.array_equal(M, rM)
    assert np.array_equal(nw, rnw)
    assert np.array_equal(m, rm)


def test_tables_bit_exact(small):
    tr, om, gm = small
    lp, ll, t = gm.tables()
    rlp, rll, rt = om.tables()
    assert np.array_equal(bits(lp), bits(rlp))
    assert np.array_equal(bits(ll), bits(rll))
    bad = np.count_nonzero(bits(t) != bits(rt))
    assert bad == 0, f"{bad} of {t.size} table entries differ"


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 64, 243, 486, 1000, 1418, 2048, 4097, 7000])
def test_boot_indices_match_java_random(ctx, n):
    got = ctx.boot_indices(n)
    k = n // 8
    want = ora.jrandom_stream(1, n, 100 * k).reshape(100, k) if k else np.zeros((100, 0), np.int32)
    assert np.array_equal(got.astype(np.int32), want)


def test_boot_indices_min_words(ctx):
    got = ctx.boot_indices(20, 5)
    want = ora.jrandom_stream(1, 20, 500).reshape(100, 5)
    assert np.array_equal(got.astype(np.int32), want)
    # back to the 2.5 rule: cache must be rebuilt, not reused
    got = ctx.boot_indices(20, 0)
    want = ora.jrandom_stream(1, 20, 200).reshape(100, 2)
    assert np.array_equal(got.astype(np.int32), want)


def edge_reads(tr):
    rng = np.random.default_rng(9)
    seqs = [tr["data"][tr["off"][i]:tr["off"][i + 1]] for i in range(40)]
    reads = []
    for i, s in enumerate(seqs):
        a = int(rng.integers(0, 300))
        ln = int(rng.integers(50, 600))
        r = s[a:a + ln].copy()
        if i % 2:
            r = synth.revcomp(r)
        if i % 5 == 0:
            r[rng.integers(0, len(r), 3)] = ord("N")
        if i % 7 == 0:
            r = np.frombuffer(r.tobytes().upper(), np.uint8)
        reads.append(r.tobytes())
    reads += [
        b"",                                   # empty record
        b"ACGT" * 12,                          # 48 bases: short (A2)
        b"ACGTTGCA" * 6 + b"AC",               # exactly 50 bases
        b"N" * 64,                             # long enough, no good word
        b"ACGTACGTAC" + b"N" * 50,             # 3 words: k = 0 replicates
        b"acgu" * 20,                          # RNA alphabet, lower case
        b"RYKMSWBDHVN" * 8,                    # IUPAC only
        seqs[0].tobytes(),                     # full training sequence
        synth.revcomp(seqs[1]).tobytes(),
    ]
    return reads


def check_against_oracle(ctx, gm, om, anc, reads, mode=0, min_boot=0):
    data, off = pack_sequences(reads)
    res, boot = ctx.classify(gm, data, off, mode=mode, min_boot_words=min_boot, want_boot=True)
    st = ctx.classify_stats()
    if mode == 1:
        assert gm.certifiable and st["certified"] > 0, st      # the certified kernels really ran
    else:
        assert st["certified"] == 0 and st["handed_back"] == 0
    ref = om.classify(data, off, min_boot)
    votes = om.votes(ref, anc)
    assert np.array_equal(res["status"], ref["status"])
    assert np.array_equal(res["genus"], ref["genus"])
    ok = ref["status"] == 0
    assert np.array_equal(res["n_words"][ok], ref["n_words"][ok])
    assert np.array_equal(res["reversed"][ok], ref["reversed"][ok])
    assert np.array_equal(boot[ok], ref["boot"][ok])
    assert np.array_equal(bits(res["score"][ok]), bits(ref["score"][ok]))
    d = anc.shape[1]
    assert np.array_equal(res["votes"][ok][:, :d].astype(np.int32), votes[ok])
    assert (res["votes"][:, d:] == 0).all()
    assert (res["depth"][ok] == d).all()
    return res


def test_pack_planes(ctx, small):
    tr, om, gm = small
    reads = edge_reads(tr)
    data, off = pack_sequences(reads)
    pk = ctx.pack(data, off)
    assert len(pk) == len(reads)
    code = {ord(c): v for c, v in zip("ATUGCatugc", [0, 1, 1, 2, 3, 0, 1, 1, 2, 3])}
    for i, r in enumerate(reads):
        ln, codes, mask = pk.unpack(i, len(r))
        assert ln == len(r)
        for p, ch in enumerate(r):
            valid = (mask[p // 32] >> (p % 32)) & 1
            assert valid == (1 if ch in code else 0)
            if valid:
                assert (codes[p // 16] >> (2 * (p % 16))) & 3 == code[ch]
    pk.free()


def test_extract_words_and_orientation(ctx, small):
    tr, om, gm = small
    reads = edge_reads(tr)
    data, off = pack_sequences(reads)
    words, nw, rev = ctx.extract_words(gm, data, off)
    ref = om.classify(data, off)
    L = ora.lib()
    for i, r in enumerate(reads):
        if len(r) < 50:
            assert nw[i] == 0
            continue
        w = ora.words(r)
        if ref[i]["reversed"]:
            w = np.array([L.rdp_revcomp_word(int(x)) for x in w[::-1]], np.int32)
        assert nw[i] == len(w) == ref[i]["n_words"]
        assert rev[i] == ref[i]["reversed"]
        assert np.array_equal(words[off[i]:off[i] + nw[i]].astype(np.int32), w)


MODES = pytest.mark.parametrize("mode", [0, 1], ids=["strict", "certified"])


@MODES
def test_classify_edge_cases(ctx, small, mode):
    tr, om, gm = small
    check_against_oracle(ctx, gm, om, tr["anc"], edge_reads(tr), mode=mode)


def test_classify_empty_batch(ctx, small):
    tr, om, gm = small
    res = ctx.classify(gm, np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(res) == 0


@MODES
def test_classify_min_boot_words(ctx, small, mode):
    tr, om, gm = small
    check_against_oracle(ctx, gm, om, tr["anc"], edge_reads(tr)[:12], min_boot=5, mode=mode)
    check_against_oracle(ctx, gm, om, tr["anc"], edge_reads(tr)[:12], min_boot=0, mode=mode)


@MODES
def test_classify_illumina_reads(ctx, small, mode):
    tr, om, gm = small
    for paired in (False, True):
        data, off, src = synth.synth_reads(31, tr, 400, paired=paired)
        reads = [data[off[i]:off[i + 1]].tobytes() for i in range(400)]
        res = check_against_oracle(ctx, gm, om, tr["anc"], reads, mode=mode)
        assert (res["n_words"] == (486 if paired else 243)).mean() > 0.2      # some carry an 'n'
        assert (res["genus"] == src).mean() > 0.5


@MODES
def test_classify_self_full_length(ctx, small, mode):
    """config 2 in miniature: classify the training set against its own model."""
    tr, om, gm = small
    reads = [tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes() for i in range(len(tr["genus"]))]
    res = check_against_oracle(ctx, gm, om, tr["anc"], reads, mode=mode)
    assert (res["genus"] == tr["genus"]).mean() >= 0.99


@MODES
def test_long_reads_every_bucket(ctx, mode):
    """reads of 1.4k .. 6.9k words exercise the 1-CTA/SM and narrow-tile launches
    (certified mode hands the narrow-tile buckets to the strict kernels)."""
    tr = synth.synth16s(seed=77, seqs=24, genera=9, length=7000)
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    seqs = [tr["data"][tr["off"][i]:tr["off"][i + 1]] for i in range(24)]
    reads = [s[:ln].tobytes() for s, ln in zip(seqs, [1400, 1790, 1830, 2500, 3590, 3650, 5000, 6300])]
    reads.append(synth.revcomp(seqs[8][:6999]).tobytes())
    reads.append(seqs[9][:300].tobytes())
    check_against_oracle(ctx, gm, om, tr["anc"], reads, mode=mode)
    too_long = np.tile(seqs[0], 2)[:7300].tobytes()
    data, off = pack_sequences([too_long])
    with pytest.raises(pg.PangeaError) as e:
        ctx.classify(gm, data, off)
    assert e.value.code == -6                                        # PG_ERANGE, not a silent truncation
    om.free()
    gm.free()


def test_single_genus_and_no_lineage(ctx):
    tr = synth.synth16s(seed=5, seqs=6, genera=1, length=300)
    om = ora.Model(tr["data"], tr["off"], tr["genus"], 1)
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], 1)
    reads = [tr["data"][tr["off"][i]:tr["off"][i + 1]][:200].tobytes() for i in range(6)]
    data, off = pack_sequences(reads)
    res, boot = ctx.classify(gm, data, off, want_boot=True)
    assert (res["genus"] == 0).all() and (boot == 0).all()
    assert (res["depth"] == 1).all() and (res["votes"][:, 0] == 100).all()
    ref = om.classify(data, off)
    assert np.array_equal(bits(res["score"]), bits(ref["score"]))
    om.free()
    gm.free()


def test_train_rejects_bad_genus(ctx):
    data, off = pack_sequences([b"ACGT" * 30, b"TTGA" * 30])
    with pytest.raises(pg.PangeaError):
        ctx.train(data, off, np.array([0, 5], np.int32), 2)


def test_model_save_load_roundtrip(ctx, small, tmp_path):
    tr, om, gm = small
    p = tmp_path / "m.pgm"
    gm.save(p, b"taxonomy-blob")
    gm2, blob = ctx.model_load(p)
    assert blob == b"taxonomy-blob"
    assert gm2.G == gm.G and gm2.N == gm.N
    a, b = gm.tables(), gm2.tables()
    for x, y in zip(a, b):
        assert np.array_equal(bits(x), bits(y))
    check_against_oracle(ctx, gm2, om, tr["anc"], edge_reads(tr)[:10])
    gm2.free()


def test_model_replication_buffers(ctx, small):
    """the multi-GPU path: copy the count buffers into an empty model, commit, same tables."""
    import ctypes as C

    tr, om, gm = small
    gm2 = ctx.model_create(gm.G)
    cudart = C.CDLL("libcudart.so.12")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    ctx.sync()
    for (src, n), (dst, n2) in zip(gm.buffers(), gm2.buffers()):
        assert n == n2
        assert cudart.cudaMemcpy(dst, src, n, 3) == 0                # cudaMemcpyDeviceToDevice
    gm2.commit()
    gm2.set_lineage(tr["anc"])
    assert gm2.N == gm.N
    for x, y in zip(gm.tables(), gm2.tables()):
        assert np.array_equal(bits(x), bits(y))
    gm2.free()


def test_real_16s_subset(ctx):
    """real type-strain 16S (subset of validation_dataset/rdp_download_373seqs.fa kept as a
    fixture): train with genus = 2nd header token, classify every sequence and windows of it."""
    from pathlib import Path

    fa = Path(__file__).parent / "golden" / "rdp_373_subset.fa"
    ids, hdr, seqs = synth.read_fasta(fa)
    names = sorted({h.split()[1] for h in hdr})
    gi = {n: i for i, n in enumerate(names)}
    genus = np.array([gi[h.split()[1]] for h in hdr], np.int32)
    data, off = pack_sequences(seqs)
    om = ora.Model(data, off, genus, len(names))
    gm = ctx.train(data, off, genus, len(names))
    anc = np.stack([np.zeros(len(names), np.int32), 1 + np.arange(len(names), dtype=np.int32)], axis=1)
    gm.set_lineage(anc)
    m, nw, M, N = gm.counts()
    rm, rnw, rM, rN = om.counts()
    assert np.array_equal(m, rm) and np.array_equal(nw, rnw) and np.array_equal(M, rM) and N == rN
    assert np.array_equal(bits(gm.tables()[2]), bits(om.tables()[2]))
    reads = list(seqs) + [s[100:350] for s in seqs] + [s[-400:] for s in seqs]
    check_against_oracle(ctx, gm, om, anc, reads, mode=0)
    check_against_oracle(ctx, gm, om, anc, reads, mode=1)
    om.free()
    gm.free()


def _twin_model(ctx, copies):
    """genera that are exact copies of one another: every sum ties, the reference's
    first-maximum rule decides, and the certified path sees near-ties everywhere."""
    base = synth.synth16s(seed=404, seqs=60, genera=12, length=600)
    G0 = base["G"]
    datas, offs, gen = [], [0], []
    for c in range(copies):
        for i in range(len(base["genus"])):
            s = base["data"][base["off"][i]:base["off"][i + 1]]
            datas.append(s)
            offs.append(offs[-1] + len(s))
            gen.append(base["genus"][i] + c * G0)
    data = np.concatenate(datas)
    off = np.array(offs, np.int64)
    genus = np.array(gen, np.int32)
    G = G0 * copies
    anc = np.concatenate([base["anc"]] * copies, axis=0).copy()
    anc[:, -1] = 1000 + np.arange(G)                     # distinct genus nodes, shared ancestors
    om = ora.Model(data, off, genus, G)
    gm = ctx.train(data, off, genus, G)
    gm.set_lineage(anc)
    reads = [base["data"][base["off"][i]:base["off"][i + 1]][40:40 + 260].tobytes() for i in range(30)]
    return om, gm, anc, reads


@pytest.mark.parametrize("copies", [2, 3, 7])
def test_certified_with_exact_ties(ctx, copies):
    """2 copies: the twin is a near-tie in every replicate; 3 and 7 copies overflow the
    near-tie list and exercise the hand-back to the strict kernels."""
    om, gm, anc, reads = _twin_model(ctx, copies)
    res = check_against_oracle(ctx, gm, om, anc, reads, mode=1)
    st = ctx.classify_stats()
    assert (st["handed_back"] > 0) == (copies > 2), st   # 3+ copies overflow the near-tie list
    assert (res["genus"] < 12).all()                     # lowest index among identical genera
    check_against_oracle(ctx, gm, om, anc, reads, mode=0)
    om.free()
    gm.free()


def test_certified_close_genera(ctx):
    """genera 0.4 % apart: many replicates have several survivors inside the margin."""
    tr = synth.synth16s(seed=99, seqs=400, genera=90, length=700)
    rng = np.random.default_rng(1)
    src = [tr["data"][tr["off"][i]:tr["off"][i + 1]].copy() for i in range(len(tr["genus"]))]
    proto = src[0][:630]
    seqs = []
    for i in range(len(src)):
        s = proto.copy()
        hit = rng.random(s.size) < 0.004
        s[hit] = synth.BASES[rng.integers(0, 4, int(hit.sum()))]
        seqs.append(s.tobytes())
    data, off = pack_sequences(seqs)
    om = ora.Model(data, off, tr["genus"], tr["G"])
    gm = ctx.train(data, off, tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    reads = [s[30:30 + 250] for s in seqs[:120]]
    check_against_oracle(ctx, gm, om, tr["anc"], reads, mode=1)
    om.free()
    gm.free()


# ------------------------------------------------------------------ BASELINE configs at their full sizes

@pytest.fixture(scope="module")
def baseline_model(ctx):
    """the bench's model: synth16s(0x9178, 9178 seqs, 1219 genera) -- the stand-in for the
    reference's absent rdp_download_9178seqs.fa"""
    tr = synth.synth16s(0x9178, 9178, 1219)
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    yield tr, om, gm
    om.free()
    gm.free()


def test_tables_bit_exact_full_size_model(ctx, baseline_model):
    """the device's double-precision log() against glibc's on every cell of the 1 219-genus bench model (80 M cells):
    A6 narrows ln() of a float quotient to fp32, so the two libraries would have to disagree in the last bits of the
    double AND sit on a rounding boundary of the float to differ; the whole table is compared bit for bit"""
    tr, om, gm = baseline_model
    lp, ll, t = gm.tables()
    rlp, rll, rt = om.tables()
    assert np.array_equal(bits(lp), bits(rlp)) and np.array_equal(bits(ll), bits(rll))
    assert np.array_equal(bits(t), bits(np.ascontiguousarray(rt)))


def test_config0_real_16s_queries_vs_9178_model(ctx, baseline_model):
    """configs[0]: real full-length 16S queries (subset of rdp_download_373seqs.fa) against the 9178-seq model"""
    from pathlib import Path

    tr, om, gm = baseline_model
    ids, hdr, seqs = synth.read_fasta(Path(__file__).parent / "golden" / "rdp_373_subset.fa")
    for mode in (0, 1):
        check_against_oracle(ctx, gm, om, tr["anc"], seqs, mode=mode)


def test_config1_self_classification_full_size(ctx, baseline_model):
    """configs[1]: all 9178 training sequences against their own model.  Oracle parity on a
    600-read sample; on the full set strict == certified record for record, >= 99 % self-genus,
    votes(root) = 100 and monotone non-increasing down the lineage."""
    tr, om, gm = baseline_model
    n = len(tr["genus"])
    sample = [tr["data"][tr["off"][i]:tr["off"][i + 1]].tobytes() for i in range(0, n, n // 600)]
    check_against_oracle(ctx, gm, om, tr["anc"], sample, mode=1)
    a, ba = ctx.classify(gm, tr["data"], tr["off"], mode=0, want_boot=True)
    b, bb = ctx.classify(gm, tr["data"], tr["off"], mode=1, want_boot=True)
    assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb)
    assert (a["genus"] == tr["genus"]).mean() >= 0.99
    d = tr["anc"].shape[1]
    assert (a["votes"][:, 0] == 100).all() and (np.diff(a["votes"][:, :d].astype(int), axis=1) <= 0).all()


def test_config2_illumina_reads_full_size(ctx, baseline_model):
    """configs[2]: 2^18 joined 250 bp pairs (the bench's step).  Oracle parity on a sample;
    strict == certified on every record; checksum of the records is order-stable."""
    from pangea_b200 import dist as pgdist

    tr, om, gm = baseline_model
    data, off, src = synth.synth_reads(0x250, tr, 1 << 18, paired=True)
    reads = [data[off[i]:off[i + 1]].tobytes() for i in range(0, 1 << 18, 257)]
    check_against_oracle(ctx, gm, om, tr["anc"], reads, mode=1)
    a, ba = ctx.classify(gm, data, off, mode=0, want_boot=True)
    b, bb = ctx.classify(gm, data, off, mode=1, want_boot=True)
    st = ctx.classify_stats()
    assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb)
    assert st["certified"] == 1 << 18
    assert (a["genus"] == src).mean() > 0.95
    assert pgdist.records_checksum(a) == pgdist.records_checksum(b)
    # idempotence: classifying the same batch again gives the same records
    c = ctx.classify(gm, data, off, mode=1)
    assert c.tobytes() == b.tobytes()


def test_config3_rdp_scale_genera(ctx):
    """configs[3] in its genus dimension: a 10 000-genus model (157 genus blocks of 64, 313 table tiles).
    Oracle parity on a sample in both modes; strict == certified on a larger batch."""
    tr = synth.synth16s(0x3000000, 30000, 10000, length=600)
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    assert gm.certifiable
    m, nw, M, N = gm.counts(dense=False)
    rm, rnw, rM, rN = om.counts()
    assert np.array_equal(nw, rnw) and np.array_equal(M, rM) and N == rN
    lp, ll, t = gm.tables()                                  # 655 M cells: device log() == glibc log() after narrowing
    rlp, rll, rt = om.tables()
    assert np.array_equal(bits(lp), bits(rlp)) and np.array_equal(bits(ll), bits(rll))
    assert np.array_equal(bits(t), bits(np.ascontiguousarray(rt)))
    del t, rt
    data, off, src = synth.synth_reads(0x3000001, tr, 4096, paired=False)
    reads = [data[off[i]:off[i + 1]].tobytes() for i in range(0, 4096, 32)]
    check_against_oracle(ctx, gm, om, tr["anc"], reads, mode=1)
    check_against_oracle(ctx, gm, om, tr["anc"], reads[:32], mode=0)
    a, ba = ctx.classify(gm, data, off, mode=0, want_boot=True)
    b, bb = ctx.classify(gm, data, off, mode=1, want_boot=True)
    assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb)
    # joined pairs (two block groups of bounds per read and ~490 words): the coarse first level (default here) and
    # the 16-bit bounds must both return the strict kernels' records, and a sample must match the oracle
    data, off, src = synth.synth_reads(0x3000002, tr, 3000, paired=True, gap=40)
    a, ba = ctx.classify(gm, data, off, mode=0, want_boot=True)
    for kw in (dict(), dict(bound_level=1), dict(bound_level=2, light_max=40), dict(bound_level=3)):
        b, bb = ctx.classify(gm, data, off, mode=1, want_boot=True, **kw)
        st = ctx.classify_stats()
        assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb), kw
        assert st["certified"] == 3000, st
    ref = om.classify(data[: off[100]], off[:101])
    assert np.array_equal(a["genus"][:100], ref["genus"]) and np.array_equal(ba[:100], ref["boot"])
    assert np.array_equal(bits(a["score"][:100]), bits(ref["score"]))
    om.free()
    gm.free()


# ------------------------------------------------------------------ certified plans (block lower bounds)

def test_certified_plans_agree(ctx, baseline_model):
    """The certified path's work plans differ only in how they prove genera irrelevant: the default plan
    (best block + block lower bounds + items), the same with every read with an open item sent back to the
    all-block kernel (light_max = -1), and the all-block kernel alone (cert_plan = 1) must return the strict
    kernels' records byte for byte."""
    tr, om, gm = baseline_model
    data, off, src = synth.synth_reads(0x251, tr, 6000, paired=True)
    want, wb = ctx.classify(gm, data, off, mode=0, want_boot=True)
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True)
    st = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st["certified"] == 6000 and st["items"] > 0, st
    assert st["heavy"] < 600, st                         # the bound dismisses nearly every block on this workload
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, light_max=-1)
    st = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st["heavy"] > 0 and st["items"] == 0, st
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, light_max=3)
    st = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st["heavy"] > 0 and st["items"] > 0, st
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, cert_plan=1)
    st = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st["heavy"] == 0 and st["items"] == 0, st
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, cert_plan=2)     # whole best block, one read per CTA
    st = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st["items"] > 0, st
    # the coarse 8-bit first level + exact second level (the default of models with more than one block group)
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, bound_level=2)
    st = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st["certified"] == 6000 and st["items"] > 0 and st["heavy"] < 600, st


def test_certified_without_lineage(ctx):
    """no lineage -> the quantised table keeps the genus order of the training file (relatives scattered
    over the blocks, weak bounds, many items); results are the same and match the oracle."""
    tr = synth.synth16s(seed=31, seqs=900, genera=300, length=900)
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    data, off, src = synth.synth_reads(5, tr, 1500, paired=True)
    a, ba = ctx.classify(gm, data, off, mode=0, want_boot=True)
    b, bb = ctx.classify(gm, data, off, mode=1, want_boot=True)
    assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb)
    ref = om.classify(data[: off[200]], off[:201])
    assert np.array_equal(b["genus"][:200], ref["genus"]) and np.array_equal(bb[:200], ref["boot"])
    gm.set_lineage(tr["anc"])                            # re-lays the table out; results must not move
    c, bc = ctx.classify(gm, data, off, mode=1, want_boot=True)
    assert np.array_equal(c["genus"], b["genus"]) and np.array_equal(bc, bb)
    assert np.array_equal(bits(c["score"]), bits(b["score"]))
    om.free()
    gm.free()


@pytest.mark.parametrize("seed,genera,seqs,length", [(11, 70, 150, 500), (12, 130, 500, 1200), (13, 700, 1400, 800),
                                                     (14, 2150, 4300, 420)])
def test_certified_equals_strict_on_random_models(ctx, seed, genera, seqs, length):
    """random taxonomies of 2 .. 40 table blocks (2 150 genera = two block groups in k_bound), reads of every
    length bucket (single mates, joined pairs, long fragments, fragments with runs of N): the certified plans and
    the strict kernels must agree byte for byte, and a sample must agree with the oracle."""
    tr = synth.synth16s(seed=seed, seqs=seqs, genera=genera, length=length)
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    assert gm.certifiable
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(1200):
        s = tr["data"][tr["off"][i % seqs]:tr["off"][i % seqs + 1]]
        ln = int(rng.choice([60, 120, 250, min(700, len(s)), len(s)]))
        a = int(rng.integers(0, len(s) - ln + 1))
        r = s[a:a + ln].copy()
        hit = rng.random(ln) < 0.01
        r[hit] = synth.BASES[rng.integers(0, 4, int(hit.sum()))]
        if i % 9 == 0 and ln > 80:
            r[30:30 + int(rng.integers(1, 40))] = ord("N")
        if i % 2:
            r = synth.revcomp(r)
        reads.append(r.tobytes())
    data, off = pack_sequences(reads)
    want, wb = ctx.classify(gm, data, off, mode=0, want_boot=True)
    for kw in (dict(), dict(cert_plan=1), dict(cert_plan=2), dict(light_max=5), dict(bound_level=1), dict(bound_level=2),
               dict(bound_level=2, light_max=5), dict(bound_level=3), dict(bound_level=3, light_max=5)):
        got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, **kw)
        assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb), kw
    om = ora.Model(tr["data"], tr["off"], tr["genus"], tr["G"])
    ref = om.classify(data[: off[150]], off[:151])
    assert np.array_equal(want["genus"][:150], ref["genus"]) and np.array_equal(wb[:150], ref["boot"])
    om.free()
    gm.free()


def test_large_batches_are_sliced(ctx, small, monkeypatch):
    """pg_classify cuts a batch into slices so that device scratch stays bounded; slices of 700 reads
    (PG_SLICE_READS) of a 3 000-read batch must give the records of one pass, statistics summed."""
    tr, om, gm = small
    data, off, _ = synth.synth_reads(21, tr, 3000, paired=False, read_len=200)
    want, wb = ctx.classify(gm, data, off, mode=1, want_boot=True)
    st0 = ctx.classify_stats()
    monkeypatch.setenv("PG_SLICE_READS", "700")
    got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True)
    st1 = ctx.classify_stats()
    assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb)
    assert st1["certified"] + st1["strict"] == 3000 == st0["certified"] + st0["strict"]


def test_model_from_tables_round_trip(ctx, small):
    """a model exported as RDP-style tables (word priors, leave counts, the sparse word-major
    list of conditional log probabilities of the cells with a count) and built again from them holds the same dense
    table bit for bit -- the absent cells are fp32(logPrior - logLeave) on both sides -- and classifies to the same
    records; it has no counts, so it refuses to be saved or committed."""
    tr, om, gm = small
    m, nw, M, N = gm.counts()
    lp, ll, t = gm.tables()
    present = m > 0
    idx = np.zeros(65537, np.int64)
    idx[1:] = np.cumsum(present.sum(axis=1))
    w_, g_ = np.nonzero(present)                           # row-major: word-major, genus ascending
    tm = ctx.model_from_tables(lp, M, idx, g_.astype(np.int32), t[w_, g_])
    lp2, ll2, t2 = tm.tables()
    assert np.array_equal(bits(lp2), bits(lp)) and np.array_equal(bits(ll2), bits(ll)) and np.array_equal(bits(t2), bits(t))
    tm.set_lineage(tr["anc"])
    assert tm.certifiable
    data, off, src = synth.synth_reads(77, tr, 500, paired=True)
    for mode in (0, 1):
        a, ba = ctx.classify(gm, data, off, mode=mode, want_boot=True)
        b, bb = ctx.classify(tm, data, off, mode=mode, want_boot=True)
        assert a.tobytes() == b.tobytes() and np.array_equal(ba, bb)
    with pytest.raises(pg.PangeaError):
        tm.commit()
    with pytest.raises(pg.PangeaError):
        tm.save("/tmp/never.pgm")
    # malformed input is refused, not scattered
    bad = idx.copy()
    bad[100] = bad[101] + 5
    with pytest.raises(pg.PangeaError):
        ctx.model_from_tables(lp, M, bad, g_.astype(np.int32), t[w_, g_])
    gbad = g_.astype(np.int32).copy()
    gbad[0] = tr["G"]
    with pytest.raises(pg.PangeaError):
        ctx.model_from_tables(lp, M, idx, gbad, t[w_, g_])
    tm.free()


@pytest.mark.parametrize("seed", [5, 6])
def test_ties_without_lineage_and_min_boot_words(ctx, seed):
    """everything awkward at once: genera that are exact copies of each other (ties in the full sum and in every
    replicate), a few near-copies, NO lineage (the table keeps the training order, so twins sit in different blocks
    and the bounds cannot dismiss them), min_boot_words = 5 with reads short enough for it to matter (k = max(n/8, 5)),
    reads with runs of N, both strands.  Certified (every plan) == strict == oracle."""
    rng = np.random.default_rng(seed)
    base = synth.synth16s(seed=40 + seed, seqs=150, genera=150, length=420)
    seqs, genus = [], []
    for i in range(150):
        s = base["data"][base["off"][i]:base["off"][i + 1]]
        seqs.append(s.tobytes())
        genus.append(int(base["genus"][i]))
    G = 150
    for c in range(60):                                      # exact twins of the first 60 sequences, as new genera
        seqs.append(seqs[c])
        genus.append(G)
        G += 1
    for c in range(30):                                      # near copies: two substitutions
        s = np.frombuffer(seqs[c], np.uint8).copy()
        for p in rng.integers(0, len(s), 2):
            s[p] = synth.BASES[rng.integers(0, 4)]
        seqs.append(s.tobytes())
        genus.append(G)
        G += 1
    data, off = pack_sequences(seqs)
    genus = np.array(genus, np.int32)
    om = ora.Model(data, off, genus, G)
    gm = ctx.train(data, off, genus, G)                      # no set_lineage
    assert gm.certifiable
    reads = []
    for i in range(400):
        s = np.frombuffer(seqs[int(rng.integers(0, len(seqs)))], np.uint8)
        ln = int(rng.choice([50, 52, 55, 60, 90, 250, len(s)]))
        a = int(rng.integers(0, len(s) - ln + 1))
        r = s[a:a + ln].copy()
        if i % 7 == 0 and ln >= 90:
            r[20:20 + int(rng.integers(1, 30))] = ord("N")
        if i % 2:
            r = synth.revcomp(r)
        reads.append(r.tobytes())
    rdata, roff = pack_sequences(reads)
    ref = om.classify(rdata, roff, 5)
    want, wb = ctx.classify(gm, rdata, roff, mode=0, min_boot_words=5, want_boot=True)
    ok = ref["status"] == 0
    assert np.array_equal(want["genus"], ref["genus"]) and np.array_equal(wb[ok], ref["boot"][ok])
    assert np.array_equal(bits(want["score"][ok]), bits(ref["score"][ok]))
    assert (want["votes"][ok][:, 0].astype(int) == (wb[ok] == want["genus"][ok, None]).sum(axis=1)).all()   # one-level lineage {g}
    assert ((want["n_words"][ok] // 8) < 5).any()            # min_boot_words really decided k for some reads
    for kw in (dict(), dict(cert_plan=1), dict(cert_plan=2), dict(light_max=3), dict(bound_level=2), dict(bound_level=3)):
        got, gb = ctx.classify(gm, rdata, roff, mode=1, min_boot_words=5, want_boot=True, **kw)
        assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb), kw
    om.free()
    gm.free()


def test_tensor_core_plan_every_read_length(ctx):
    """Plan 4 (the default of certified mode): best part and block bounds as one integer matrix product per read on the
    tensor cores.  Reads of every length from 9 to 640 words -- every count image, every number of ring stages, partial
    last K steps, a new image at nearly every read -- and a few longer ones (plan 3 takes those) must come back with the
    strict kernels' records and replicate winners, for k = n/8 and for min_boot_words = 5."""
    tr = synth.synth16s(seed=0x44A, seqs=900, genera=300, length=1500)          # 300 genera: 5+ blocks, clade-aligned padding
    gm = ctx.train(tr["data"], tr["off"], tr["genus"], tr["G"])
    gm.set_lineage(tr["anc"])
    rng = np.random.default_rng(0x44A)
    reads = []
    for ln in list(range(16, 648)) + [700, 900, 1200]:                        # ln bases -> ln - 7 words
        i = int(rng.integers(0, 900))
        s = tr["data"][tr["off"][i]:tr["off"][i + 1]]
        p = int(rng.integers(0, len(s) - ln))
        r = s[p:p + ln].copy()
        flip = rng.random(ln) < 0.02
        r[flip] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(flip.sum()))]
        reads.append((synth.revcomp(r) if rng.random() < 0.5 else r).tobytes())
    order = rng.permutation(len(reads))
    data, off = pack_sequences([reads[i] for i in order])
    for mb in (0, 5):
        want, wb = ctx.classify(gm, data, off, mode=0, want_boot=True, min_boot_words=mb)
        got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, min_boot_words=mb)
        st = ctx.classify_stats()
        assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb), mb
        assert st["tensor_core"] == 632 and st["certified"] == len(reads), st     # every read of up to 640 words (647 bases)
        got3, gb3 = ctx.classify(gm, data, off, mode=1, want_boot=True, min_boot_words=mb, cert_plan=3)
        st3 = ctx.classify_stats()
        assert got3.tobytes() == want.tobytes() and np.array_equal(gb3, wb) and st3["tensor_core"] == 0, (mb, st3)
    gm.free()


def test_tensor_core_plan_heavy_and_full_lists(ctx, baseline_model):
    """Plan 4's escape routes: every read with an open pair sent to the all-block kernel (light_max = -1), a small
    budget (light_max = 3), and reads far from every genus (their lists of open pairs and near-ties overflow)."""
    tr, om, gm = baseline_model
    data, off, src = synth.synth_reads(0x252, tr, 3000, paired=True)
    rng = np.random.default_rng(5)
    junk = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 500)].tobytes() for _ in range(40)]
    jd, jo = pack_sequences(junk)
    data = np.concatenate([data, jd])
    off = np.concatenate([off, off[-1] + jo[1:]])
    want, wb = ctx.classify(gm, data, off, mode=0, want_boot=True)
    for kw in (dict(), dict(light_max=-1), dict(light_max=3)):
        got, gb = ctx.classify(gm, data, off, mode=1, want_boot=True, **kw)
        st = ctx.classify_stats()
        assert got.tobytes() == want.tobytes() and np.array_equal(gb, wb), kw
        assert st["tensor_core"] == 3040, (kw, st)
        if kw:
            assert st["heavy"] > 0, (kw, st)
