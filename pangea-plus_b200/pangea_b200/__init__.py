"""ctypes binding of libpangea_b200.so (include/pangea_b200.h).

This is plumbing for the tests and bench.py; the product is the C-ABI library
and the C command-line tools under pangea-plus_b200/host/.  There is no Python
or CPU implementation of the hot path here: if the library is missing, or no
B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent          # pangea-plus_b200/
REPO = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libpangea_b200.so"
HEADER = REPO / "include" / "pangea_b200.h"

PG_NWORDS = 65536
PG_NUM_BOOT = 100
PG_MAX_DEPTH = 32
PG_MIN_SEQ_LEN = 50

RESULT_DTYPE = np.dtype(
    [
        ("genus", "<i4"),
        ("n_words", "<i4"),
        ("score", "<f4"),
        ("reversed", "u1"),
        ("status", "u1"),
        ("depth", "u1"),
        ("_pad0", "u1"),
        ("votes", "u1", (PG_MAX_DEPTH,)),
        ("_pad1", "u1", (16,)),
    ]
)
assert RESULT_DTYPE.itemsize == 64


class PangeaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpangea_b200 error {code}: {msg}")
        self.code = code


class _SeqBatch(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("off", C.c_void_p), ("count", C.c_int64)]


class _Opts(C.Structure):
    _fields_ = [("min_boot_words", C.c_int32), ("mode", C.c_int32), ("cert_plan", C.c_int32),
                ("light_max", C.c_int32), ("bound_level", C.c_int32), ("reserved", C.c_int32 * 3)]


class _TrimOpts(C.Structure):
    _fields_ = [("gap", C.c_int32), ("truncate", C.c_int32), ("reserved", C.c_int32 * 6)]


class _MegaclustOpts(C.Structure):
    _fields_ = [("sim_threshold", C.c_double), ("eval_threshold", C.c_double), ("bitscore_threshold", C.c_double),
                ("count_every_hit", C.c_int32), ("reserved", C.c_int32 * 5)]


class _ConsensusIn(C.Structure):
    _fields_ = [
        ("nreads", C.c_int64),
        ("hit_off", C.c_void_p),
        ("lineage_bytes", C.c_void_p),
        ("lineage_off", C.c_void_p),
        ("pident_bytes", C.c_void_p),
        ("pident_off", C.c_void_p),
        ("rdp_bytes", C.c_void_p),
        ("rdp_off", C.c_void_p),
        ("first_is_fresh", C.c_int32),
        ("reserved", C.c_int32),
    ]


def build(verbose: bool = False) -> Path:
    """Compile the library in-tree (nvcc, sm_100a).  Used by __graft_entry__.build()."""
    out = subprocess.run(["make", "-C", str(PKG_DIR), "-j8", "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libpangea_b200.so failed")
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: run `make -C {PKG_DIR}` (or __graft_entry__.build()). "
            "There is no fallback implementation."
        )
    lib = C.CDLL(str(LIB_PATH))
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.pg_init.restype = vp
    lib.pg_init.argtypes = [C.c_int]
    lib.pg_shutdown.argtypes = [vp]
    lib.pg_shutdown.restype = None
    lib.pg_last_error.restype = C.c_char_p
    lib.pg_last_error.argtypes = [vp]
    lib.pg_set_stream.argtypes = [vp, vp]
    lib.pg_sync.argtypes = [vp]
    lib.pg_launch_count.restype = i64
    lib.pg_launch_count.argtypes = [vp]
    lib.pg_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.pg_kernel_time_reset.argtypes = [vp]
    lib.pg_train.argtypes = [vp, C.POINTER(_SeqBatch), vp, C.c_int, C.POINTER(vp)]
    lib.pg_train_dev.argtypes = [vp, C.POINTER(_SeqBatch), vp, C.c_int, C.POINTER(vp)]
    lib.pg_model_free.argtypes = [vp]
    lib.pg_model_free.restype = None
    lib.pg_train_accumulate.argtypes = [vp, vp, C.POINTER(_SeqBatch), vp]
    lib.pg_train_accumulate_dev.argtypes = [vp, vp, C.POINTER(_SeqBatch), vp]
    u64 = C.c_uint64
    lib.pg_synth_members.argtypes = [vp, u64, vp, C.c_int, vp, vp, i64, i64, vp]
    lib.pg_synth_reads.argtypes = [vp, u64, vp, vp, vp, i64, i64, i64, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.pg_model_set_lineage.argtypes = [vp, vp, C.c_int]
    lib.pg_model_genera.argtypes = [vp]
    lib.pg_model_certifiable.argtypes = [vp]
    lib.pg_model_bound_columns.argtypes = [vp]
    lib.pg_classify_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.pg_classify_stats2.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.pg_classify_stats3.argtypes = [vp, C.POINTER(i64)]
    lib.pg_model_sequences.restype = i64
    lib.pg_model_sequences.argtypes = [vp]
    lib.pg_model_counts.argtypes = [vp, vp, vp, vp, C.POINTER(i64)]
    lib.pg_model_tables.argtypes = [vp, vp, vp, vp]
    lib.pg_model_save.argtypes = [vp, C.c_char_p, vp, i64]
    lib.pg_model_load.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    lib.pg_free.argtypes = [vp]
    lib.pg_free.restype = None
    lib.pg_model_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    lib.pg_model_from_tables.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.POINTER(vp)]
    lib.pg_model_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_int)]
    lib.pg_model_commit.argtypes = [vp]
    lib.pg_reads_pack.argtypes = [vp, C.POINTER(_SeqBatch), C.POINTER(vp)]
    lib.pg_reads_pack_dev.argtypes = [vp, C.POINTER(_SeqBatch), i64, C.POINTER(vp)]
    lib.pg_reads_free.argtypes = [vp]
    lib.pg_reads_free.restype = None
    lib.pg_reads_count.restype = i64
    lib.pg_reads_count.argtypes = [vp]
    lib.pg_reads_unpack.restype = i64
    lib.pg_reads_unpack.argtypes = [vp, i64, vp, vp, i64]
    lib.pg_classify.argtypes = [vp, vp, C.POINTER(_SeqBatch), C.POINTER(_Opts), vp, vp]
    lib.pg_classify_packed.argtypes = [vp, vp, vp, C.POINTER(_Opts), vp, vp]
    lib.pg_extract_words.argtypes = [vp, vp, C.POINTER(_SeqBatch), vp, vp, vp]
    lib.pg_boot_indices.argtypes = [vp, i32, i32, vp]
    lib.pg_tax_build.argtypes = [C.c_char_p]
    lib.pg_tax_load.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.pg_tax_free.argtypes = [vp]
    lib.pg_tax_free.restype = None
    lib.pg_tax_leaf.argtypes = [vp, vp, vp, i64, vp]
    lib.pg_tax_lineage.argtypes = [vp, vp, vp, i64, vp, i64, vp]
    lib.pg_consensus.argtypes = [vp, C.POINTER(_ConsensusIn), vp, vp]
    lib.pg_fasta_ingest.argtypes = [vp, C.c_char_p, i64, i64, C.POINTER(i64), vp, vp, vp, C.POINTER(vp)]
    lib.pg_fasta_last_bytes.argtypes = [vp, i64, vp, i64, vp]
    lib.pg_classify_packed_host.argtypes = [vp, vp, vp, C.POINTER(_Opts), vp, vp]
    lib.pg_first_hits.argtypes = [vp, C.c_char_p, i64, vp, i64, C.POINTER(i64), vp, i64, C.POINTER(i64)]
    lib.pg_megaclust.argtypes = [vp, C.c_char_p, i64, C.POINTER(_MegaclustOpts), i64, C.POINTER(i64), vp, vp, vp,
                                 C.POINTER(i64), C.POINTER(i64)]
    lib.pg_trim_join.argtypes = [vp, C.c_char_p, i64, C.c_char_p, i64, C.c_int, C.POINTER(_TrimOpts), vp, i64, C.POINTER(i64), C.POINTER(vp)]
    lib.pg_tax_chain.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.pg_tax_node_record.argtypes = [vp, i32, vp]
    lib.pg_tax_name_records.argtypes = [vp, i32, vp, i32]
    lib.pg_tax_max_gi.restype = i64
    lib.pg_tax_max_gi.argtypes = [vp]
    lib.pg_tax_max_taxid.restype = i64
    lib.pg_tax_max_taxid.argtypes = [vp]
    _lib = lib
    return lib


def pack_sequences(seqs) -> tuple[np.ndarray, np.ndarray]:
    """list of bytes/str -> (uint8 bytes, int64 offsets[count+1])."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return data, off


def _ptr(a) -> int:
    """address of a numpy array (host) or anything with data_ptr() (torch, device)."""
    if a is None:
        return 0
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class Model:
    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.h = ctx, handle

    @property
    def G(self) -> int:
        return self.ctx.lib.pg_model_genera(self.h)

    @property
    def N(self) -> int:
        return self.ctx.lib.pg_model_sequences(self.h)

    @property
    def certifiable(self) -> bool:
        return bool(self.ctx.lib.pg_model_certifiable(self.h))

    @property
    def bound_columns(self) -> int:
        """0 = block columns, 1 = part columns, -1 = not tuned yet (see pg_model_bound_columns)"""
        return int(self.ctx.lib.pg_model_bound_columns(self.h))

    def set_lineage(self, anc: np.ndarray) -> None:
        anc = np.ascontiguousarray(anc, dtype=np.int32)
        assert anc.ndim == 2 and anc.shape[0] == self.G
        self.ctx._chk(self.ctx.lib.pg_model_set_lineage(self.h, anc.ctypes.data, anc.shape[1]))

    def counts(self, dense: bool = True):
        G = self.G
        m = np.empty((PG_NWORDS, G), np.int32) if dense else None
        nw = np.empty(PG_NWORDS, np.int32)
        M = np.empty(G, np.int32)
        N = C.c_int64()
        self.ctx._chk(self.ctx.lib.pg_model_counts(self.h, _ptr(m), nw.ctypes.data, M.ctypes.data, C.byref(N)))
        return m, nw, M, N.value

    def tables(self, dense: bool = True):
        G = self.G
        lp = np.empty(PG_NWORDS, np.float32)
        ll = np.empty(G, np.float32)
        t = np.empty((PG_NWORDS, G), np.float32) if dense else None
        self.ctx._chk(self.ctx.lib.pg_model_tables(self.h, lp.ctypes.data, ll.ctypes.data, _ptr(t)))
        return lp, ll, t

    def save(self, path: str, blob: bytes = b"") -> None:
        buf = C.create_string_buffer(blob, len(blob)) if blob else None
        self.ctx._chk(self.ctx.lib.pg_model_save(self.h, str(path).encode(), C.cast(buf, C.c_void_p) if buf else None, len(blob)))

    def buffers(self):
        ptrs = (C.c_void_p * 8)()
        sizes = (C.c_size_t * 8)()
        n = C.c_int()
        self.ctx._chk(self.ctx.lib.pg_model_buffers(self.h, ptrs, sizes, 8, C.byref(n)))
        return [(ptrs[i], sizes[i]) for i in range(n.value)]

    def commit(self) -> None:
        self.ctx._chk(self.ctx.lib.pg_model_commit(self.h))

    def free(self) -> None:
        if self.h:
            self.ctx.lib.pg_model_free(self.h)
            self.h = 0


class Reads:
    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.h = ctx, handle

    def __len__(self) -> int:
        return self.ctx.lib.pg_reads_count(self.h)

    def unpack(self, i: int, length_hint: int):
        nch = (length_hint + 31) // 32 + 1
        codes = np.zeros(2 * nch, np.uint32)
        mask = np.zeros(nch, np.uint32)
        ln = self.ctx.lib.pg_reads_unpack(self.h, i, codes.ctypes.data, mask.ctypes.data, 2 * nch)
        if ln < 0:
            raise PangeaError(ln, self.ctx.last_error())
        return ln, codes, mask

    def free(self) -> None:
        if self.h:
            self.ctx.lib.pg_reads_free(self.h)
            self.h = 0


class Context:
    """One pg_ctx == one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = self.lib.pg_init(device)
        if not self.h:
            raise PangeaError(-5, self.lib.pg_last_error(None).decode())

    def last_error(self) -> str:
        return self.lib.pg_last_error(self.h).decode()

    def _chk(self, rc: int) -> None:
        if rc != 0:
            raise PangeaError(rc, self.last_error())

    def close(self) -> None:
        if self.h:
            self.lib.pg_shutdown(self.h)
            self.h = 0

    def set_stream(self, stream_ptr: int) -> None:
        self._chk(self.lib.pg_set_stream(self.h, stream_ptr))

    def sync(self) -> None:
        self._chk(self.lib.pg_sync(self.h))

    def launch_count(self) -> int:
        return self.lib.pg_launch_count(self.h)

    def kernel_time(self):
        ms, n = C.c_double(), C.c_int64()
        self._chk(self.lib.pg_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def classify_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._chk(self.lib.pg_classify_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        d, e = C.c_int64(), C.c_int64()
        self._chk(self.lib.pg_classify_stats2(self.h, C.byref(d), C.byref(e)))
        f = C.c_int64()
        self._chk(self.lib.pg_classify_stats3(self.h, C.byref(f)))
        return {"certified": a.value, "strict": b.value, "handed_back": c.value, "heavy": d.value, "items": e.value,
                "tensor_core": f.value}

    def kernel_time_reset(self) -> None:
        self._chk(self.lib.pg_kernel_time_reset(self.h))

    # ---- Stage A
    def train(self, data, off, genus, G: int, device: bool = False) -> Model:
        sb = _SeqBatch(_ptr(data), _ptr(off), len(off) - 1)
        h = C.c_void_p()
        fn = self.lib.pg_train_dev if device else self.lib.pg_train
        if not device:
            genus = np.ascontiguousarray(genus, dtype=np.int32)
        self._chk(fn(self.h, C.byref(sb), _ptr(genus), G, C.byref(h)))
        return Model(self, h.value)

    def train_accumulate(self, model: Model, data, off, genus, device: bool = False) -> None:
        """add the counts of one more batch to `model` (sharded / streamed training); commit() derives the tables"""
        sb = _SeqBatch(_ptr(data), _ptr(off), len(off) - 1)
        if not device:
            genus = np.ascontiguousarray(genus, dtype=np.int32)
        fn = self.lib.pg_train_accumulate_dev if device else self.lib.pg_train_accumulate
        self._chk(fn(self.h, model.h, C.byref(sb), _ptr(genus)))

    # ---- synthetic workloads generated on the device (include/pangea_b200_synth.h; numpy twins in synth.py)
    def synth_members(self, seed: int, centroids_dev, length: int, genus_dev, off_dev, first: int, count: int, bytes_dev) -> None:
        self._chk(self.lib.pg_synth_members(self.h, seed, _ptr(centroids_dev), length, _ptr(genus_dev), _ptr(off_dev), first, count,
                                            _ptr(bytes_dev)))

    def synth_reads(self, seed: int, members_dev, member_off_dev, member_genus_dev, nmembers: int, first: int, count: int,
                    out_dev, src_dev=None, read_len: int = 250, gap: int = 189, paired: bool = True) -> None:
        self._chk(self.lib.pg_synth_reads(self.h, seed, _ptr(members_dev), _ptr(member_off_dev), _ptr(member_genus_dev), nmembers,
                                          first, count, read_len, gap, int(paired), _ptr(out_dev), _ptr(src_dev)))

    def model_from_tables(self, log_prior, leave_count, idx, genus_of_entry, logp_of_entry) -> Model:
        """a model given as RDP-style tables (pg_model_from_tables): logPrior[65536], leaveCount[G], sparse word-major cells"""
        lp = np.ascontiguousarray(log_prior, np.float32)
        lc = np.ascontiguousarray(leave_count, np.int32)
        ix = np.ascontiguousarray(idx, np.int64)
        eg = np.ascontiguousarray(genus_of_entry, np.int32)
        ep = np.ascontiguousarray(logp_of_entry, np.float32)
        assert lp.size == PG_NWORDS and ix.size == PG_NWORDS + 1 and eg.size == ep.size == ix[-1]
        h = C.c_void_p()
        self._chk(self.lib.pg_model_from_tables(self.h, lc.size, lp.ctypes.data, lc.ctypes.data, ix.ctypes.data, eg.ctypes.data,
                                                ep.ctypes.data, C.byref(h)))
        return Model(self, h.value)

    def model_create(self, G: int) -> Model:
        h = C.c_void_p()
        self._chk(self.lib.pg_model_create(self.h, G, C.byref(h)))
        return Model(self, h.value)

    def model_load(self, path: str):
        h, blob, blen = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._chk(self.lib.pg_model_load(self.h, str(path).encode(), C.byref(h), C.byref(blob), C.byref(blen)))
        b = C.string_at(blob.value, blen.value) if blob.value else b""
        if blob.value:
            self.lib.pg_free(blob)
        return Model(self, h.value), b

    def pack(self, data, off, device: bool = False, total_bytes: int | None = None) -> Reads:
        sb = _SeqBatch(_ptr(data), _ptr(off), len(off) - 1)
        h = C.c_void_p()
        if device:
            self._chk(self.lib.pg_reads_pack_dev(self.h, C.byref(sb), int(total_bytes), C.byref(h)))
        else:
            self._chk(self.lib.pg_reads_pack(self.h, C.byref(sb), C.byref(h)))
        return Reads(self, h.value)

    def classify(self, model: Model, data: np.ndarray, off: np.ndarray, mode: int = 0, min_boot_words: int = 0,
                 want_boot: bool = False, out: np.ndarray | None = None, cert_plan: int = 0, light_max: int = 0,
                 bound_level: int = 0):
        n = len(off) - 1
        sb = _SeqBatch(_ptr(data), _ptr(off), n)
        opts = _Opts(min_boot_words, mode, cert_plan, light_max, bound_level)
        res = out if out is not None else np.zeros(n, RESULT_DTYPE)
        boot = np.zeros((n, PG_NUM_BOOT), np.int32) if want_boot else None
        self._chk(self.lib.pg_classify(self.h, model.h, C.byref(sb), C.byref(opts), _ptr(res), _ptr(boot)))
        return (res, boot) if want_boot else res

    def classify_packed(self, model: Model, reads: Reads, results_dev, boot_dev=None, mode: int = 0,
                        min_boot_words: int = 0, cert_plan: int = 0, light_max: int = 0, bound_level: int = 0) -> None:
        opts = _Opts(min_boot_words, mode, cert_plan, light_max, bound_level)
        self._chk(self.lib.pg_classify_packed(self.h, model.h, reads.h, C.byref(opts), _ptr(results_dev), _ptr(boot_dev)))

    def extract_words(self, model: Model, data: np.ndarray, off: np.ndarray):
        n = len(off) - 1
        sb = _SeqBatch(_ptr(data), _ptr(off), n)
        words = np.zeros(int(off[-1]), np.uint16)
        nw = np.zeros(n, np.int32)
        rev = np.zeros(n, np.uint8)
        self._chk(self.lib.pg_extract_words(self.h, model.h, C.byref(sb), words.ctypes.data, nw.ctypes.data, rev.ctypes.data))
        return words, nw, rev

    def boot_indices(self, n: int, min_boot_words: int = 0) -> np.ndarray:
        k = max(n // 8, min_boot_words)
        out = np.zeros((PG_NUM_BOOT, k), np.uint16)
        self._chk(self.lib.pg_boot_indices(self.h, n, min_boot_words, out.ctypes.data))
        return out


    # ---- FASTA ingest
    def fasta_ingest(self, text: bytes, want_reads: bool = True):
        """-> (ids list[bytes], headers list[bytes], sequence bytes uint8, offsets int64, Reads or None)"""
        cap = max(16, text.count(b">") + 1)
        hoff, idl, hl = np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        n, h = C.c_int64(), C.c_void_p()
        self._chk(self.lib.pg_fasta_ingest(self.h, text, len(text), cap, C.byref(n), hoff.ctypes.data, idl.ctypes.data,
                                           hl.ctypes.data, C.byref(h) if want_reads else None))
        n = n.value
        off = np.zeros(n + 1, np.int64)
        buf = np.zeros(max(int(len(text)), 1), np.uint8)
        self._chk(self.lib.pg_fasta_last_bytes(self.h, n, buf.ctypes.data, buf.size, off.ctypes.data))
        ids = [text[hoff[i]:hoff[i] + idl[i]] for i in range(n)]
        hdr = [text[hoff[i]:hoff[i] + hl[i]] for i in range(n)]
        reads = Reads(self, h.value) if (want_reads and h.value) else None
        return ids, hdr, buf[: int(off[n])], off, reads

    def classify_packed_host(self, model: Model, reads: Reads, mode: int = 0, min_boot_words: int = 0, want_boot: bool = False):
        n = len(reads)
        opts = _Opts(min_boot_words, mode, 0, 0)
        res = np.zeros(n, RESULT_DTYPE)
        boot = np.zeros((n, PG_NUM_BOOT), np.int32) if want_boot else None
        self._chk(self.lib.pg_classify_packed_host(self.h, model.h, reads.h, C.byref(opts), _ptr(res), _ptr(boot)))
        return (res, boot) if want_boot else res

    # ---- first hit per read
    def first_hits(self, text: bytes):
        """-> (kept text bytes, kept 0-based line numbers)"""
        out = np.zeros(max(len(text), 1), np.uint8)
        lines = np.zeros(text.count(b"\n") + 2, np.int64)
        n, k = C.c_int64(), C.c_int64()
        self._chk(self.lib.pg_first_hits(self.h, text, len(text), out.ctypes.data, out.size, C.byref(n), lines.ctypes.data,
                                         lines.size, C.byref(k)))
        return out[: n.value].tobytes(), lines[: k.value].tolist()

    # ---- Megaclust
    def megaclust_raw(self, text: bytes, sim: float = 95.0, ev: float = 1e-20, bits: float = 200.0, every: bool = False,
                      bufs=None):
        """the C ABI call itself -> (n, off, len, count, examined, beyond); bufs = (off, len, count) arrays to reuse"""
        opts = _MegaclustOpts(sim, ev, bits, 1 if every else 0)
        if bufs is None:
            cap = max(16, text.count(b"\n") + 2)
            bufs = (np.zeros(cap, np.int64), np.zeros(cap, np.int32), np.zeros(cap, np.int64))
        off, ln, cnt = bufs
        n, ex, by = C.c_int64(), C.c_int64(), C.c_int64()
        self._chk(self.lib.pg_megaclust(self.h, text, len(text), C.byref(opts), len(off), C.byref(n), off.ctypes.data,
                                        ln.ctypes.data, cnt.ctypes.data, C.byref(ex), C.byref(by)))
        return n.value, off, ln, cnt, ex.value, by.value

    def megaclust(self, text: bytes, sim: float = 95.0, ev: float = 1e-20, bits: float = 200.0, every: bool = False):
        """-> (list[(subject bytes, count)] in order of first appearance, lines examined, lines beyond)"""
        n, off, ln, cnt, ex, by = self.megaclust_raw(text, sim, ev, bits, every)
        return [(text[off[i]:off[i] + ln[i]], int(cnt[i])) for i in range(n)], ex, by

    # ---- Trim join
    def trim_join(self, a: bytes, b: bytes | None = None, paired: bool = False, gap: int = 189, truncate: int = 11,
                  want_reads: bool = False):
        opts = _TrimOpts(gap, truncate)
        n = C.c_int64()
        rh = C.c_void_p()
        cap = 2 * (len(a) + (len(b) if b else 0)) + 4096
        while True:
            buf = C.create_string_buffer(cap)
            rc = self.lib.pg_trim_join(self.h, a, len(a), b, len(b) if b else 0, 1 if paired else 0, C.byref(opts), buf, cap,
                                       C.byref(n), C.byref(rh) if want_reads else None)
            if rc == -6 and n.value > cap:
                cap = n.value + 64
                continue
            self._chk(rc)
            break
        text = buf.raw[: n.value]
        return (text, Reads(self, rh.value)) if want_reads else text

    # ---- Stage B
    def tax_load(self, directory) -> "Tax":
        h = C.c_void_p()
        self._chk(self.lib.pg_tax_load(self.h, str(directory).encode(), C.byref(h)))
        return Tax(self, h.value)

    # ---- Stage C
    def consensus(self, hit_off, lineages, pidents, rdps, first_is_fresh=True):
        """lineages/pidents: list[bytes] per hit; rdps: list[bytes] per read (text after the 5 TABs)."""
        return self.consensus_packed(hit_off, pack_sequences(lineages), pack_sequences(pidents), pack_sequences(rdps),
                                     first_is_fresh)

    def consensus_packed(self, hit_off, lineages, pidents, rdps, first_is_fresh=True):
        """the C ABI call itself: every argument already a (bytes uint8, offsets int64) pair of host arrays."""
        (lb, lo), (pb, po), (rb, ro) = lineages, pidents, rdps
        hit_off = np.ascontiguousarray(hit_off, np.int64)
        n = len(hit_off) - 1
        inp = _ConsensusIn(n, hit_off.ctypes.data, lb.ctypes.data, lo.ctypes.data, pb.ctypes.data, po.ctypes.data,
                           rb.ctypes.data, ro.ctypes.data, 1 if first_is_fresh else 0, 0)
        win = np.zeros(n, np.int64)
        nm = np.zeros(n, np.int32)
        self._chk(self.lib.pg_consensus(self.h, C.byref(inp), win.ctypes.data, nm.ctypes.data))
        return win, nm


def tax_build(directory) -> None:
    """== tax_class -c: pure file conversion, needs no GPU."""
    lib = load_library()
    rc = lib.pg_tax_build(str(directory).encode())
    if rc != 0:
        raise PangeaError(rc, lib.pg_last_error(None).decode())


class Tax:
    def __init__(self, ctx: Context, handle: int):
        self.ctx, self.h = ctx, handle

    def leaf(self, gi) -> np.ndarray:
        gi = np.ascontiguousarray(gi, np.int32)
        out = np.zeros(len(gi), np.int32)
        self.ctx._chk(self.ctx.lib.pg_tax_leaf(self.ctx.h, self.h, gi.ctypes.data, len(gi), out.ctypes.data))
        return out

    def lineage_raw(self, gi, buf=None, off=None):
        """the C ABI call itself: lineage strings of all hits back to back in `buf`, hit i at off[i]..off[i+1]."""
        gi = np.ascontiguousarray(gi, np.int32)
        n = len(gi)
        if off is None:
            off = np.zeros(n + 1, np.int64)
        if buf is None:
            buf = np.zeros(max(4096, 160 * n), np.uint8)
        while True:
            rc = self.ctx.lib.pg_tax_lineage(self.ctx.h, self.h, gi.ctypes.data, n, buf.ctypes.data, buf.size, off.ctypes.data)
            if rc == -6 and off[n] > buf.size:
                buf = np.zeros(int(off[n]) + 64, np.uint8)
                continue
            self.ctx._chk(rc)
            break
        return buf, off

    def lineage(self, gi) -> list[bytes]:
        buf, off = self.lineage_raw(gi)
        raw = buf.tobytes()
        return [raw[off[i]:off[i + 1]] for i in range(len(off) - 1)]

    def free(self) -> None:
        if self.h:
            self.ctx.lib.pg_tax_free(self.h)
            self.h = 0
