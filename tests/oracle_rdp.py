"""ctypes wrapper of oracle/librdp_ref.so -- the Stage A CPU checker (tests only)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"
NUM_BOOT = 100

RESULT = np.dtype([("genus", "<i4"), ("n_words", "<i4"), ("score", "<f4"), ("reversed", "<i4"),
                   ("status", "<i4"), ("boot", "<i4", (NUM_BOOT,))])


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR), str(ORACLE_DIR / "librdp_ref.so")], check=True,
                   capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = ORACLE_DIR / "librdp_ref.so"
        src = ORACLE_DIR / "rdp_ref.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            build()
        L = C.CDLL(str(so))
        vp = C.c_void_p
        L.rdp_words.argtypes = [C.c_char_p, C.c_int, vp]
        L.rdp_revcomp_word.argtypes = [C.c_int32]
        L.rdp_jrandom_stream.argtypes = [C.c_uint64, C.c_int32, C.c_int, vp]
        L.rdp_jrandom_ints.argtypes = [C.c_uint64, C.c_int, vp]
        L.rdp_model_new.restype = vp
        L.rdp_model_new.argtypes = [C.c_int]
        L.rdp_model_free.argtypes = [vp]
        L.rdp_train_batch.argtypes = [vp, vp, vp, C.c_int64, vp]
        L.rdp_model_N.restype = C.c_int64
        L.rdp_model_N.argtypes = [vp]
        for f in ("m", "nw", "M", "logPrior", "logLeave", "logP"):
            getattr(L, "rdp_model_" + f).restype = vp
            getattr(L, "rdp_model_" + f).argtypes = [vp]
        L.rdp_classify_batch.argtypes = [vp, vp, vp, C.c_int64, C.c_int, vp, C.c_int]
        L.rdp_votes.argtypes = [vp, vp, C.c_int, vp]
        L.rdp_is_reversed.argtypes = [vp, vp, C.c_int]
        assert L.rdp_sizeof_result() == RESULT.itemsize
        _lib = L
    return _lib


def words(seq: bytes) -> np.ndarray:
    out = np.zeros(max(len(seq), 1), np.int32)
    n = lib().rdp_words(seq, len(seq), out.ctypes.data)
    return out[:n]


def jrandom_stream(seed, n, count):
    out = np.zeros(count, np.int32)
    lib().rdp_jrandom_stream(seed, n, count, out.ctypes.data)
    return out


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Model:
    def __init__(self, data, off, genus, G):
        L = lib()
        self.G = G
        self.h = L.rdp_model_new(G)
        data = np.ascontiguousarray(data, np.uint8)
        off = np.ascontiguousarray(off, np.int64)
        genus = np.ascontiguousarray(genus, np.int32)
        L.rdp_train_batch(self.h, data.ctypes.data, off.ctypes.data, len(off) - 1, genus.ctypes.data)

    @property
    def N(self):
        return lib().rdp_model_N(self.h)

    def counts(self):
        L = lib()
        return (_view(L.rdp_model_m(self.h), (65536, self.G), np.int32), _view(L.rdp_model_nw(self.h), (65536,), np.int32),
                _view(L.rdp_model_M(self.h), (self.G,), np.int32), self.N)

    def tables(self):
        L = lib()
        return (_view(L.rdp_model_logPrior(self.h), (65536,), np.float32), _view(L.rdp_model_logLeave(self.h), (self.G,), np.float32),
                _view(L.rdp_model_logP(self.h), (65536, self.G), np.float32))

    def classify(self, data, off, min_boot_words=0, threads=8):
        data = np.ascontiguousarray(data, np.uint8)
        off = np.ascontiguousarray(off, np.int64)
        n = len(off) - 1
        res = np.zeros(n, RESULT)
        lib().rdp_classify_batch(self.h, data.ctypes.data, off.ctypes.data, n, min_boot_words, res.ctypes.data, threads)
        return res

    def votes(self, res, anc):
        anc = np.ascontiguousarray(anc, np.int32)
        out = np.zeros((len(res), anc.shape[1]), np.int32)
        for i in range(len(res)):
            lib().rdp_votes(res[i:i + 1].ctypes.data, anc.ctypes.data, anc.shape[1], out[i].ctypes.data)
        return out

    def free(self):
        if self.h:
            lib().rdp_model_free(self.h)
            self.h = None
