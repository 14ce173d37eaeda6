"""ctypes wrapper of oracle/libpipeline_ref.so (Stage B / C CPU checkers) and helpers that
drive the REAL reference (Perl + compiled ncbitc.c) when it is available.  Tests only."""
import ctypes as C
import shutil
import subprocess
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
ORACLE_DIR = REPO / "oracle"
REF = Path("/root/reference")
REF_TAX_CLASS = ORACLE_DIR / "_ref" / "tax_class"

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = ORACLE_DIR / "libpipeline_ref.so"
        srcs = [ORACLE_DIR / "taxcollector_ref.c", ORACLE_DIR / "consensus_ref.c", ORACLE_DIR / "trim_ref.c",
                ORACLE_DIR / "megaclust_ref.c", ORACLE_DIR / "uniq_ref.c"]
        if not so.exists() or any(so.stat().st_mtime < s.stat().st_mtime for s in srcs):
            subprocess.run(["make", "-C", str(ORACLE_DIR), str(so)], check=True, capture_output=True)
        L = C.CDLL(str(so))
        L.txc_file.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.cns_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64]
        L.trim_run.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.mc_ref_megaclust.restype = C.c_longlong
        L.mc_ref_megaclust.argtypes = [C.c_char_p, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_int, C.c_longlong,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.uq_ref_first_hits.restype = C.c_longlong
        L.uq_ref_first_hits.argtypes = [C.c_char_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mc_ref_number.restype = C.c_double
        L.mc_ref_number.argtypes = [C.c_char_p, C.c_int]
        _lib = L
    return _lib


def have_reference() -> bool:
    return (REF / "Tax_class" / "NCBI-taxcollector-0.01.pl").exists() and shutil.which("perl") is not None


def ref_tax_class() -> Path:
    """the reference's tax_class: prebuilt oracle/_ref/tax_class, rebuilt here when the sources are present"""
    if not REF_TAX_CLASS.exists() and (REF / "Tax_class" / "ncbitc.c").exists():
        subprocess.run(["make", "-C", str(ORACLE_DIR), "ref"], check=True, capture_output=True)
    return REF_TAX_CLASS


def ref_build_bins(dump_dir: Path, out_dir: Path):
    """tax_class -c of the reference in a scratch copy of the dumps"""
    out_dir.mkdir(parents=True, exist_ok=True)
    for f in ("nodes.dmp", "names.dmp", "gi_taxid_nucl.dmp"):
        shutil.copy(Path(dump_dir) / f, out_dir / f)
    subprocess.run([str(ref_tax_class()), "-c"], cwd=out_dir, check=True)


def oracle_taxcollector(bin_dir, hits, out) -> int:
    return lib().txc_file(str(bin_dir).encode(), str(hits).encode(), str(out).encode())


def oracle_consensus(blast, rdp, out, grace=16) -> int:
    return lib().cns_run(str(blast).encode(), str(rdp).encode(), str(out).encode(), grace)


def real_taxcollector(dump_dir, hits, out):
    with tempfile.TemporaryDirectory() as wd:
        wd = Path(wd)
        ref_build_bins(Path(dump_dir), wd / "Tax_class")
        shutil.copy(ref_tax_class(), wd / "Tax_class" / "tax_class")
        subprocess.run(["perl", str(REF / "Tax_class" / "NCBI-taxcollector-0.01.pl"), "-f", str(hits), "-o", str(out)],
                       cwd=wd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def real_consensus(blast, rdp, out, timeout=120):
    subprocess.run(["perl", str(REF / "Consensus" / "Consensus_BLAST_SOAP_RDP-1.1.pl"), "-b", str(blast), "-r", str(rdp),
                    "-o", str(out)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)


def group_lineages(class_file):
    ids, by = [], []
    for l in Path(class_file).read_text().split("\n"):
        if not l:
            continue
        f = l.split("\t")
        if not ids or ids[-1] != f[0]:
            ids.append(f[0])
            by.append([])
        by[-1].append(f[1] if len(f) > 1 else "")
    return ids, by


def oracle_trim(a, b, gap, truncate, out) -> int:
    return lib().trim_run(str(a).encode(), str(b).encode() if b else None, gap, truncate, str(out).encode())


def real_trim(a, b, gap, out, truncate=None):
    """runs the reference's Trim/trim2.4.pl in a scratch cwd and copies its _runblast.fasta to `out`"""
    a, out = Path(a).resolve(), Path(out).resolve()
    with tempfile.TemporaryDirectory() as wd:
        cmd = ["perl", str(REF / "Trim" / "trim2.4.pl"), "-a", str(a)]
        if b:
            cmd += ["-b", str(Path(b).resolve())]
        if gap is not None:
            cmd += ["-g", str(gap)]
        if truncate is not None:
            cmd += ["-t", str(truncate)]
        r = subprocess.run(cmd, cwd=wd, capture_output=True, text=True, timeout=300)
        shutil.copy(Path(wd) / "output_files" / "trim2" / (a.name + "_runblast.fasta"), out)
        # the script also litters <dir of a>/singletons/
        sd = a.parent / "singletons"
        if sd.exists():
            shutil.rmtree(sd, ignore_errors=True)
        return r.stdout


# ------------------------------------------------------------------ Megaclust

def oracle_megaclust(text: bytes, sim=95.0, ev=1e-20, bits=200.0, every=False):
    """-> (list[(subject bytes, count)] in order of first appearance, lines examined, lines beyond)"""
    import numpy as np

    cap = text.count(b"\n") + 2
    off = np.zeros(cap, np.int64)
    ln = np.zeros(cap, np.int32)
    cnt = np.zeros(cap, np.int64)
    ex, by = C.c_longlong(), C.c_longlong()
    n = lib().mc_ref_megaclust(text, len(text), sim, ev, bits, 1 if every else 0, cap, off.ctypes.data, ln.ctypes.data,
                               cnt.ctypes.data, C.byref(ex), C.byref(by))
    assert n >= 0
    return [(text[off[i]:off[i] + ln[i]], int(cnt[i])) for i in range(n)], ex.value, by.value


def oracle_number(s: bytes) -> float:
    return lib().mc_ref_number(s, len(s))


def have_megaclust_reference() -> bool:
    return (REF / "Megaclust" / "megaclust2.pl").exists() and shutil.which("perl") is not None


def real_megaclust(text: bytes, args=()):
    """the live script -> (set of output lines after the header, header, stdout)"""
    with tempfile.TemporaryDirectory() as wd:
        wd = Path(wd)
        (wd / "in.txt").write_bytes(text)
        r = subprocess.run(["perl", str(REF / "Megaclust" / "megaclust2.pl"), "-i", "in.txt", "-o", "out.txt", *args],
                           cwd=wd, capture_output=True)
        out = (wd / "out.txt").read_bytes() if (wd / "out.txt").exists() else b""
    lines = out.split(b"\n")
    return sorted(l for l in lines[1:] if l), lines[0], r.stdout


def real_megaclustable(files: dict, level: str, order=None):
    """files: name -> bytes; the live megaclustable.pl -> output bytes"""
    with tempfile.TemporaryDirectory() as wd:
        wd = Path(wd)
        for k, v in files.items():
            (wd / k).write_bytes(v)
        names = order or list(files)
        subprocess.run(["perl", str(REF / "Megaclustable" / "megaclustable.pl"), "-m", *names, "-t", level, "-o", "table.txt"],
                       cwd=wd, capture_output=True)
        return (wd / "table.txt").read_bytes() if (wd / "table.txt").exists() else None


# ------------------------------------------------------------------ first hit per read

def oracle_first_hits(text: bytes):
    """-> (kept text bytes, kept 0-based line numbers)"""
    import numpy as np

    out = C.create_string_buffer(max(len(text), 1))
    lines = np.zeros(text.count(b"\n") + 2, np.int64)
    nk = C.c_longlong()
    n = lib().uq_ref_first_hits(text, len(text), out, lines.ctypes.data, C.byref(nk))
    assert n >= 0
    return out.raw[:n], lines[: nk.value].tolist()


def real_get_uniq(text: bytes):
    """the live Scripts/get_uniq.pl -> bytes of <file>.unique"""
    with tempfile.TemporaryDirectory() as wd:
        wd = Path(wd)
        (wd / "hits.txt").write_bytes(text)
        subprocess.run(["perl", str(REF / "Scripts" / "get_uniq.pl"), "-f", "hits.txt"], cwd=wd, capture_output=True)
        return (wd / "hits.txt.unique").read_bytes()
