"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/pangea_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import re
import subprocess

import pytest

import pangea_b200 as pg


def declared_symbols():
    syms = set()
    for h in sorted(pg.HEADER.parent.glob("*.h")):          # pangea_b200.h and pangea_b200_synth.h
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        syms |= set(re.findall(r"\b(pg_[a-z_0-9]+)\s*\(", text))
    return sorted(syms)


@pytest.fixture(scope="module")
def lib():
    if not pg.LIB_PATH.exists():
        pg.build()
    return pg.load_library()


def test_header_declares_the_three_stages():
    syms = declared_symbols()
    for s in ("pg_init", "pg_train", "pg_classify", "pg_tax_build", "pg_tax_lineage", "pg_consensus"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_no_torch_or_cxx_types_in_signatures():
    code = re.sub(r"/\*.*?\*/", "", pg.HEADER.read_text(), flags=re.S)
    for banned in ("torch", "at::", "std::", "template", "class "):
        assert banned not in code


def test_result_record_is_64_bytes():
    assert pg.RESULT_DTYPE.itemsize == 64


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(pg.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_fails_loudly_without_a_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is for CPU-only hosts")
    h = lib.pg_init(0)
    assert not h
    msg = lib.pg_last_error(None).decode()
    assert "no CUDA device" in msg and "no CPU path" in msg
    with pytest.raises(pg.PangeaError):
        pg.Context(0)
