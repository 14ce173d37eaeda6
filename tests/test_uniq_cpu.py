"""First hit per read (Scripts/get_uniq.pl): the C restatement against the live Perl script."""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))

import oracle_pipeline as op  # noqa: E402

EDGE = (b"q1\tgi|1|\t99.0\nq1\tgi|2|\t98.0\nq2\tgi|3|\t97\n\nq3\nq3\tx\nq3\n\t\tlead\n\tother\nq2 \ty\nq1\tgi|9|\t50\n"
        b"noeol")                                            # "q3\n" (no TAB) and "q3" + TAB are different keys


def hits_text(seed: int, reads: int) -> bytes:
    rng = np.random.default_rng(seed)
    out = []
    for r in range(reads):
        q = f"read{int(rng.integers(0, max(reads // 2, 1))):06d}"
        for h in range(1 + int(rng.integers(0, 4))):
            out.append(f"{q}\tgi|{int(rng.integers(1, 10**6))}|gb|X|\t{rng.uniform(80, 100):.2f}\t250\t{int(rng.integers(50, 900))}")
    return ("\n".join(out) + "\n").encode()


@pytest.mark.skipif(not op.have_megaclust_reference(), reason="reference tree or perl absent")
def test_oracle_matches_live_script():
    for text in (EDGE, EDGE + b"\n", hits_text(1, 300), hits_text(2, 2000), b"a\n", b"a"):
        got, lines = op.oracle_first_hits(text)
        assert got == op.real_get_uniq(text)
        assert len(lines) == len(set(lines)) and lines == sorted(lines)


def test_oracle_known_answer():
    got, lines = op.oracle_first_hits(EDGE)
    assert lines == [0, 2, 3, 4, 5, 7, 9, 11]
    assert got == b"q1\tgi|1|\t99.0\nq2\tgi|3|\t97\n\nq3\nq3\tx\n\t\tlead\nq2 \ty\nnoeol"
