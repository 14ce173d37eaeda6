import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "pangea-plus_b200"))
sys.path.insert(0, str(REPO / "tests"))
sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One pg_ctx on cuda:0 for the whole GPU session."""
    import pangea_b200 as pg

    c = pg.Context(0)
    yield c
    c.close()
