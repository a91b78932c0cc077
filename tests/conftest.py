import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libggml_ref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {p.stem: np.load(p) for p in GOLDEN.glob("*.npz")}
