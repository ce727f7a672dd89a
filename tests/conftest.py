import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Tests never build implicitly on the GPU box; here (CPU) make sure the .so exists."""
    import __graft_entry__ as g

    if not (ROOT / "catalax_b200" / "libctxb200.so").exists():
        g.build()
    yield
