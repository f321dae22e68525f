"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """A fresh checkout has no libbemb200.so (built artefacts are not tracked): build it once, in-tree, with nvcc for sm_100a
    (cross-compiles without a GPU).  An existing library is left alone -- on the GPU box it is the one that travelled."""
    from math_audio_b200 import build as _b

    if not _b.LIB.exists():
        _b.build()
    return _b.LIB
