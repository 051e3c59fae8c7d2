import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "hf_t5_tiny.npz"), allow_pickle=False))
    g["meta"] = json.load(open(os.path.join(ROOT, "tests", "golden", "hf_t5_tiny.json")))
    return g


@pytest.fixture(scope="session")
def ref_helpers():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "reference_helpers.json"), encoding="utf-8"))


@pytest.fixture(scope="session")
def built_lib():
    """the in-tree libp5b200.so (built by __graft_entry__.build(); rebuilt here if nvcc is around and it is missing)"""
    from openp5_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from openp5_b200.build import build
        build()
    return _lib.load()
