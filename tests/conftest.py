"""pytest configuration: `-m gpu` tests need a B200 and call through the C ABI (libt4b.so);
`-m "not gpu"` tests run on CPU (oracle vs golden vectors, the host-only entry points of the C ABI - rank
rules, sweep plans, adaptive cutoffs -, the world-size-2 gloo run of the sharded patch driver, C-ABI symbol
export)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def ctx():
    import t4b
    c = t4b.Context(0)
    yield c
    c.close()
