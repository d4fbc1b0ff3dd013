"""pytest configuration: the `gpu` marker and shared fixture loaders."""
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_fixture(name: str) -> bytes:
    """Committed copy of a reference testdata file (tests/golden/make_golden.py)."""
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".gz"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def urls():
    return load_fixture("urls.10K")


@pytest.fixture(scope="session")
def urls_snappy():
    return load_fixture("urls.10K.snappy")


@pytest.fixture(scope="session")
def baddata3():
    return load_fixture("baddata3.snappy")


@pytest.fixture(scope="session")
def unaligned_pair():
    return load_fixture("unaligned_uint64_test.snappy"), load_fixture("unaligned_uint64_test.bin")
