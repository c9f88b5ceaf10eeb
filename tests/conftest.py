import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import __graft_entry__ as entry  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (ctypes over lib/libaprilgrid_b200.so)."""
    p = entry.load_package()
    if not os.path.exists(p.LIB_PATH):
        entry.build()
    return p


@pytest.fixture(scope="session")
def oracle():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def expected():
    with open(os.path.join(GOLDEN, "expected.json")) as f:
        return json.load(f)


def load_image(name):
    from PIL import Image
    return np.ascontiguousarray(np.array(Image.open(os.path.join(GOLDEN, "images", name + ".png"))))


@pytest.fixture(scope="session")
def images():
    class _Images(dict):
        def __missing__(self, k):
            self[k] = load_image(k)
            return self[k]
    return _Images()


@pytest.fixture(scope="session")
def detector(pkg):
    """One T36H11 detector on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    det = pkg.TagDetector(pkg.TagFamily.T36H11, None, device=0)
    yield det
    det.close()


FIXTURE_NAMES = ["iphone", "EuRoC", "TUM_VI", "right", "r45", "top", "two_boards", "top_right",
                 "demo_1520525725372653511"]

# tests/tools/ holds stand-alone checking scripts (they use the oracle, so they live with the tests)
collect_ignore_glob = ["tools/*"]
