import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("ADPRES_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_problem(name):
    """A reference sample deck, from the committed fixture (works on the GPU box too)."""
    from adpres_b200.deck import Problem
    with open(os.path.join(GOLDEN, name + ".spec.json")) as fh:
        return Problem.from_spec(json.load(fh))


@pytest.fixture(scope="session")
def iaea3ds():
    return load_problem("IAEA3Ds")


@pytest.fixture(scope="session")
def golden_trace():
    with open(os.path.join(GOLDEN, "iaea3ds_trace.json")) as fh:
        return json.load(fh)
