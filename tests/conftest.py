import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MESHES = os.path.join(GOLDEN, "meshes")
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "domain-decomposed-pde-solver_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden_values.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


def mesh_path(name):
    return os.path.join(MESHES, name + ".exo")
