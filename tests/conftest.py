import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, like the driver's build() step.
    # (Never rebuilds what is there: on the GPU box the prebuilt libraries travel with the snapshot.)
    need = [os.path.join(ROOT, "project-desert-tortoise_b200", f"libpdt_{p}.so") for p in ("f32", "f64")] + \
           [os.path.join(ROOT, "oracle", f"liboracle_{p}.so") for p in ("f32", "f64")]
    if not all(os.path.exists(f) for f in need):
        import importlib
        importlib.import_module("__graft_entry__").build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle32():
    import pyoracle
    return pyoracle.Oracle("f32")


@pytest.fixture(scope="session")
def oracle64():
    import pyoracle
    return pyoracle.Oracle("f64")
