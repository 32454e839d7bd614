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
    # A fresh checkout has no built artefacts (they are git-ignored): build what is missing once, each part on its own, and
    # never abort the session over a missing toolchain - the oracle-only tests need gcc alone, and a test whose library
    # could not be built fails (or skips) by itself with the loader's message.  Nothing that exists is rebuilt: on the GPU
    # box the prebuilt libraries travel with the snapshot.
    import shutil
    import subprocess
    import warnings
    pkg = os.path.join(ROOT, "project-desert-tortoise_b200")
    jobs = []
    if not all(os.path.exists(os.path.join(ROOT, "oracle", f"liboracle_{p}.so")) for p in ("f32", "f64")):
        jobs.append(("gcc", ["make", "-C", os.path.join(ROOT, "oracle"), "all"]))
    if not all(os.path.exists(os.path.join(pkg, f"libpdt_{p}.so")) for p in ("f32", "f64")):
        jobs.append(("nvcc", ["make", "-C", os.path.join(pkg, "csrc"), "all"]))
    for tool, cmd in jobs:
        if shutil.which(tool) is None or shutil.which("make") is None:
            warnings.warn(f"{tool}/make not found: {' '.join(cmd[1:3])} not built; the tests that need it will say so")
            continue
        r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            warnings.warn(f"{' '.join(cmd)} failed:\n{r.stderr[-800:]}")


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
