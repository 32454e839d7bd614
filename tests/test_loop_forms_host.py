"""The branch-free loop-filter forms of the kernels against the reference-shaped functions, ON THE HOST: the very functions
of csrc/ (`__host__ __device__`) compiled by nvcc into a CPU program (tests/host/loop_forms.cu) and compared bit for bit over
millions of random and edge states and over closed-loop trajectories, float and double builds.  No device."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "project-desert-tortoise_b200", "csrc")


@pytest.mark.parametrize("floats", [1, 0])
def test_select_forms_equal_reference_shaped_functions(tmp_path, floats):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not found: the host harness is built from the CUDA headers")
    exe = tmp_path / f"loop_forms_{floats}"
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-fmad=false", "-Xcompiler", "-ffp-contract=off", f"-DPDT_USE_FLOATS={floats}",
                    "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "host", "loop_forms.cu")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-400:]
    assert int(r.stdout.split()[1]) > 5_000_000


def test_tiled_engine_building_blocks_equal_reference_shaped_functions(tmp_path):
    """pdt_tiled.cuh on the host: the unrolled 26-tap block FIR (L = 1…8, both forms) against the rotating-order reference FIR,
    the streaming track-mode PLL loop against the per-sample loop filter, the AGC tile (common-regime chain, side proof,
    fallback) against the per-sample AGC — tests/host/tiled_forms.cu."""
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not found: the host harness is built from the CUDA headers")
    exe = tmp_path / "tiled_forms"
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-fmad=false", "-Xcompiler", "-ffp-contract=off", "-DPDT_USE_FLOATS=1",
                    "-I" + CSRC, "-o", str(exe), os.path.join(ROOT, "tests", "host", "tiled_forms.cu")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout[-400:]
    assert int(r.stdout.split()[1]) > 2_000_000

