"""The drop-in boundary is plain C: both headers must compile as pedantic C99 (the reference's language), and a C program
written against include/pdt.h must link against the shared object and fail cleanly without a device.  No device work."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "project-desert-tortoise_b200")
CFLAGS = ["-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include")]


@pytest.mark.parametrize("decimal", ["float", "double"])
def test_headers_are_pedantic_c99(tmp_path, decimal):
    src = tmp_path / "tu.c"
    src.write_text(f'#define DECIMAL_TYPE {decimal}\n#include "pdt.h"\n#include "pdt_legacy.h"\n'
                   "int probe(void) { pdt_params p; pdt_stream_plan s; (void)p; (void)s; "
                   "return (int)sizeof(pdt_frame) + (int)sizeof(pdt_capture_stats) + (int)sizeof(pdt_frame_quality); }\n")
    subprocess.run(["gcc", *CFLAGS, "-c", str(src), "-o", str(tmp_path / "tu.o")], check=True)


def test_c_example_links_and_fails_cleanly_without_arguments(tmp_path):
    exe = tmp_path / "demod_pcm"
    subprocess.run(["gcc", *CFLAGS, "-o", str(exe), os.path.join(ROOT, "examples", "demod_pcm.c"), "-L" + PKG, "-lpdt_f32",
                    "-Wl,-rpath," + PKG, "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and "usage:" in r.stderr and "pdt-b200" in r.stderr      # pdt_version() came from the library


def test_c_live_example_links_and_fails_cleanly_without_arguments(tmp_path):
    exe = tmp_path / "live_stdin"
    subprocess.run(["gcc", *CFLAGS, "-o", str(exe), os.path.join(ROOT, "examples", "live_stdin.c"), "-L" + PKG, "-lpdt_f32",
                    "-Wl,-rpath," + PKG, "-lm"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and "usage:" in r.stderr and "pdt-b200" in r.stderr

