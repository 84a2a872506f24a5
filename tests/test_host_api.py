"""The public C++ API acceptance program (tests/cpp/api_tests.cpp) — restated reference unit
tests plus error-convention checks — compiled against the reference (CPU, proves the program)
and against this repo's drop-in headers + CUDA library (GPU, proves the drop-in)."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "api_tests.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "build")


def _run(cmd, **kw):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)


def test_api_program_passes_on_the_reference():
    """CPU: needs /root/reference (this container).  Skipped on the GPU box, which has no sources."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("/root/reference is not present on this machine")
    os.makedirs(OUT, exist_ok=True)
    objs = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj", "*", "*.o")))
    objs = [o for o in objs if not o.endswith("ref_harness.o")]
    if not objs:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"])
        objs = [o for o in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj", "*", "*.o")))]
    exe = os.path.join(OUT, "api_ref")
    r = _run(["g++", "-O2", "-std=c++11", "-I/root/reference/include", SRC] + objs + ["-o", exe])
    assert r.returncode == 0, r.stdout
    r = _run([exe])
    assert r.returncode == 0, r.stdout
    assert "all API checks passed" in r.stdout


def test_api_program_compiles_against_the_drop_in_headers():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(ROOT, "box2d_optimized_b200")
    exe = os.path.join(OUT, "api_gpu")
    r = _run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + lib, "-lb2gpu_scenes",
              "-lb2cuda", "-Wl,-rpath," + lib, "-o", exe])
    assert r.returncode == 0, r.stdout


@pytest.mark.gpu
def test_api_program_passes_on_the_gpu():
    test_api_program_compiles_against_the_drop_in_headers()
    r = _run([os.path.join(OUT, "api_gpu")])
    assert r.returncode == 0, r.stdout
    assert "all API checks passed" in r.stdout
