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


def _traces(text):
    out = {}
    for line in text.splitlines():
        if line.startswith("TRACE "):
            head, *bodies = line[6:].split(" | ")
            tag, contacts = head.split(" contacts=")
            out[tag] = (int(contacts), [[float(x) for x in b.split()] for b in bodies])
    return out


@pytest.mark.gpu
def test_api_program_passes_on_the_gpu():
    test_api_program_compiles_against_the_drop_in_headers()
    r = _run([os.path.join(OUT, "api_gpu")])
    assert r.returncode == 0, r.stdout
    assert "all API checks passed" in r.stdout
    # the scripted editing session (DestroyBody, SetTransform, DestroyFixture, SetType, SetEnabled,
    # forces, impulses, ...) against the same program linked with the reference: tests/cpp/build/api_ref
    # is built by the CPU test in the container that holds /root/reference and travels with the tree
    ref_exe = os.path.join(OUT, "api_ref")
    if not os.path.exists(ref_exe):
        pytest.skip("tests/cpp/build/api_ref has not been built (needs /root/reference)")
    ref, gpu = _traces(_run([ref_exe]).stdout), _traces(r.stdout)
    assert set(ref) == set(gpu) and len(ref) >= 10
    worst = 0.0
    for tag in ref:
        (rc, rb), (gc, gb) = ref[tag], gpu[tag]
        assert abs(rc - gc) <= 1, f"{tag}: contact count {gc} vs {rc}"
        assert len(rb) == len(gb)
        for x, y in zip(rb, gb):
            if tag != "resting":   # positions and angles; the awake flag only once everything is asleep or not
                d = max(abs(a - b) for a, b in zip(x[:3], y[:3]))
                worst = max(worst, d)
                assert d < 0.02, f"{tag}: {y} vs {x}"
    print(f"editing session: worst position/angle difference to the reference {worst:.5f}")


BATCH_SRC = os.path.join(ROOT, "tests", "cpp", "batch_tests.cpp")


def test_batch_program_compiles_against_the_drop_in_headers():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(ROOT, "box2d_optimized_b200")
    r = _run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), BATCH_SRC, "-L" + lib, "-lb2gpu_scenes",
              "-lb2cuda", "-Wl,-rpath," + lib, "-o", os.path.join(OUT, "batch_gpu")])
    assert r.returncode == 0, r.stdout


@pytest.mark.gpu
def test_worlds_of_a_b2WorldBatch_equal_the_same_worlds_stepped_alone():
    """b2WorldBatch (the drop-in's route to BASELINE config 4): six different worlds with a motor joint, spawns, an
    impulse and a destroyed body, stepped alone and as one batch — bitwise equal world by world; then slot growth,
    a member's QueryAABB, and a member leaving the batch (tests/cpp/batch_tests.cpp)."""
    test_batch_program_compiles_against_the_drop_in_headers()
    r = _run([os.path.join(OUT, "batch_gpu")])
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "all batch checks passed" in r.stdout
