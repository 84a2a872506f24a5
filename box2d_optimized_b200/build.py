"""Build recipe for the native libraries (sm_100a only, no other targets).

  libb2cuda.so        CUDA kernels + the C-ABI of include/b2cuda.h        (nvcc)
  libb2cuda_dist.so   NCCL transport of the slab halo exchange (b2g_dist_*)   (nvcc, links NCCL)
  libb2gpu_scenes.so  drop-in C++ API (include/box2d) + scene shim over it  (g++)

Both are built IN-TREE next to this file so they travel with a gpurun snapshot.
nvcc cross-compiles without a GPU, so build() works on the CPU-only container.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-parity with the reference's x86-64 build: no FMA contraction, IEEE div/sqrt
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("[build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _find(tool, fallback):
    return shutil.which(tool) or fallback


def build_cuda(force=False, out_name="libb2cuda.so", extra_flags=()):
    """extra_flags / out_name: debug variants (e.g. -DB2G_BIG_TRACE into libb2cuda_trace.so, loaded by
    scripts/gpu_big_trace.py through B2G_CUDA_LIB)"""
    out = os.path.join(HERE, out_name)
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "b2cuda.h")]
    if force or _newer(out, srcs):
        nvcc = _find("nvcc", "/usr/local/cuda/bin/nvcc")
        _run([nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", out, os.path.join(CSRC, "b2g_capi.cu")])
    return out


def _nccl_dirs():
    """NCCL as bundled with torch (the process then holds ONE NCCL, torch's), else the system one"""
    import sysconfig
    sp = sysconfig.get_paths()["purelib"]
    inc, lib = os.path.join(sp, "nvidia", "nccl", "include"), os.path.join(sp, "nvidia", "nccl", "lib")
    if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
        return inc, lib
    return "/usr/include", "/usr/lib/x86_64-linux-gnu"


def build_dist(force=False):
    """libb2cuda_dist.so: the NCCL transport of the halo exchange (b2g_dist_*), over libb2cuda.so"""
    out = os.path.join(HERE, "libb2cuda_dist.so")
    src = os.path.join(CSRC, "b2g_dist.cu")
    if force or _newer(out, [src, os.path.join(ROOT, "include", "b2cuda.h"), os.path.join(HERE, "libb2cuda.so")]):
        nvcc = _find("nvcc", "/usr/local/cuda/bin/nvcc")
        inc, lib = _nccl_dirs()
        _run([nvcc] + NVCC_FLAGS + ["-I" + inc, "-o", out, src, "-L" + HERE, "-lb2cuda", "-L" + lib, "-l:libnccl.so.2",
                                    "-Xlinker", "-rpath,$ORIGIN", "-Xlinker", "-rpath," + lib])
    return out


def build_host(force=False):
    out = os.path.join(HERE, "libb2gpu_scenes.so")
    srcs = [os.path.join(HOST, f) for f in ("b2_world_host.cpp", "b2_shapes.cpp", "gpu_scene_shim.cpp")]
    deps = srcs + [os.path.join(ROOT, "scenes", f) for f in os.listdir(os.path.join(ROOT, "scenes"))]
    deps += [os.path.join(ROOT, "include", "box2d", f) for f in os.listdir(os.path.join(ROOT, "include", "box2d"))]
    deps += [os.path.join(HERE, "libb2cuda.so")]
    if force or _newer(out, deps):
        gxx = _find("g++", "g++")
        _run([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wl,-Bsymbolic",
              "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "scenes")] + srcs +
             ["-L" + HERE, "-lb2cuda", "-Wl,-rpath,$ORIGIN", "-o", out])
    return out


def build_oracle(force=False):
    """Builds the checker (test infrastructure): the C restatement always, the compiled
    reference only where /root/reference exists (this container; the GPU box uses the prebuilt
    oracle/_ref/libb2ref.so that travelled with the snapshot)."""
    odir = os.path.join(ROOT, "oracle")
    built = []
    if os.path.exists(os.path.join(odir, "b2_oracle.c")):
        _run(["make", "-C", odir, "port"])
        built.append(os.path.join(odir, "libb2oracle.so"))
    if os.path.isdir("/root/reference/src"):
        _run(["make", "-C", odir, "ref", "-j8"])
        built.append(os.path.join(odir, "_ref", "libb2ref.so"))
    return built


def build_reference_benchmarks(force=False):
    """The reference's own benchmark programs, compiled UNCHANGED from where they lie under /root/reference
    (testbed/benchmarks/single.cpp and benchmarks.h: scenes b1..b14) against (a) this repo's drop-in headers +
    CUDA library and (b) the reference's headers + its CPU objects (oracle/_ref/obj, the checker).  Plus
    tests/cpp/bench_suite.cpp, which drives the same unchanged benchmarks.h and prints end states.  Only where
    /root/reference exists; the binaries (tests/cpp/build/, git-ignored) travel to the GPU box."""
    import glob
    ref = "/root/reference"
    bdir = os.path.join(ref, "testbed", "benchmarks")
    if not os.path.isdir(bdir):
        return []
    out = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    gxx = _find("g++", "g++")
    objs = [o for o in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "obj", "*", "*.o")))
            if not o.endswith("ref_harness.o")]
    suite = os.path.join(ROOT, "tests", "cpp", "bench_suite.cpp")
    gpu_deps = [os.path.join(HERE, "libb2gpu_scenes.so"), os.path.join(HERE, "libb2cuda.so")]
    built = []
    for name, src in (("single", os.path.join(bdir, "single.cpp")), ("bench_suite", suite)):
        deps = [src, os.path.join(bdir, "benchmarks.h")]
        exe = os.path.join(out, name + "_gpu")
        if force or _newer(exe, deps + gpu_deps):
            _run([gxx, "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + bdir, src, "-L" + HERE,
                  "-lb2gpu_scenes", "-lb2cuda", "-Wl,-rpath,$ORIGIN/../../../box2d_optimized_b200", "-o", exe])
        built.append(exe)
        exe = os.path.join(out, name + "_ref")
        if objs and (force or _newer(exe, deps + objs[:1])):
            _run([gxx, "-O2", "-std=c++11", "-I" + os.path.join(ref, "include"), "-I" + bdir, src] + objs + ["-o", exe])
        if objs:
            built.append(exe)
    # config 4 through b2WorldBatch: W instances of the reference's unchanged b3 in one batch (drop-in build only)
    src = os.path.join(ROOT, "tests", "cpp", "batch_tumbler.cpp")
    exe = os.path.join(out, "batch_tumbler_gpu")
    if force or _newer(exe, [src, os.path.join(bdir, "benchmarks.h")] + gpu_deps):
        _run([gxx, "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + bdir, src, "-L" + HERE,
              "-lb2gpu_scenes", "-lb2cuda", "-Wl,-rpath,$ORIGIN/../../../box2d_optimized_b200", "-o", exe])
    built.append(exe)
    return built


def build_all(force=False):
    return ([build_cuda(force), build_dist(force), build_host(force)] + build_oracle(force) +
            build_reference_benchmarks(force))


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
