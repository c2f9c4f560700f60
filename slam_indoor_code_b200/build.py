"""Builds libslamb200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m slam_indoor_code_b200.build [--force]

The shared object lands in slam_indoor_code_b200/lib/ (git-ignored, but it travels to the GPU box
with the repo snapshot).  cudart is linked statically and the driver API is reached through
cudaGetDriverEntryPoint, so the library loads on a machine without libcuda (the symbol-export
test runs without a GPU).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libslamb200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-ccbin", "/usr/bin/g++", "-DSLAMB200_BUILD"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "slamb200.h"))
    if src.endswith(".cpp") and _stale(obj, [os.path.join(CSRC, src)] + headers):
        # host-only units (SIMD intrinsics): the host compiler directly
        cmd = ["/usr/bin/g++", "-O3", "-std=c++17", "-fPIC", "-Wall", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed on {src}:\n{r.stdout}\n{r.stderr}")
    elif _stale(obj, [os.path.join(CSRC, src)] + headers):
        cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
    return obj


def build(force=False, verbose=False, defines=()):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    global CFLAGS
    if defines:
        CFLAGS = CFLAGS + [f"-D{d}" for d in defines]
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ccbin",
                                                                "/usr/bin/g++", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_host_shim()
    return LIB


def build_host_shim():
    """The C++ drop-in TU (host/featureMatchingB200.cpp) compiled against host/cv_shim.h plus C
    entry points for the tests.  With the real OpenCV the TU is compiled inside the reference
    tree instead."""
    host = os.path.join(HERE, "host")
    out = os.path.join(LIBDIR, "libslamb200_hostshim.so")
    deps = [os.path.join(host, f) for f in os.listdir(host)] + [LIB]
    if _stale(out, deps):
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-DSLAMB200_CV_SHIM",
               "-I" + os.path.join(os.path.dirname(HERE), "include"), "-o", out,
               os.path.join(host, "host_shim_test.cpp"), "-L" + LIBDIR, "-lslamb200",
               "-Wl,-rpath,$ORIGIN", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host shim build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
