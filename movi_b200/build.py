"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  _lib/libmovi_gpu.so    CUDA kernels + C ABI (include/movi_gpu.h), nvcc, sm_100a only
  _lib/libmovi_host.so   host helpers (include/movi_host.h), g++
  _lib/movi-b200         command line front end mirroring `movi query` (host/movi_cli.cpp)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(HERE, "_lib")
GPU_LIB = os.path.join(LIBDIR, "libmovi_gpu.so")
HOST_LIB = os.path.join(LIBDIR, "libmovi_host.so")
CLI = os.path.join(LIBDIR, "movi-b200")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared", "-Xptxas", "-v"]
HOST_CXX = os.environ.get("MOVI_HOST_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _glob(d, exts):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(exts)]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build_gpu(force=False, verbose=False):
    srcs = _glob(os.path.join(HERE, "csrc"), (".cu", ".cuh")) + [os.path.join(ROOT, "include", "movi_gpu.h")]
    if force or _newer(GPU_LIB, srcs):
        os.makedirs(LIBDIR, exist_ok=True)
        cmd = [nvcc_path()] + NVCC_FLAGS + [os.path.join(HERE, "csrc", "movi_gpu.cu"), "-o", GPU_LIB]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        with open(os.path.join(LIBDIR, "nvcc.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout)
        if r.returncode:
            raise RuntimeError("nvcc failed")
    return GPU_LIB


def build_host(force=False):
    srcs = _glob(os.path.join(HERE, "host"), (".cpp", ".hpp")) + _glob(os.path.join(ROOT, "include"), (".h",))
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _newer(HOST_LIB, srcs):
        subprocess.check_call([HOST_CXX, "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared",
                               os.path.join(HERE, "host", "movi_host.cpp"), "-lz", "-o", HOST_LIB])
    cli_src = os.path.join(HERE, "host", "movi_cli.cpp")
    if os.path.exists(cli_src) and (force or _newer(CLI, srcs + [GPU_LIB])):
        subprocess.check_call([HOST_CXX, "-std=c++17", "-O2", "-Wall", cli_src, os.path.join(HERE, "host", "movi_host.cpp"),
                               "-L" + LIBDIR, "-lmovi_gpu", "-lz", "-Wl,-rpath,$ORIGIN", "-pthread", "-o", CLI])
    return HOST_LIB


SYNTH_LIB = os.path.join(LIBDIR, "libmovi_synth.so")


def build_synth(force=False):
    """Bench data generator: synthetic human-scale tables built directly in HBM (datagen/synth_gpu.cu)."""
    src = os.path.join(HERE, "datagen", "synth_gpu.cu")
    if force or _newer(SYNTH_LIB, [src, GPU_LIB]):
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--compiler-options", "-fPIC",
               "-shared", src, "-L" + LIBDIR, "-lmovi_gpu", "-Xlinker", "-rpath=$ORIGIN", "-o", SYNTH_LIB]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout)
            raise RuntimeError("nvcc failed (synth_gpu.cu)")
    return SYNTH_LIB


def build_all(force=False, verbose=False):
    build_gpu(force, verbose)
    build_host(force)
    build_synth(force)
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "datagen")])


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
