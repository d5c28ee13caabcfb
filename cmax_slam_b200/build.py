"""In-tree build of libcmax_b200.so (hand-written CUDA for sm_100a + the C ABI)."""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libcmax_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

# -fmad=false: the f64 event geometry must not be FMA-contracted (cell indices bit-identical to the
# reference on baseline x86-64); explicit fmaf() is used where OpenCV's filter uses FMA.
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into cmax_slam_b200/libcmax_b200.so.  nvcc cross-compiles
    without a GPU.  Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # prebuilt library travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libcmax_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-o", LIB_PATH] + sources()
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
