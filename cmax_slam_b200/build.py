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
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(_HERE, "build")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def _compile_and_link(out, obj_dir, extra, verbose, force):
    """One nvcc -c per translation unit (in parallel, only the stale ones) + one link."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    env = dict(os.environ)
    env.pop("CXX", None)
    hdr_t = max(os.path.getmtime(p) for p in glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")))
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=8) as ex:
        for rc in ex.map(lambda c: subprocess.run(c, env=env).returncode, jobs):
            if rc != 0:
                raise subprocess.CalledProcessError(rc, "nvcc -c")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, env=env)
    return out


def build_variant(name, defines, verbose=False):
    """A/B build with extra -D flags into scratch/variants/libcmax_b200_<name>.so (select it at run time with
    CMAXB_LIB_PATH=<path>); tuning aid, never the shipped library."""
    out_dir = os.path.join(os.path.dirname(_HERE), "scratch", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libcmax_b200_{name}.so")
    return _compile_and_link(out, os.path.join(out_dir, "obj_" + name), [f"-D{d}" for d in defines], verbose, True)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into cmax_slam_b200/libcmax_b200.so.  nvcc cross-compiles
    without a GPU.  Returns the library path."""
    override = os.environ.get("CMAXB_LIB_PATH")
    if override:
        if not os.path.exists(override):
            raise RuntimeError(f"CMAXB_LIB_PATH={override} does not exist")
        return override
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # prebuilt library travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libcmax_b200.so")
    return _compile_and_link(LIB_PATH, OBJ_DIR, [], verbose, force)


if __name__ == "__main__":
    import sys
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if not a.startswith("-")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
