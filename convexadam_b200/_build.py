"""In-tree nvcc build of libconvexadam_b200.so (sm_100a only; no JIT cache, the .so travels with the repo)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libconvexadam_b200.so")
SOURCES = ["common.cu", "pool.cu", "mind.cu", "correlate.cu", "correlate_fused.cu", "correlate_pipe.cu", "coupled.cu", "field.cu", "adam.cu", "mask.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",            # every order-sensitive op is an explicit _rn intrinsic; the rest may fuse
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA toolkit is required to build convexadam_b200")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "convexadam_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    env = dict(os.environ)
    env.pop("CC", None)     # the image exports CC=/opt/gcc/bin/gcc whose spec files are incomplete; use the system gcc
    env.pop("CXX", None)
    res = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
