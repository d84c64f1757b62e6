"""Builds tests/emu/libca_emu.so: the product kernel SOURCES compiled with g++ -DCA_EMU against the CUDA-on-CPU shim.
Test infrastructure only (see cuda_emu.h); the package never loads this library."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "convexadam_b200", "csrc")
LIB = os.path.join(HERE, "libca_emu.so")


def build(force=False):
    from convexadam_b200._build import SOURCES
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "ca_platform.h"), os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_emu.cpp"),
                   os.path.join(ROOT, "include", "convexadam_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    objs = []
    procs = []
    for s in srcs + [os.path.join(HERE, "cuda_emu.cpp")]:
        o = os.path.join(HERE, "_obj_" + os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["/usr/bin/g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
               "-DCA_EMU", "-I", HERE, "-Wno-attributes", "-Wno-unknown-pragmas", "-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("emu build failed: " + " ".join(cmd) + "\n" + out)
    subprocess.check_call(["/usr/bin/g++", "-shared", "-fopenmp", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    import sys
    sys.path.insert(0, ROOT)
    print(build(force=True))
