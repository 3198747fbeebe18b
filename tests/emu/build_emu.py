"""TEST INFRASTRUCTURE ONLY.  Builds tests/emu/_build/libmcb200_emu.so: the product's CUDA sources (csrc/mcb_api.cu with
mcb_kernels.cuh, mcb_lower.cpp, mcb_jit.cpp) compiled by g++ against tests/emu/cuda_runtime.h, which executes the
kernels' source on the host (fibers per CUDA thread).  Used by tests/test_emu_*.py in the CPU tier to run the device
code at small grid sizes against the oracle.  The product never loads it: the package only ever opens libmcb200.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "marching-cube-for-implicit-surfaces_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libmcb200_emu.so")
SOURCES = ["mcb_api.cu", "mcb_lower.cpp", "mcb_jit.cpp"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_runtime.h"), __file__,
                                                                   os.path.join(ROOT, "include", "mcb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = os.path.join(CSRC, "mcb_pow_src.inc")
    if not os.path.exists(inc):  # normally written by the product build
        with open(os.path.join(CSRC, "mcb_pow.h")) as f, open(inc, "w") as o:
            o.write('R"MCBPOWSRC(' + f.read() + ')MCBPOWSRC"\n')
    stamp = os.path.join(CSRC, "mcb_build_stamp.inc")
    if not os.path.exists(stamp):  # normally written by the product build
        with open(stamp, "w") as o:
            o.write('"emulated build"\n')
    objs = []
    for s in SOURCES:
        o = os.path.join(OUT_DIR, s + ".o")
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-w", "-I", HERE, "-I", CSRC, "-x", "c++", "-c",
               os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("emulated build failed: " + s)
        objs.append(o)
    r = subprocess.run(["g++", "-shared", "-o", LIB] + objs + ["-ldl", "-pthread"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("emulated link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
