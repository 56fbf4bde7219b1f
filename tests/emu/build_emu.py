"""tests/emu/build_emu.py -- TEST INFRASTRUCTURE.  Builds tests/emu/_build/libipb200_emu.so:
the product's kernel sources compiled with g++ -DIPB_EMULATE against cuda_emu.h."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "imageprocess_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libipb200_emu.so")


def newest_src():
    t = 0.0
    for d in (CSRC, HERE):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest_src():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-DIPB_EMULATE",
           "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-unused-function",
           "-I", HERE, "-I", CSRC, "-x", "c++", os.path.join(CSRC, "ipb_api.cu"),
           "-x", "c++", os.path.join(HERE, "cuda_emu.cpp"), "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
