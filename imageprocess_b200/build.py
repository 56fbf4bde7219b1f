"""Builds imageprocess_b200/lib/libipb200.so with nvcc for sm_100a (cross-compiles without a
GPU).  One translation unit: csrc/ipb_api.cu includes every kernel header."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libipb200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared", "-Xptxas", "-v"]


def newest_src():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
               if f.endswith((".cu", ".cuh", ".h")))


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest_src():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(os.path.dirname(HERE), "include"),
                                 "-o", OUT, os.path.join(CSRC, "ipb_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "lib", "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
