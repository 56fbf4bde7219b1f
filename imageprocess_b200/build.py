"""Builds imageprocess_b200/lib/libipb200.so with nvcc for sm_100a (cross-compiles without a
GPU).  One translation unit: csrc/ipb_api.cu includes every kernel header."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libipb200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared", "-Xptxas", "-v"]


def source_digest():
    """sha256 over the kernel sources + the C header: lets _lib.load() refuse a stale library."""
    import hashlib
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "ipb200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    return h.hexdigest()


HASH = os.path.join(HERE, "lib", "libipb200.sha256")


def newest_src():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
               if f.endswith((".cu", ".cuh", ".h")))


def build(force=False, verbose=False):
    digest = source_digest()
    if not force and os.path.exists(OUT) and os.path.exists(HASH) and open(HASH).read().strip() == digest:
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
    with open(HASH, "w") as f:
        f.write(digest + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
