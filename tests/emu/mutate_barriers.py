"""tests/emu/mutate_barriers.py -- TEST INFRASTRUCTURE (not part of the pytest suite; run by hand).

How much do the emulator's thread orders see?  Every `__syncthreads()` of one kernel file is removed in
turn (in a scratch copy of csrc/), the emulated library is rebuilt, and a few parity checks run under each
IPB_EMU_ORDER.  Output: one line per barrier with the orders under which a check failed (or the emulator
reported a deadlock).  A barrier no order notices is either redundant or guards a hazard that needs true
concurrency to show (the emulator runs one thread at a time) -- it stays in the kernel either way.

    python tests/emu/mutate_barriers.py [--lines=119,121] ipb_fa_smem.cuh check_fa_overflow check_fa_wide_crop
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "imageprocess_b200", "csrc")
ORDERS = ["forward", "reverse", "random:5:preempt"]

RUNNER = r"""
import sys
sys.path.insert(0, {root!r})
from tests.emu import build_emu
build_emu.build = lambda force=False: {so!r}
from imageprocess_b200.ops import Engine
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests import checks
eng = Engine(emu_lib(), NumpyMem())
for name in {names!r}:
    if name.startswith("fa_path"):
        checks.check_fa_batch(eng, checks.FA_CASES[0], fa_path=int(name[-1]))
    else:
        getattr(checks, name)(eng)
"""


def build(csrc, out):
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-DIPB_EMULATE", "-ffp-contract=off",
           "-fno-strict-aliasing", "-Wno-unused-function", "-I", HERE, "-I", csrc, "-x", "c++",
           os.path.join(csrc, "ipb_api.cu"), "-x", "c++", os.path.join(HERE, "cuda_emu.cpp"), "-o", out]
    subprocess.check_call(cmd)


def run(so, names, order):
    env = dict(os.environ, IPB_EMU_ORDER=order)
    try:
        r = subprocess.run([sys.executable, "-c", RUNNER.format(root=ROOT, so=so, names=names)], env=env,
                           capture_output=True, text=True, timeout=900)
    except subprocess.TimeoutExpired:
        return "timeout"
    if r.returncode == 0:
        return None
    if r.returncode < 0:
        return f"signal {-r.returncode}"                     # e.g. an index built from stale shared memory
    tail = (r.stderr.strip().splitlines() or ["?"])[-1]
    return "deadlock" if "deadlock" in r.stderr else ("assert" if "Assert" in r.stderr else tail[:60])


def main():
    argv = sys.argv[1:]
    only = None
    if argv and argv[0].startswith("--lines="):            # restrict to these source lines (e.g. the ones a first run did not see)
        only = {int(v) for v in argv.pop(0)[8:].split(",")}
    fname, names = argv[0], argv[1:]
    tmp = tempfile.mkdtemp(prefix="ipb_mut_")
    try:
        csrc = os.path.join(tmp, "csrc")
        shutil.copytree(CSRC, csrc)
        src = open(os.path.join(CSRC, fname)).read().split("\n")
        lines = [i for i, ln in enumerate(src) if re.search(r"__syncthreads\(\);", ln) and (only is None or i + 1 in only)]
        so = os.path.join(tmp, "emu.so")
        build(csrc, so)
        base = [run(so, names, o) for o in ORDERS]
        print(f"{fname}: {len(lines)} barriers; checks: {' '.join(names)}; unmodified: {base}", flush=True)
        seen = {o: 0 for o in ORDERS}
        any_seen = 0
        for i in lines:
            mut = list(src)
            mut[i] = mut[i].replace("__syncthreads();", "/* removed */;", 1)
            open(os.path.join(csrc, fname), "w").write("\n".join(mut))
            build(csrc, so)
            res = [run(so, names, o) for o in ORDERS]
            for o, r in zip(ORDERS, res):
                seen[o] += r is not None
            any_seen += any(r is not None for r in res)
            print(f"  line {i + 1:4d}: " + "  ".join(f"{o}={r or 'ok'}" for o, r in zip(ORDERS, res)) +
                  f"    | {src[i].strip()[:70]}", flush=True)
        open(os.path.join(csrc, fname), "w").write("\n".join(src))
        print(f"seen by: {seen}; by at least one order: {any_seen} of {len(lines)}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
