#!/usr/bin/env python
"""Host-side cost of the folder entry points, measured WITHOUT a GPU: a folder of small TIFF pairs
goes through fret_ratio_builder.run_headless / Fluor_INT.run_headless on the emulated build
(tests/emu); the kernels' (emulated, slow) time is excluded by reporting only the host functions.

    python profiles/host_profile.py [n_frames]
"""
import cProfile
import json
import os
import pstats
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    from imageprocess_b200.host import Fluor_INT, common, fret_ratio_builder
    from imageprocess_b200.ops import Engine
    from oracle.gen_golden import small_scene
    from tests.emu.emu_backend import NumpyMem, emu_lib
    eng = Engine(emu_lib(), NumpyMem())
    H, W = 96, 128
    d, a, polys = small_scene(5, H=H, W=W, n_cells=2, blobs=3)
    polys = (polys * 12)[:24]                                     # 24 ROI rows per frame, as the C4 workload
    root = tempfile.mkdtemp(prefix="ipb_hostprof_", dir=os.path.join(ROOT, "gpurun_out", "hostprof"))
    try:
        roi_dir = os.path.join(root, "roi")
        os.makedirs(roi_dir)
        roi_json = json.dumps({"name": "S01", "image_shape": {"height": H, "width": W},
                               "rois": [np.asarray(P).tolist() for P in polys]})
        for t in range(n):
            common.write_tiff(os.path.join(root, f"S01_t{t:03d}_1.tif"), d)
            common.write_tiff(os.path.join(root, f"S01_t{t:03d}_2.tif"), a)
            with open(os.path.join(roi_dir, f"S01_t{t:02d}.json"), "w") as fh:
                fh.write(roi_json)
        for name, fn in (("fret_ratio_builder", lambda: fret_ratio_builder.run_headless(
                              root, roi_dir, out_root=os.path.join(root, "RES_FRET"), eng=eng, log=lambda s: None,
                              p={"timelapse": True, "out_tif": False, "ratio_mode": "Donor/FRET"}, frames_per_batch=32)),
                         ("Fluor_INT", lambda: Fluor_INT.run_headless(
                              root, roi_dir, out_root=os.path.join(root, "RES_INT"), eng=eng, log=lambda s: None,
                              cfg={"timelapse": True, "channels_to_quant": [1, 2]}))):
            pr = cProfile.Profile()
            t0 = time.perf_counter()
            pr.enable()
            rows = fn()
            pr.disable()
            dt = time.perf_counter() - t0
            st = pstats.Stats(pr)
            kern = sum(v[3] for k, v in st.stats.items() if k[2] in ("call", "__call__") and "ops.py" in k[0] or "emu" in k[0] and k[2] == "call")
            print(f"== {name}: {len(rows)} rows, {n} frames, wall {dt:.2f} s")
            st.sort_stats("cumulative").print_stats(45)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
