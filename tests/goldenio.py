"""Loaders for tests/golden (written by oracle/gen_golden.py)."""
import csv
import json
import lzma
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_u16(path, shape=(1536, 2048)):
    raw = lzma.decompress(open(path, "rb").read())
    n = shape[0] * shape[1]
    hi = np.frombuffer(raw[:n], dtype=np.uint8).astype(np.uint16)
    lo = np.frombuffer(raw[n:], dtype=np.uint8).astype(np.uint16)
    return ((hi << 8) | lo).reshape(shape)


def load_intensity(exp):
    d = os.path.join(GOLD, "intensity", exp)
    rois = json.load(open(os.path.join(d, "rois.json")))
    H, W = rois["image_shape"]["height"], rois["image_shape"]["width"]
    imgs = {ch: load_u16(os.path.join(d, f"ch{ch}.u16.xz"), (H, W)) for ch in (2, 3)}
    polys = [np.asarray(p, dtype=float) for p in rois["rois"] if len(p) >= 3]
    with open(os.path.join(d, "expected.csv"), newline="") as f:
        rows = list(csv.DictReader(f))
    bits = np.frombuffer(lzma.decompress(open(os.path.join(d, "mask.bits.xz"), "rb").read()),
                         dtype=np.uint8)
    mask = np.unpackbits(bits)[: H * W].reshape(H, W).astype(bool)
    return imgs, polys, rows, mask


def load_fa_rois():
    d = json.load(open(os.path.join(GOLD, "fa_rois.json")))
    return {k: (v["image_shape"], [np.asarray(p, dtype=float) for p in v["rois"]])
            for k, v in d.items()}


def load_ref_vectors():
    return np.load(os.path.join(GOLD, "ref_vectors.npz"), allow_pickle=False)


STAT_KEYS = ("mean", "median", "std", "p5", "p95", "vmin", "vmax", "vsum", "npx")
