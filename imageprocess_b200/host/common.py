"""Shared host plumbing of the mirrors: TIFF I/O, ROI JSON, table writers.  File-name grammars
live in each mirror because the reference scripts each have their own."""
import csv
import glob
import json
import os
import re
import struct

import numpy as np


def ensure_dir(p):
    os.makedirs(p, exist_ok=True)
    return p


def natural_key(s):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]


def list_tifs(folder):
    """Every *.tif / *.tiff of a folder once (case-insensitive file systems list a file under
    several patterns), natural order -- Fluor_INT.py:265-275."""
    seen = {}
    for pat in ("*.tif", "*.tiff", "*.TIF", "*.TIFF"):
        for p in glob.glob(os.path.join(folder, pat)):
            seen.setdefault(os.path.normcase(os.path.abspath(p)), p)
    return sorted(seen.values(), key=natural_key)


def fmt_stage(n):
    return f"S{int(n):02d}"


def fmt_time(n):
    return f"t{int(n):02d}"


# ---------------------------------------------------------------------- TIFF
def read_image_raw(path, page=0):
    """First page of a TIFF as a numpy array in its stored dtype (PIL: the reference's own
    fallback reader, Fluor_INT.py:350-362; it decodes the LZW uint16 fixtures identically)."""
    from PIL import Image
    with Image.open(path) as im:
        try:
            im.seek(page)
        except EOFError:
            im.seek(0)
        a = np.array(im)
    if a.ndim > 2:
        a = a[..., 0] if a.ndim == 3 else a[0, ...]
    return a


def image_shape(path):
    """(H, W) of an image file from its header, without decoding the pixels."""
    from PIL import Image
    with Image.open(path) as im:
        w, h = im.size
    return int(h), int(w)


def read_2d(path):
    """float32 view of read_image_raw -- what the reference's read_2d returns (Fluor_INT.py:364-368)."""
    return read_image_raw(path).astype(np.float32, copy=False)


def as_u16_plane(a, what="image"):
    """The device path works on the uint16 samples themselves.  Integer TIFFs pass through;
    anything else (float TIFFs) is refused loudly -- no CPU fallback."""
    a = np.asarray(a)
    if a.dtype == np.uint16:
        return np.ascontiguousarray(a)
    if a.dtype in (np.uint8, np.bool_):
        return a.astype(np.uint16)
    if np.issubdtype(a.dtype, np.integer) or np.issubdtype(a.dtype, np.floating):
        r = a.astype(np.uint16)
        if np.array_equal(r.astype(a.dtype), a):
            return r
    raise ValueError(f"{what}: only 8/16-bit unsigned integer samples are supported on the device path "
                     f"(got dtype {a.dtype} with non-integer or out-of-range values)")


_TIFF_TYPES = {np.dtype(np.uint8): (8, 1), np.dtype(np.uint16): (16, 1), np.dtype(np.float32): (32, 3),
               np.dtype(np.int32): (32, 2), np.dtype(np.uint32): (32, 1)}


def write_tiff(path, arr):
    """Baseline little-endian TIFF, one uncompressed strip, minisblack (what tifffile.imwrite
    produces for a 2-D array modulo metadata; SURVEY.md 8(f) item 1: no tifffile here)."""
    a = np.ascontiguousarray(arr)
    if a.ndim != 2:
        raise ValueError("write_tiff: 2-D arrays only")
    bits, fmt = _TIFF_TYPES[a.dtype]
    h, w = a.shape
    data = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, bits), (259, 3, 1, 1), (262, 3, 1, 1),
            (273, 4, 1, 8), (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(data)), (339, 3, 1, fmt)]
    ifd_off = 8 + len(data) + (len(data) & 1)
    with open(path, "wb") as f:
        f.write(struct.pack("<2sHI", b"II", 42, ifd_off))
        f.write(data)
        if len(data) & 1:
            f.write(b"\0")
        f.write(struct.pack("<H", len(tags)))
        for tag, typ, cnt, val in tags:
            f.write(struct.pack("<HHI", tag, typ, cnt))
            f.write(struct.pack("<HH", val, 0) if typ == 3 else struct.pack("<I", val))
        f.write(struct.pack("<I", 0))


# ---------------------------------------------------------------------- ROI JSON
def load_roi_json(path):
    """{"rois": [[[x, y], ...], ...]} -> list of (V, 2) float arrays; polygons with fewer than
    three points are dropped (Fluor_INT.py:417-422).  Returns None when nothing is usable."""
    with open(path, "r", encoding="utf-8") as f:
        data = json.load(f)
    polys = []
    for poly in data.get("rois", []):
        P = np.asarray(poly, dtype=float)
        if P.ndim == 2 and P.shape[0] >= 3:
            polys.append(P)
    return polys or None


def write_rows_csv(path, rows, columns=None):
    """pandas.DataFrame(rows).to_csv(index=False) when pandas is there (the reference's writer),
    else the csv module with the same column order."""
    if not rows:
        return
    try:
        import pandas as pd
        df = pd.DataFrame(rows)
        if columns:
            df = df[[c for c in columns if c in df.columns] + [c for c in df.columns if c not in columns]]
        df.to_csv(path, index=False)
    except ImportError:
        cols = list(columns or [])
        for r in rows:
            for k in r:
                if k not in cols:
                    cols.append(k)
        with open(path, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=cols)
            w.writeheader()
            w.writerows(rows)
