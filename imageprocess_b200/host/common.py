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
def _tiff_plain_layout(path):
    """(dtype, H, W, offset) when the first page of a classic TIFF is ONE uncompressed, contiguous,
    single-sample grey-scale raster (what microscopes, tifffile.imwrite and write_tiff below produce
    for 2-D uint8 / uint16 / float32 images), else None.  Header only: a few hundred bytes read."""
    try:
        with open(path, "rb") as f:
            head = f.read(8)
            if len(head) < 8 or head[:2] not in (b"II", b"MM"):
                return None
            e = "<" if head[:2] == b"II" else ">"
            magic, ifd = struct.unpack(e + "HI", head[2:8])
            if magic != 42:                                       # BigTIFF (43) and anything else: the general reader
                return None
            f.seek(ifd)
            n = struct.unpack(e + "H", f.read(2))[0]
            raw = f.read(12 * n)
            if len(raw) < 12 * n:
                return None
            size = {1: 1, 2: 1, 3: 2, 4: 4, 6: 1, 7: 1, 8: 2, 9: 4, 16: 8}
            code = {1: "B", 3: "H", 4: "I", 16: "Q"}
            tags = {}
            for i in range(n):
                tag, typ, cnt = struct.unpack(e + "HHI", raw[12 * i: 12 * i + 8])
                if typ not in code:
                    tags[tag] = None
                    continue
                nb = size[typ] * cnt
                if nb <= 4:
                    vals = struct.unpack(e + code[typ] * cnt, raw[12 * i + 8: 12 * i + 8 + nb])
                else:
                    off = struct.unpack(e + "I", raw[12 * i + 8: 12 * i + 12])[0]
                    if cnt > 1 << 20:
                        return None
                    f.seek(off)
                    vals = struct.unpack(e + code[typ] * cnt, f.read(nb))
                tags[tag] = vals
            one = lambda t, d=None: (tags[t][0] if tags.get(t) else d)
            w, h = one(256), one(257)
            if not w or not h or 322 in tags or 324 in tags:      # tiled
                return None
            if one(259, 1) != 1 or one(277, 1) != 1 or one(262, 1) != 1 or one(266, 1) != 1 or one(284, 1) != 1:
                return None                                       # compressed / RGB / white-is-zero / bit-reversed / planar
            bits, fmt = one(258, 1), one(339, 1)
            dt = {(8, 1): "u1", (16, 1): "u2", (32, 3): "f4"}.get((bits, fmt))
            offs, cnts = tags.get(273), tags.get(279)
            if dt is None or not offs or not cnts or len(offs) != len(cnts):
                return None
            if any(offs[i] + cnts[i] != offs[i + 1] for i in range(len(offs) - 1)):
                return None                                       # strips not back to back
            if sum(cnts) != w * h * (bits // 8):
                return None
            return np.dtype(e + dt if bits > 8 else dt), int(h), int(w), int(offs[0])
    except (OSError, struct.error, KeyError):
        return None


def read_image_raw(path, page=0):
    """First page of a TIFF as a numpy array in its stored dtype.  Plain uncompressed grey-scale
    rasters are read straight from the file (one read, no decoder: 8 MB in ~3 ms instead of ~25 ms
    through PIL's decoder and its copies -- the folder entry points are decode-bound); everything
    else goes through PIL, the reference's own fallback reader (Fluor_INT.py:350-362; it decodes
    the LZW uint16 fixtures identically).  tests: checks_host.check_plain_tiff_reader."""
    if page == 0:
        lay = _tiff_plain_layout(path)
        if lay is not None:
            dt, h, w, off = lay
            a = np.fromfile(path, dtype=dt, count=h * w, offset=off)
            if a.size == h * w:
                return a.reshape(h, w).astype(dt.newbyteorder("="), copy=False)
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


def read_plane_into(path, out):
    """Reads a plain uncompressed uint16 TIFF (see _tiff_plain_layout) straight into `out`, a
    C-contiguous uint16 [H][W] array -- the stream's page-locked ring buffer: file -> pinned memory
    with no array in between.  Returns False (nothing usable written) when the file is anything
    else or its shape differs; the caller then decodes it the general way."""
    lay = _tiff_plain_layout(path)
    if lay is None:
        return False
    dt, h, w, off = lay
    if dt != np.dtype("<u2") or out.dtype != np.uint16 or out.shape != (h, w) or not out.flags.c_contiguous:
        return False
    with open(path, "rb", buffering=0) as f:
        f.seek(off)
        view = memoryview(out).cast("B")
        got = 0
        while got < len(view):
            n = f.readinto(view[got:])
            if not n:
                return False
            got += n
    return True


def image_shape(path):
    """(H, W) of an image file from its header, without decoding the pixels."""
    lay = _tiff_plain_layout(path)
    if lay is not None:
        return lay[1], lay[2]
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
