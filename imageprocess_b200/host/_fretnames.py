"""File-name grammar shared by fret_ratio_builder, Nesprin2_FRET_Builder, MOR_by_ROI and
roi_channel_cropper (each reference script carries its own copy: fret_ratio_builder.py:244-256,
Nesprin2_FRET_Builder.py:292-307, MOR_by_ROI.py:55-83, roi_channel_cropper.py:211-252):
channel = trailing `_<n>` / `_ch<n>` / `_c<n>`; stage = first S<n>; time = first t<n>."""
import os
import re

from . import common
from .common import fmt_stage, fmt_time


def parse_tokens(basename, timelapse):
    name = os.path.splitext(basename)[0]
    ch = None
    m = re.search(r"(?:[_-](\d+)$)|(?:[_-](?:ch|c)(\d+)$)", name, flags=re.IGNORECASE)
    if m:
        ch = int(next(g for g in m.groups() if g is not None))
    ms = re.search(r"(?i)S(\d+)", name)
    s_num = int(ms.group(1)) if ms else None
    t_num = None
    if timelapse:
        mt = re.search(r"(?i)t(\d+)", name)
        t_num = int(mt.group(1)) if mt else None
    return s_num, t_num, ch


def roi_json_path(roi_dir, s, t_code, timelapse):
    """S01[_t00].json first, legacy S1[_t0].json second."""
    base = f"{s}_{t_code}" if (timelapse and t_code is not None) else s
    cands = [os.path.join(roi_dir, base + ".json")]
    ms = re.search(r"(\d+)", s)
    if ms:
        legacy = f"S{int(ms.group(1))}"
        if timelapse and t_code is not None:
            legacy += f"_t{int(re.search(r'(\d+)', t_code).group(1))}"
        cands.append(os.path.join(roi_dir, legacy + ".json"))
    for c in cands:
        if os.path.exists(c):
            return c
    return None


def load_roi_polys(roi_dir, s, t_code, timelapse):
    p = roi_json_path(roi_dir, s, t_code, timelapse)
    return common.load_roi_json(p) if p else None


def build_pairs_by_channel(files, timelapse, donor_ch, fret_ch):
    """[((Sxx, txx | None), donor path, fret path)] for keys that have both channels, in
    (stage, time) order (fret_ratio_builder.py:910-928, Nesprin2_FRET_Builder.py:309-330)."""
    key2 = {}
    for p in files:
        s_num, t_num, ch = parse_tokens(os.path.basename(p), timelapse)
        if s_num is None or ch is None:
            continue
        key = (fmt_stage(s_num), fmt_time(t_num) if (timelapse and t_num is not None) else None)
        key2.setdefault(key, {})[ch] = p

    def order(k):
        s, tt = k
        return (int(re.search(r"\d+", s).group()), int(re.search(r"\d+", tt).group()) if tt else -1)
    pairs = [(k, key2[k][donor_ch], key2[k][fret_ch]) for k in sorted(key2, key=order)
             if donor_ch in key2[k] and fret_ch in key2[k]]
    return pairs, key2
