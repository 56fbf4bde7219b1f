"""File-name grammars of the FRET-side scripts (every reference script carries its own regex set and
the mirrors keep each script's own):

* parse_tokens            fret_ratio_builder.py:244-256 and MOR_by_ROI.py:55-83 (identical on every name tried:
                          tests/test_oracle_vs_reference.py): channel = trailing `_<n>` / `_ch<n>` / `_c<n>`;
                          stage = first S<n> anywhere; time = first t<n> anywhere.
* parse_tokens_delimited  Nesprin2_FRET_Builder.py:292-307: the S<n> / t<n> tokens must stand between `_` / `-` / the
                          ends of the name (`XS1_1.tif`, `S1t3_1.tif` have no stage there).
* parse_tokens_cropper    roi_channel_cropper.py:211-252 (`parse_stage_time` + `detect_channel`): delimited S / t of at
                          most three digits; channel = a delimited `ch<n>` / `c<n>` token anywhere, else the LAST
                          all-digit token (the time token's digits excluded in time-lapse mode)."""
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


def parse_tokens_delimited(basename, timelapse):
    """Nesprin2_FRET_Builder.parse_tokens (Nesprin2_FRET_Builder.py:292-307)."""
    name = os.path.splitext(basename)[0]
    ch = None
    m = re.search(r"(?:[_-](\d+)$)|(?:[_-](?:ch|c)(\d+)$)", name, flags=re.IGNORECASE)
    if m:
        ch = int(next(g for g in m.groups() if g is not None))
    ms = re.search(r"(?i)(?:^|[_-])S(\d+)(?=$|[_-])", name)
    s_num = int(ms.group(1)) if ms else None
    t_num = None
    if timelapse:
        mt = re.search(r"(?i)(?:^|[_-])t(\d+)(?=$|[_-])", name)
        t_num = int(mt.group(1)) if mt else None
    return s_num, t_num, ch


_CH_TOKEN = re.compile(r"(?i)(?:^|[_-])(ch|c)(\d{1,3})(?=$|[_-])")


def parse_tokens_cropper(basename, timelapse):
    """roi_channel_cropper.parse_stage_time + detect_channel (roi_channel_cropper.py:211-252) as one
    (stage number, time number, channel) triple."""
    name = os.path.splitext(basename)[0]
    ms = re.search(r"(?i)(?:^|[_-])S(\d{1,3})(?=$|[_-])", name)
    s_num = int(ms.group(1)) if ms else None
    mt = re.search(r"(?i)(?:^|[_-])t(\d{1,3})(?=$|[_-])", name) if timelapse else None
    t_num = int(mt.group(1)) if mt else None
    m = _CH_TOKEN.search(name)
    if m:
        return s_num, t_num, int(m.group(2))
    nums = [tok for tok in re.split(r"[_-]", name) if tok.isdigit()]
    if mt:
        nums = [v for v in nums if v != mt.group(1)]          # the time token's digits, leading zeros kept
    return s_num, t_num, (int(nums[-1]) if nums else None)


def roi_json_path(roi_dir, s, t_code, timelapse, legacy=True):
    """S01[_t00].json first, legacy S1[_t0].json second (not for the cropper, which only looks for the first:
    roi_channel_cropper.py:270-275)."""
    base = f"{s}_{t_code}" if (timelapse and t_code is not None) else s
    cands = [os.path.join(roi_dir, base + ".json")]
    ms = re.search(r"(\d+)", s) if legacy else None
    if ms:
        legacy = f"S{int(ms.group(1))}"
        if timelapse and t_code is not None:
            legacy += f"_t{int(re.search(r'(\d+)', t_code).group(1))}"
        cands.append(os.path.join(roi_dir, legacy + ".json"))
    for c in cands:
        if os.path.exists(c):
            return c
    return None


def load_roi_polys(roi_dir, s, t_code, timelapse, legacy=True):
    p = roi_json_path(roi_dir, s, t_code, timelapse, legacy)
    return common.load_roi_json(p) if p else None


def build_pairs_by_channel(files, timelapse, donor_ch, fret_ch, parse=parse_tokens):
    """[((Sxx, txx | None), donor path, fret path)] for keys that have both channels, in
    (stage, time) order (fret_ratio_builder.py:910-928, Nesprin2_FRET_Builder.py:1264-1285 with
    parse = parse_tokens_delimited)."""
    key2 = {}
    for p in files:
        s_num, t_num, ch = parse(os.path.basename(p), timelapse)
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
