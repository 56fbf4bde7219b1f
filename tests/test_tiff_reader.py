"""host/common.read_image_raw: the direct reader for plain uncompressed grey-scale TIFFs returns
exactly what PIL (the reference's fallback reader, Fluor_INT.py:350-362) decodes, and steps aside
for everything else (compressed, multi-sample, big-endian handled, tiled, BigTIFF)."""
import os
import struct

import numpy as np
import pytest
from PIL import Image

from imageprocess_b200.host import common


def _pil(path):
    with Image.open(path) as im:
        return np.array(im)


@pytest.mark.parametrize("name,dtype,shape,kw,plain", [
    ("u16", np.uint16, (300, 517), {}, True),
    ("u8", np.uint8, (120, 77), {}, True),
    ("f32", np.float32, (64, 65), {}, True),
    ("u16_lzw", np.uint16, (300, 517), {"compression": "tiff_lzw"}, False),
    ("u16_deflate", np.uint16, (90, 33), {"compression": "tiff_adobe_deflate"}, False),
    ("u16_1row", np.uint16, (1, 7), {}, True),
])
def test_plain_reader_equals_pil(tmp_path, name, dtype, shape, kw, plain):
    rng = np.random.default_rng(5)
    arr = (rng.random(shape) * (1 if dtype == np.float32 else np.iinfo(dtype).max)).astype(dtype)
    p = str(tmp_path / f"{name}.tif")
    Image.fromarray(arr).save(p, **kw)                       # PIL writes many 64 KiB strips, back to back
    assert (common._tiff_plain_layout(p) is not None) == plain
    got = common.read_image_raw(p)
    assert got.dtype == _pil(p).dtype and np.array_equal(got, _pil(p)) and np.array_equal(got, arr)
    assert common.image_shape(p) == shape


def test_own_writer_round_trip_and_big_endian(tmp_path):
    rng = np.random.default_rng(6)
    arr = rng.integers(0, 65535, (37, 53), dtype=np.uint16)
    p = str(tmp_path / "w.tif")
    common.write_tiff(p, arr)
    assert common._tiff_plain_layout(p) is not None and np.array_equal(common.read_image_raw(p), arr)
    assert np.array_equal(_pil(p), arr)
    # the same image as a big-endian TIFF, written by hand
    h, w = arr.shape
    data = arr.astype(">u2").tobytes()
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8),
            (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(data))]
    pb = str(tmp_path / "be.tif")
    with open(pb, "wb") as f:
        f.write(struct.pack(">2sHI", b"MM", 42, 8 + len(data)))
        f.write(data)
        f.write(struct.pack(">H", len(tags)))
        for tag, typ, cnt, val in tags:
            f.write(struct.pack(">HHI", tag, typ, cnt))
            f.write(struct.pack(">HH", val, 0) if typ == 3 else struct.pack(">I", val))
        f.write(struct.pack(">I", 0))
    got = common.read_image_raw(pb)
    assert got.dtype == np.uint16 and np.array_equal(got, arr) and np.array_equal(_pil(pb).astype(np.uint16), arr)


def test_rgb_and_truncated_go_to_pil(tmp_path):
    rgb = np.zeros((8, 9, 3), np.uint8)
    rgb[..., 0] = 7
    p = str(tmp_path / "rgb.tif")
    Image.fromarray(rgb).save(p)
    assert common._tiff_plain_layout(p) is None
    assert np.array_equal(common.read_image_raw(p), rgb[..., 0])       # first sample, as before
    # a plain TIFF cut short: the layout says more bytes than the file holds -> not the direct path's answer
    arr = np.arange(64 * 64, dtype=np.uint16).reshape(64, 64)
    q = str(tmp_path / "cut.tif")
    common.write_tiff(q, arr)
    raw = open(q, "rb").read()
    ifd = struct.unpack("<I", raw[4:8])[0]
    cut = raw[:8] + raw[8: 8 + 1000] + raw[ifd:]                        # pixel data shortened, IFD kept (offsets now lie)
    with open(q, "wb") as f:
        f.write(cut[:4] + struct.pack("<I", 8 + 1000) + cut[8:])
    with pytest.raises(Exception):
        common.read_image_raw(q)


def test_read_plane_into_pinned_like_buffer(tmp_path):
    """File -> caller's buffer without an array in between (the stream's pinned ring); anything that
    is not a plain little-endian uint16 raster of the buffer's shape is refused untouched."""
    rng = np.random.default_rng(7)
    arr = rng.integers(0, 65535, (40, 56), dtype=np.uint16)
    p = str(tmp_path / "a.tif")
    Image.fromarray(arr).save(p)
    ring = np.zeros((3, 2, 40, 56), np.uint16)
    assert common.read_plane_into(p, ring[1, 0]) and np.array_equal(ring[1, 0], arr) and not ring[1, 1].any()
    assert not common.read_plane_into(p, np.zeros((40, 57), np.uint16))               # shape differs
    assert not common.read_plane_into(p, np.zeros((40, 56), np.float32))              # dtype differs
    assert not common.read_plane_into(p, np.zeros((40, 112), np.uint16)[:, ::2])      # not contiguous
    q = str(tmp_path / "lzw.tif")
    Image.fromarray(arr).save(q, compression="tiff_lzw")
    assert not common.read_plane_into(q, ring[0, 0]) and not ring[0, 0].any()
    q8 = str(tmp_path / "u8.tif")
    Image.fromarray(arr.astype(np.uint8)).save(q8)
    assert not common.read_plane_into(q8, ring[0, 0])
