"""Seeded synthetic inputs of SURVEY.md 8(d): FRET donor/acceptor frames with star-polygon
cell ROIs and focal-adhesion-like blobs (configs C3/C4) and the large FA mosaic (C5).

numpy only (host); bench.py uploads the result once, outside the timed region.
"""
import numpy as np


def star_polygon(rng, cx, cy, r_min, r_max, n_vert):
    """Random star-shaped polygon around (cx, cy); vertices snapped to the .5 grid the
    reference's manual ROI drawer produces (roi_manual_drawer.py:1316-1324)."""
    ang = np.sort(rng.uniform(0.0, 2.0 * np.pi, n_vert))
    rad = rng.uniform(r_min, r_max, n_vert)
    x = cx + rad * np.cos(ang)
    y = cy + rad * np.sin(ang)
    P = np.stack([x, y], axis=1)
    return np.round(P * 2.0) / 2.0


def place_cells(rng, H, W, n_cells, r_min, r_max, v_min=20, v_max=40, max_tries=20000):
    """Non-overlapping star polygons (bounding circles do not intersect)."""
    cells, centres = [], []
    tries = 0
    while len(cells) < n_cells and tries < max_tries:
        tries += 1
        cx = rng.uniform(r_max + 2, W - r_max - 2)
        cy = rng.uniform(r_max + 2, H - r_max - 2)
        if any((cx - ox) ** 2 + (cy - oy) ** 2 < (2 * r_max + 4) ** 2 for ox, oy in centres):
            continue
        centres.append((cx, cy))
        cells.append(star_polygon(rng, cx, cy, r_min, r_max, int(rng.integers(v_min, v_max + 1))))
    return cells, centres


def _fill_poly_mask(P, H, W):
    """Even-odd scanline fill at pixel centres; only used to paint synthetic signal (it
    is NOT the ROI rule under test)."""
    m = np.zeros((H, W), dtype=bool)
    x, y = P[:, 0], P[:, 1]
    y0 = max(int(np.floor(y.min())), 0)
    y1 = min(int(np.ceil(y.max())), H - 1)
    xn, yn = np.roll(x, -1), np.roll(y, -1)
    for r in range(y0, y1 + 1):
        cond = (y <= r) != (yn <= r)
        if not cond.any():
            continue
        xs = np.sort(x[cond] + (r - y[cond]) * (xn[cond] - x[cond]) / (yn[cond] - y[cond]))
        for a, b in zip(xs[0::2], xs[1::2]):
            ia, ib = max(int(np.ceil(a)), 0), min(int(np.floor(b)), W - 1)
            if ib >= ia:
                m[r, ia:ib + 1] = True
    return m


def add_blobs(rng, img, cell_mask, centre, radius, n_blobs, amp_lambda=6000.0,
              area_min=120, area_max=800):
    """FA-like elliptical blobs inside one cell (C4: 60 per cell, 120-800 px)."""
    H, W = img.shape
    cx, cy = centre
    for _ in range(n_blobs):
        area = rng.uniform(area_min, area_max)
        ar = rng.uniform(1.5, 4.0)
        b = np.sqrt(area / (np.pi * ar))
        a = ar * b
        th = rng.uniform(0, np.pi)
        rr = rng.uniform(0.2, 0.8) * radius
        ph = rng.uniform(0, 2 * np.pi)
        bx, by = cx + rr * np.cos(ph), cy + rr * np.sin(ph)
        R = int(np.ceil(a)) + 1
        x0, x1 = max(int(bx) - R, 0), min(int(bx) + R + 1, W)
        y0, y1 = max(int(by) - R, 0), min(int(by) + R + 1, H)
        if x0 >= x1 or y0 >= y1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        u = (xx - bx) * np.cos(th) + (yy - by) * np.sin(th)
        v = -(xx - bx) * np.sin(th) + (yy - by) * np.cos(th)
        e = ((u / a) ** 2 + (v / b) ** 2 <= 1.0) & cell_mask[y0:y1, x0:x1]
        n = int(e.sum())
        if n:
            sub = img[y0:y1, x0:x1]
            sub[e] = np.minimum(sub[e].astype(np.int64) + rng.poisson(amp_lambda, n), 65535)


def add_blobs_spread(rng, img, cell_mask, centre, radius, n_blobs, amp_lambda=6000.0,
                     area_min=120, area_max=800, gap=3, tries=40):
    """FA-like elliptical blobs spread over the WHOLE cell and kept apart: a blob is placed only
    where its ellipse, grown by `gap` pixels, meets no earlier blob and lies inside the cell, so a
    closing with disk(1) cannot fuse neighbours and every placed blob stays one adhesion.  Blobs
    that find no room after `tries` draws are dropped.  Returns the number placed."""
    H, W = img.shape
    cx, cy = centre
    ys, xs = np.nonzero(cell_mask)
    if ys.size == 0:
        return 0
    y_lo, y_hi, x_lo, x_hi = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    occ = np.zeros((y_hi - y_lo, x_hi - x_lo), dtype=bool)
    cm = cell_mask[y_lo:y_hi, x_lo:x_hi]
    placed = 0
    for _ in range(n_blobs):
        area = rng.uniform(area_min, area_max)
        ar = rng.uniform(1.5, 4.0)
        b = np.sqrt(area / (np.pi * ar))
        a = ar * b
        for _t in range(tries):
            th = rng.uniform(0, np.pi)
            k = int(rng.integers(0, ys.size))                    # a uniformly drawn cell pixel as the centre
            bx, by = xs[k] + rng.uniform(-0.5, 0.5), ys[k] + rng.uniform(-0.5, 0.5)
            R = int(np.ceil(a)) + gap + 1
            x0, x1 = int(bx) - R, int(bx) + R + 1
            y0, y1 = int(by) - R, int(by) + R + 1
            if x0 < x_lo or y0 < y_lo or x1 > x_hi or y1 > y_hi:
                continue
            yy, xx = np.mgrid[y0:y1, x0:x1]
            u = (xx - bx) * np.cos(th) + (yy - by) * np.sin(th)
            v = -(xx - bx) * np.sin(th) + (yy - by) * np.cos(th)
            grown = (u / (a + gap)) ** 2 + (v / (b + gap)) ** 2 <= 1.0
            sl = (slice(y0 - y_lo, y1 - y_lo), slice(x0 - x_lo, x1 - x_lo))
            if (grown & (occ[sl] | ~cm[sl])).any():
                continue
            e = (u / a) ** 2 + (v / b) ** 2 <= 1.0
            n = int(e.sum())
            if n < area_min:
                continue
            occ[sl] |= e
            sub = img[y0:y1, x0:x1]
            sub[e] = np.minimum(sub[e].astype(np.int64) + rng.poisson(amp_lambda, n), 65535)
            placed += 1
            break
    return placed


def fret_frame(seed=1234, H=2048, W=2048, n_cells=24, r_min=80, r_max=160, blobs_per_cell=0,
               drift=1.0, cells=None, centres=None, sat_frac=1e-4, blob_area=(120, 800), blob_layout="centre",
               info=None):
    """One C3/C4 frame.  Returns (donor u16 HxW, acceptor u16 HxW, polys list[(V,2) f64]).
    Channel 1 (donor) additionally carries FA blobs when blobs_per_cell > 0: blob_layout "centre"
    draws them around the cell centre (they overlap into a few large adhesions; the layout of the
    small parity scenes), "spread" keeps them apart all over the cell (add_blobs_spread).
    info (a dict) receives "blobs_placed"."""
    rng = np.random.default_rng(seed)
    if cells is None:
        cells, centres = place_cells(np.random.default_rng(seed ^ 0x5EED), H, W, n_cells,
                                     r_min, r_max)
    donor = (rng.poisson(300.0 * drift, (H, W)) + 100).astype(np.int64)
    acc = (rng.poisson(300.0 * drift, (H, W)) + 100).astype(np.int64)
    for P, c in zip(cells, centres):
        m = _fill_poly_mask(P, H, W)
        n = int(m.sum())
        sig = rng.poisson(2500.0 * drift, n)
        r = rng.uniform(0.3, 1.5)
        donor[m] += sig
        acc[m] += np.round(r * sig).astype(np.int64) + rng.poisson(300.0, n)
        if blobs_per_cell:
            d16 = np.minimum(donor, 65535)
            if blob_layout == "spread":
                k = add_blobs_spread(rng, d16, m, c, r_min, blobs_per_cell, area_min=blob_area[0], area_max=blob_area[1])
                if info is not None:
                    info["blobs_placed"] = info.get("blobs_placed", 0) + k
            else:
                add_blobs(rng, d16, m, c, r_min, blobs_per_cell, area_min=blob_area[0],
                          area_max=blob_area[1])
            donor = d16
    donor = np.minimum(donor, 65535)
    acc = np.minimum(acc, 65535)
    if sat_frac > 0:
        k = int(round(sat_frac * H * W))
        if k:
            idx = rng.choice(H * W, size=k, replace=False)
            donor.ravel()[idx[: k // 2]] = 65535
            acc.ravel()[idx[k // 2:]] = 65535
    return donor.astype(np.uint16), acc.astype(np.uint16), [np.array(P) for P in cells]


def fa_cells_frame(seed, H, W, polys, blobs_per_cell=40, info=None):
    """C2-style FA image for GIVEN cell outlines (the shipped 2200x3200 ROI JSONs have no images,
    SURVEY.md 8(d)): Poisson background, cell signal inside every polygon, spread adhesion blobs."""
    rng = np.random.default_rng(seed)
    img = (rng.poisson(300.0, (H, W)) + 100).astype(np.int64)
    for P in polys:
        P = np.asarray(P, dtype=float)
        m = _fill_poly_mask(P, H, W)
        n = int(m.sum())
        if n == 0:
            continue
        img[m] += rng.poisson(2500.0, n)
        d16 = np.minimum(img, 65535)
        k = add_blobs_spread(rng, d16, m, (P[:, 0].mean(), P[:, 1].mean()), 0.0, blobs_per_cell)
        if info is not None:
            info["blobs_placed"] = info.get("blobs_placed", 0) + k
        img = d16
    return np.minimum(img, 65535).astype(np.uint16)


def fa_mosaic(seed=99, H=8192, W=8192, n_blobs=100000, inset=8):
    """C5: Poisson(500) background, ~n_blobs elliptical blobs, one rectangular ROI."""
    rng = np.random.default_rng(seed)
    img = rng.poisson(500.0, (H, W)).astype(np.uint16)
    full = np.ones((1, 1), dtype=bool)
    side = int(np.sqrt(n_blobs))
    gy, gx = H / side, W / side
    for i in range(side):
        for j in range(side):
            cx = (j + 0.5) * gx + rng.uniform(-0.15, 0.15) * gx
            cy = (i + 0.5) * gy + rng.uniform(-0.15, 0.15) * gy
            area = rng.uniform(120, min(800, 0.25 * gx * gy))
            ar = rng.uniform(1.2, 3.0)
            b = np.sqrt(area / (np.pi * ar))
            a = ar * b
            th = rng.uniform(0, np.pi)
            R = int(np.ceil(a)) + 1
            x0, x1 = max(int(cx) - R, 0), min(int(cx) + R + 1, W)
            y0, y1 = max(int(cy) - R, 0), min(int(cy) + R + 1, H)
            yy, xx = np.mgrid[y0:y1, x0:x1]
            u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
            v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
            e = (u / a) ** 2 + (v / b) ** 2 <= 1.0
            n = int(e.sum())
            if n:
                sub = img[y0:y1, x0:x1]
                sub[e] = np.minimum(sub[e].astype(np.int64) + rng.poisson(6000.0, n), 65535)
    del full
    poly = np.array([[inset, inset], [W - 1 - inset, inset], [W - 1 - inset, H - 1 - inset],
                     [inset, H - 1 - inset]], dtype=float)
    return img, [poly]
