"""Image filters of the ROI drawer's display pipeline and the optional pre-filter stages
(SURVEY.md 8(f) item 4, section 0.1): Gaussian, band-pass, unsharp mask, white top-hat, Otsu.

Every function takes and returns device float32 images through the Engine; results equal
scipy.ndimage's bit for bit (the kernels replay its float64 arithmetic).  The reference calls
these on the host: roi_manual_drawer.py:870-876 (display only).  Gaussian / top-hat / Otsu as
analysis stages do not exist in the reference (north_star names them): they are optional and OFF
by default, parity "unpinned by reference" -- their oracle is scipy.ndimage itself.
"""
import numpy as np


def gaussian_kernel1d(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(truncate * sigma + 0.5)."""
    sigma = float(sigma)
    radius = int(truncate * sigma + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return phi, radius


def gaussian_filter(eng, img, sigma, out=None):
    """ndi.gaussian_filter(img, sigma) for device float32 images [n][H][W] (or [H][W])."""
    mem = eng.mem
    shape = tuple(img.shape)
    H, W = shape[-2], shape[-1]
    n = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    phi, radius = gaussian_kernel1d(sigma)
    if radius == 0:
        if out is None:
            out = mem.empty(shape, np.float32)
        mem.copy_bytes(out, 0, img, 0, img.nbytes)
        return out
    w = mem.from_host(np.ascontiguousarray(phi[radius:], dtype=np.float64))      # centre, then offsets 1..radius
    tmp = mem.empty(shape, np.float32)
    if out is None:
        out = mem.empty(shape, np.float32)
    eng.call("ipb_gaussian_f32", img.ptr, tmp.ptr, out.ptr, n, H, W, w.ptr, radius, mem.stream)
    out._keep = (w, tmp)
    return out


def bandpass(eng, img, sigma_small, sigma_large):
    """gaussian_filter(im, sigma_small) - gaussian_filter(im, sigma_large)   (roi_manual_drawer.py:873)."""
    a, b = gaussian_filter(eng, img, sigma_small), gaussian_filter(eng, img, sigma_large)
    out = eng.mem.empty(tuple(img.shape), np.float32)
    eng.call("ipb_gauss_combine", a.ptr, b.ptr, out.ptr, int(np.prod(img.shape, dtype=np.int64)), 0, 0.0, eng.mem.stream)
    out._keep = (a, b)
    return out


def unsharp(eng, img, amount, radius):
    """im + amount * (im - gaussian_filter(im, radius))   (roi_manual_drawer.py:875)."""
    g = gaussian_filter(eng, img, radius)
    out = eng.mem.empty(tuple(img.shape), np.float32)
    eng.call("ipb_gauss_combine", img.ptr, g.ptr, out.ptr, int(np.prod(img.shape, dtype=np.int64)), 1, float(np.float32(amount)),
             eng.mem.stream)
    out._keep = g
    return out


def render_pipeline(eng, img, use_bandpass=False, sigma_small=1.2, sigma_large=9.0, use_unsharp=False,
                    unsharp_amount=0.7, unsharp_radius=2.0):
    """_render_pipeline of the ROI drawer (roi_manual_drawer.py:870-876) with its default settings
    (:1455-1456) on a device float32 image."""
    im = img
    if use_bandpass:
        im = bandpass(eng, im, sigma_small, sigma_large)
    if use_unsharp:
        im = unsharp(eng, im, unsharp_amount, unsharp_radius)
    return im


def _planes_shape(planes):
    shape = tuple(planes.shape)
    H, W = shape[-2], shape[-1]
    n = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    return shape, n, H, W


def grey_morph(eng, planes, size, dilate):
    """ndi.grey_dilation / grey_erosion(planes, size=(size, size)) per uint16 plane [..][H][W] (odd size)."""
    size = int(size)
    if size < 1 or size % 2 == 0:
        raise ValueError("grey_morph: odd footprint sizes only")
    shape, n, H, W = _planes_shape(planes)
    mem = eng.mem
    tmp, out = mem.empty(shape, np.uint16), mem.empty(shape, np.uint16)
    eng.call("ipb_graymorph_u16", planes.ptr, tmp.ptr, out.ptr, n, H, W, size // 2, int(bool(dilate)), mem.stream)
    out._keep = tmp
    return out


def white_tophat(eng, planes, size):
    """ndi.white_tophat(plane, size=size) per uint16 plane: plane - dilation(erosion(plane)).  Optional
    pre-filter stage (default OFF; the reference has none): small bright structures on a slowly
    varying background, e.g. before the FA threshold."""
    shape, n, H, W = _planes_shape(planes)
    opened = grey_morph(eng, grey_morph(eng, planes, size, False), size, True)
    out = eng.mem.empty(shape, np.uint16)
    eng.call("ipb_sub_u16", planes.ptr, opened.ptr, out.ptr, int(np.prod(shape, dtype=np.int64)), eng.mem.stream)
    out._keep = opened
    return out


def threshold_otsu(eng, planes, H, W, plane_indices):
    """skimage.filters.threshold_otsu of uint16 planes: integer images are histogrammed one bin per
    value between their minimum and maximum, and the threshold is the bin that maximises the
    between-class variance.  The exact 65536-bin histograms come from the device (ipb_hist_u16); the
    65536-term cumulative sums are host numpy, as for the hist-mode background.  Optional stage
    (default OFF; the reference thresholds at mean + alpha * std, FA_Analyzer.py:143-146)."""
    from .ops import HIST_JOB, PAT_FULL
    jobs = np.zeros(len(plane_indices), dtype=HIST_JOB)
    jobs["plane"], jobs["pattern"] = np.asarray(plane_indices, dtype=np.int32), PAT_FULL
    hh = eng.hist(planes, H, W, jobs).hist.host()
    return [otsu_from_counts(h) for h in hh]


def otsu_from_counts(counts):
    """threshold_otsu from an integer-valued histogram (counts[v] = pixels of value v)."""
    nz = np.flatnonzero(counts)
    if nz.size == 0:
        raise ValueError("threshold_otsu: empty image")
    lo, hi = int(nz[0]), int(nz[-1])
    if lo == hi:
        return lo                                        # single-valued image: skimage returns that value
    c = counts[lo: hi + 1].astype(np.float64)
    centers = np.arange(lo, hi + 1, dtype=np.float64)
    w1 = np.cumsum(c)
    w2 = np.cumsum(c[::-1])[::-1]
    m1 = np.cumsum(c * centers) / w1
    m2 = (np.cumsum((c * centers)[::-1]) / w2[::-1])[::-1]
    var12 = w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2
    return int(centers[int(np.argmax(var12))])


def gaussian_u16(eng, planes, sigma):
    """Optional Gaussian pre-filter of uint16 planes: rint(gaussian_filter(float32(plane), sigma))
    clipped to uint16 (the analysis kernels work on integer samples)."""
    shape, n, H, W = _planes_shape(planes)
    mem = eng.mem
    total = int(np.prod(shape, dtype=np.int64))
    f = mem.empty(shape, np.float32)
    eng.call("ipb_convert_planes", planes.ptr, f.ptr, total, 0, mem.stream)
    g = gaussian_filter(eng, f, sigma)
    out = mem.empty(shape, np.uint16)
    eng.call("ipb_convert_planes", g.ptr, out.ptr, total, 1, mem.stream)
    out._keep = (f, g)
    return out


def prefilter_planes(eng, planes, shape, channel, prefilter):
    """The optional pre-filter stage of the FA chain: `prefilter` = ("tophat", size) or
    ("gaussian", sigma) applied to `channel` of uint16 [F][C][H][W]; returns uint16 [F][1][H][W]."""
    F, C, H, W = (int(v) for v in shape)
    mem = eng.mem
    one = mem.empty((F, 1, H, W), np.uint16)
    for f in range(F):
        mem.copy_bytes(one, 2 * f * H * W, planes, 2 * (f * C + channel) * H * W, 2 * H * W)
    kind, arg = prefilter
    if kind == "tophat":
        return white_tophat(eng, one, int(arg))
    if kind == "gaussian":
        return gaussian_u16(eng, one, float(arg))
    raise ValueError(f"unknown pre-filter {kind!r}")
