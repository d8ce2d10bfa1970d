"""oracle/resize_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for the > 2.5 MP branch of FCN_LectureNet.binarize).

R/AccessMath/lecturenet_v1/FCN_lecturenet.py:434-437 halves images above 2.5 MP with `PIL.Image.resize(..., LANCZOS)` until
they fit, and :481-494 brings the thresholded masks back with `cv2.resize(..., INTER_NEAREST)`.  Both operators live in
third-party dependencies that the reference does not pin (Pillow, OpenCV; versions in this image: Pillow 12.2.0,
opencv-python-headless 4.13.0), so this module restates their PUBLISHED algorithms in numpy:

  * lanczos_coeffs / resample_u8 / lanczos_resize  <- Pillow src/libImaging/Resample.c: precompute_coeffs (double
    coefficients, window [center - support, center + support) rounded to int, normalised), normalize_coeffs_8bpc
    (fixed point, PRECISION_BITS = 32 - 8 - 2 = 22), ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc
    (accumulator starts at 1 << 21, result = clip8(acc >> 22), the horizontal pass is rounded to uint8 before the vertical).
  * halve_until_fits                                <- FCN_lecturenet.py:434-437 (int(w / 2), int(h / 2) per round)
  * nearest_offsets / nearest_resize                <- OpenCV modules/imgproc/src/resize.cpp resizeNN:
    src index = min(floor(dst * (1 / (dsize / (double) ssize))), ssize - 1) per axis.

Pinned by tests/golden/resize.npz (outputs of Pillow / OpenCV and of the unmodified reference's binarize() on a 2.6 MP
frame, oracle/gen_golden_resize.py) and, wherever Pillow / OpenCV are importable, directly against them in
tests/test_oracle_resize.py.  Only tests/, smoke() and bench.py's reference legs may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def lanczos_coeffs(in_size, out_size):
    """-> (bounds int32 [out][2] = (first tap, tap count), coeffs int32 [out][ksize] fixed point, ksize)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size          # box = (0, in_size) held as C floats
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img, bounds, kk, axis):
    """One resampling pass over `axis` of a uint8 array (any trailing channel dimension)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def lanczos_resize(img, out_w, out_h):
    """uint8 (H, W[, C]) -> (out_h, out_w[, C]) as PIL.Image.resize((out_w, out_h), LANCZOS): horizontal, then vertical."""
    h, w = img.shape[:2]
    out = img
    if out_w != w:
        bx, kx, _ = lanczos_coeffs(w, out_w)
        out = _pass(out, bx, kx, 1)
    if out_h != h:
        by, ky, _ = lanczos_coeffs(h, out_h)
        out = _pass(out, by, ky, 0)
    return np.ascontiguousarray(out)


def fcn_size(width, height):
    """Sizes visited by the 2.5 MP guard: [(w0, h0), (w1, h1), ...]; the FCN runs at the last one (:434-437)."""
    sizes = [(width, height)]
    while width * height > 2500000:
        width, height = int(width / 2), int(height / 2)
        sizes.append((width, height))
    return sizes


def halve_until_fits(img):
    for (w, h) in fcn_size(img.shape[1], img.shape[0])[1:]:
        img = lanczos_resize(img, w, h)
    return img


def nearest_offsets(src_size, dst_size):
    inv_scale = dst_size / float(src_size)
    ifx = 1.0 / inv_scale
    return np.minimum(np.floor(np.arange(dst_size, dtype=np.float64) * ifx).astype(np.int64), src_size - 1)


def nearest_resize(img, out_w, out_h):
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_NEAREST)."""
    h, w = img.shape[:2]
    return np.ascontiguousarray(img[nearest_offsets(h, out_h)][:, nearest_offsets(w, out_w)])


def cv2_linear_u8(src, out_w, out_h):
    """cv2.resize(src, (out_w, out_h)) with the default INTER_LINEAR on uint8 HxWxC -- OpenCV's 8-bit fixed-point algorithm
    (imgproc/resize.cpp: resizeGeneric_, HResizeLinear<uchar, int, short, 2048>, VResizeLinear for uchar), the third-party operator
    VideoProcessor.doProcessing applies when a resolution is forced (R/AccessMath/preprocessing/video_processor/video_processor.py:
    164-165).  Unpinned dependency (OpenCV 4.13.0 here).  Exact against cv2.resize wherever that build runs its generic code (all
    down-scales); its IPP path for up-scales differs by one grey level on ~0.1 % of the pixels (tests/test_oracle_resize.py)."""
    src = np.asarray(src)
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    ih, iw, _ = src.shape

    def axis(o, i):
        scale = 1.0 / (o / float(i))
        f = ((np.arange(o, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        lo, hi = s < 0, s >= i - 1
        f[lo], s[lo] = 0, 0
        f[hi], s[hi] = 0, i - 1
        c0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)        # saturate_cast<short>: round half to even
        c1 = np.rint(f * np.float32(2048)).astype(np.int64)
        return s, np.minimum(s + 1, i - 1), c0, c1

    sx, sx1, a0, a1 = axis(out_w, iw)
    sy, sy1, b0, b1 = axis(out_h, ih)
    S = src.astype(np.int64)
    rows = S[:, sx, :] * a0[None, :, None] + S[:, sx1, :] * a1[None, :, None]
    r0, r1 = rows[sy] >> 4, rows[sy1] >> 4
    out = (((r0 * b0[:, None, None]) >> 16) + ((r1 * b1[:, None, None]) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out
