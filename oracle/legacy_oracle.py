"""oracle/legacy_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for the four legacy accessmath_lib exports).

Only tests/ and bench.py's CPU legs may import this module; the product (lecturemath_b200) never does.

NumPy restatement (IEEE fp64, one rounding per operation, the reference's left-to-right order) of
(R/ = /root/reference/ACCESS2021_release/):
  * region_cdf        <- regionCumulativeDistribution, R/accessmath_lib.c:113-173
  * adapthisteq       <- adapthisteq,                  R/accessmath_lib.c:175-329
  * combine_results   <- combine_results,              R/accessmath_lib.c:331-354
  * speaker_detection <- speaker_detection_handle_frame, R/accessmath_lib.c:7-111
and `ref()` = ctypes handle of oracle/_ref/accessmath_lib_ref.so, the reference C file compiled unmodified.

Parity status: pinned against oracle/_ref (tests/test_oracle_legacy.py) and against tests/golden/legacy_ops.npz,
which oracle/gen_golden_legacy.py produced by calling oracle/_ref.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "accessmath_lib_ref.so")


def ref():
    """The reference's own compiled C (None when oracle/_ref was not built)."""
    return ctypes.CDLL(REF_SO) if os.path.exists(REF_SO) else None


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _round_u8(v):
    """C round() (half away from zero) then the x86-64 (unsigned char) conversion: low byte of the int32 value."""
    t = np.trunc(v)
    t = np.where(np.abs(v - t) >= 0.5, t + np.copysign(1.0, v), t)
    ok = np.abs(t) < 2147483648.0
    return (np.where(ok, t, 0.0).astype(np.int64) & 0xFF).astype(np.uint8)


def region_cdf(gray, min_x, max_x, min_y, max_y, slope_max):
    hist = np.bincount(gray[min_y:max_y + 1, min_x:max_x + 1].ravel(), minlength=256)      # :126-136
    cum = np.cumsum(hist)                                                                   # :139-144
    with np.errstate(invalid="ignore", divide="ignore"):
        out = cum.astype(np.float64) / np.float64(cum[-1])                                  # :148-151
    if slope_max > 0.0:                                                                     # :154-172
        dh = np.float64(0.0)
        s = np.float64(slope_max)
        for i in range(255):
            diff = out[i + 1] - out[i] - dh - s
            dh = dh + (np.float64(0.0) if diff < 0.0 else diff)
            out[i + 1] = out[i + 1] - dh
        add = (np.float64(1.0) - (out[255] - out[0])) / np.float64(2.0)
        out = out + add
    return out


def _limits(size, grid):
    mn, mx, mid = [], [], []
    step, mod, start = size // grid, size % grid, 0
    for r in range(grid):
        end = start + step + (1 if r < mod else 0) - 1
        mn.append(start); mx.append(end)
        h = (start + end) / 2.0
        mid.append(int(np.floor(h + 0.5)) if h >= 0 else int(np.ceil(h - 0.5)))             # C round(), :203/:211
        start = end + 1
    return np.array(mn), np.array(mx), np.array(mid)


def adapthisteq(gray, slope, grid_x, grid_y):
    H, W = gray.shape
    xmn, xmx, xmid = _limits(W, grid_x)
    ymn, ymx, ymid = _limits(H, grid_y)
    dist = np.empty((grid_y, grid_x, 256), np.float64)
    for ry in range(grid_y):
        for rx in range(grid_x):
            dist[ry, rx] = region_cdf(gray, xmn[rx], xmx[rx], ymn[ry], ymx[ry], slope)
    cx = np.searchsorted(xmx, np.arange(W))[None, :].repeat(H, 0)                           # :241-243
    cy = np.searchsorted(ymx, np.arange(H))[:, None].repeat(W, 1)
    X = np.arange(W)[None, :].repeat(H, 0)
    Y = np.arange(H)[:, None].repeat(W, 1)
    tone = gray.astype(np.int64)
    edge_x = ((cx == 0) & (X <= xmid[cx])) | ((cx == grid_x - 1) & (X >= xmid[cx]))         # :256-257
    edge_y = ((cy == 0) & (Y <= ymid[cy])) | ((cy == grid_y - 1) & (Y >= ymid[cy]))         # :260-261
    x0 = np.clip(cx - (X <= xmid[cx]), 0, max(grid_x - 2, 0)); x1 = np.minimum(x0 + 1, grid_x - 1)
    y0 = np.clip(cy - (Y <= ymid[cy]), 0, max(grid_y - 2, 0)); y1 = np.minimum(y0 + 1, grid_y - 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        wx1 = (X - xmid[x0]).astype(np.float64) / (xmid[x1] - xmid[x0]).astype(np.float64)
        wy1 = (Y - ymid[y0]).astype(np.float64) / (ymid[y1] - ymid[y0]).astype(np.float64)
    one = np.float64(1.0)
    corner = dist[cy, cx, tone]                                                             # :264
    with np.errstate(invalid="ignore"):                                                     # unselected branches may be NaN
        vert = dist[y0, cx, tone] * (one - wy1) + dist[y1, cx, tone] * wy1                  # :276
        horz = dist[cy, x0, tone] * (one - wx1) + dist[cy, x1, tone] * wx1                  # :290
        full = (dist[y0, x0, tone] * (one - wx1) * (one - wy1) + dist[y1, x0, tone] * (one - wx1) * wy1 +
                dist[y0, x1, tone] * wx1 * (one - wy1) + dist[y1, x1, tone] * wx1 * wy1)    # :306-309
    v = np.where(edge_x & edge_y, corner, np.where(edge_x, vert, np.where(edge_y, horz, full)))
    return _round_u8(v * np.float64(255.0))


def combine_results(only_board, equalized, threshold):
    return np.where(only_board > 128, 0, np.where(equalized < threshold, 255, 0)).astype(np.uint8)      # :341-345


def speaker_detection(frame, last_frame, threshold, jump_cells):
    """-> (total_changes, boundaries[4], avg[2], deviation[2]) as float64 arrays."""
    H, W = frame.shape[:2]
    f = frame.reshape(H, W, -1).astype(np.int32)
    l = last_frame.reshape(H, W, -1).astype(np.int32)
    changed = np.zeros((H, W), bool)
    changed[::jump_cells, ::jump_cells] = (np.abs(l - f) > threshold).any(axis=2)[::jump_cells, ::jump_cells]   # :35-48
    cnt_x = changed.sum(axis=0).astype(np.float64)
    cnt_y = changed.sum(axis=1).astype(np.float64)
    total = int(changed.sum())
    cols, rows = np.nonzero(cnt_x)[0], np.nonzero(cnt_y)[0]
    bounds = np.array([cols.min() if total else W + 1, cols.max() if total else -1,
                       rows.min() if total else H + 1, rows.max() if total else -1], np.float64)                 # :78-81
    avg = np.zeros(2, np.float64); dev = np.zeros(2, np.float64)
    if total > 0:
        avg[0] = np.float64(int((np.arange(W) * cnt_x).sum())) / np.float64(total)            # integer sums: exact
        avg[1] = np.float64(int((np.arange(H) * cnt_y).sum())) / np.float64(total)
        dx = np.float64(0.0)
        for c in range(W):                                                                    # :92-94, sequential order
            d = np.float64(c) - avg[0]
            dx = dx + d * d * cnt_x[c]
        dy = np.float64(0.0)
        for r in range(H):                                                                    # :96-98
            d = np.float64(r) - avg[1]
            dy = dy + d * d * cnt_y[r]
        dev[0] = np.sqrt(dx / np.float64(total)); dev[1] = np.sqrt(dy / np.float64(total))    # :100-104
    return total, bounds, avg, dev


# ---- the same four operators through oracle/_ref (the reference itself) ---------------------------------------
def ref_region_cdf(lib, gray, min_x, max_x, min_y, max_y, slope_max):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.zeros(256, np.float64)
    fn = lib.regionCumulativeDistribution
    fn.restype = None
    fn.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_void_p]
    fn(_p(gray), gray.shape[1], gray.shape[0], min_x, max_x, min_y, max_y, slope_max, _p(out))
    return out


def ref_adapthisteq(lib, gray, slope, grid_x, grid_y):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.zeros_like(gray)
    fn = lib.adapthisteq
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    fn(_p(gray), gray.shape[1], gray.shape[0], slope, grid_x, grid_y, _p(out))
    return out


def ref_combine_results(lib, only_board, equalized, threshold):
    only_board = np.ascontiguousarray(only_board, np.uint8); equalized = np.ascontiguousarray(equalized, np.uint8)
    out = np.zeros_like(equalized)
    fn = lib.combine_results
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_ubyte, ctypes.c_void_p]
    fn(_p(only_board), _p(equalized), equalized.shape[1], equalized.shape[0], threshold, _p(out))
    return out


def ref_speaker_detection(lib, frame, last_frame, threshold, jump_cells):
    frame = np.ascontiguousarray(frame, np.uint8); last_frame = np.ascontiguousarray(last_frame, np.uint8)
    H, W = frame.shape[:2]
    C = frame.size // (H * W)
    b, a, d = np.zeros(4, np.float64), np.zeros(2, np.float64), np.zeros(2, np.float64)
    fn = lib.speaker_detection_handle_frame
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p] * 3
    total = fn(_p(frame), _p(last_frame), W, H, C, threshold, jump_cells, _p(b), _p(a), _p(d))
    return total, b, a, d
