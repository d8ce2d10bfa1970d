"""Seeded synthetic inputs for the benchmark configurations.

* whiteboard_frames : uint8 HxWx3 BGR "whiteboard-like" (or chalkboard) lecture frames with a temporal
  script -- content accumulates, a board region is erased from time to time, a speaker-sized occluder sweeps.
* glyph_masks       : dense-handwriting binary masks (one small glyph per 20x20 cell, >5k CCs per 1080p
  frame) with per-frame churn and the same occluder, for the CC-only configuration.

There are no datasets or trained weights on the box: every benchmark says data = "synthetic".
"""
import cv2
import numpy as np

_COLOURS = [(0, 0, 0), (0, 0, 200), (200, 0, 0), (0, 140, 0)]          # BGR: black / red / blue / green


def _stroke(rng, h, w):
    pts = np.empty((6, 2), dtype=np.int32)
    pts[0] = (rng.integers(20, w - 20), rng.integers(20, h - 20))
    for i in range(1, 6):
        pts[i] = pts[i - 1] + rng.integers(-12, 13, size=2)
    pts[:, 0] = np.clip(pts[:, 0], 2, w - 3)
    pts[:, 1] = np.clip(pts[:, 1], 2, h - 3)
    return pts


def whiteboard_frames(n_frames, height, width, seed=1234, chalk=False, strokes_per_frame=2,
                      erase_every=600, occluder_w=300, occluder_step=30):
    """Yield n_frames uint8 (H, W, 3) BGR frames."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    if chalk:
        base = 40.0 + 8.0 * xx / width + 6.0 * yy / height
    else:
        base = 235.0 - 25.0 * xx / width - 15.0 * yy / height
    base = np.repeat(base[:, :, None], 3, axis=2)
    board = np.zeros((height, width, 3), dtype=np.uint8)             # ink layer
    ink = np.zeros((height, width), dtype=np.uint8)
    for t in range(n_frames):
        for _ in range(strokes_per_frame):
            pts = _stroke(rng, height, width)
            colour = (230, 230, 230) if chalk else _COLOURS[int(rng.integers(0, len(_COLOURS)))]
            thick = int(rng.integers(2, 4))
            cv2.polylines(board, [pts.reshape(-1, 1, 2)], False, colour, thick, cv2.LINE_AA)
            cv2.polylines(ink, [pts.reshape(-1, 1, 2)], False, 255, thick, cv2.LINE_AA)
        if erase_every and t > 0 and t % erase_every == 0:
            x0 = int(rng.integers(0, max(1, width - width // 3)))
            y0 = int(rng.integers(0, max(1, height - height // 2)))
            board[y0:y0 + height // 2, x0:x0 + width // 3] = 0
            ink[y0:y0 + height // 2, x0:x0 + width // 3] = 0
        a = (ink.astype(np.float32) / 255.0)[:, :, None]
        frame = base * (1.0 - a) + board.astype(np.float32) * a
        frame += rng.normal(0.0, 3.0, size=frame.shape).astype(np.float32)
        ox = (t * occluder_step) % (width + occluder_w) - occluder_w
        x0, x1 = max(0, ox), min(width, ox + occluder_w)
        if x1 > x0:
            frame[height // 5:, x0:x1] = (90.0, 70.0, 60.0)           # the "speaker"
        yield np.clip(frame, 0, 255).astype(np.uint8)


def _glyph(rng, mask, cx, cy):
    pts = np.stack([rng.integers(cx + 3, cx + 17, size=4), rng.integers(cy + 3, cy + 17, size=4)], axis=1).astype(np.int32)
    mask[cy:cy + 20, cx:cx + 20] = 0
    cv2.polylines(mask, [pts.reshape(-1, 1, 2)], False, 255, 2)


def glyph_masks(n_frames, height, width, seed=0, churn=0.03, occluder_w=300, occluder_step=30):
    """Yield n_frames uint8 (H, W) masks, ink = 255."""
    rng = np.random.default_rng(seed)
    cells = [(x, y) for y in range(0, height - 19, 20) for x in range(0, width - 19, 20)]
    mask = np.zeros((height, width), dtype=np.uint8)
    for cx, cy in cells:
        _glyph(rng, mask, cx, cy)
    for t in range(n_frames):
        if t > 0:
            for i in rng.choice(len(cells), size=max(1, int(churn * len(cells))), replace=False):
                _glyph(rng, mask, *cells[int(i)])
        out = mask.copy()
        if occluder_w:
            ox = (t * occluder_step) % (width + occluder_w) - occluder_w
            x0, x1 = max(0, ox), min(width, ox + occluder_w)
            if x1 > x0:
                out[height // 5:, x0:x1] = 0
        yield out


def random_blob_masks(n_frames, height, width, seed=0, density=0.45, jitter=0.02):
    """Small-case masks for parity tests: smoothed noise thresholded, with frame-to-frame pixel jitter,
    occasional blank frames and gaps (exercises multi-match, expiry and re-appearance)."""
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.random((height, width)).astype(np.float32), (0, 0), 1.6)
    thr = np.quantile(base, 1.0 - density * 0.5)
    core = base > thr
    for t in range(n_frames):
        if t % 17 == 11:
            yield np.zeros((height, width), dtype=np.uint8)
            continue
        m = core.copy()
        flip = rng.random((height, width)) < jitter
        m ^= flip
        if t % 5 == 3:
            m[:, (t * 7) % width:(t * 7) % width + width // 4] = False
        if t % 23 == 0 and t > 0:                                     # new content
            base = cv2.GaussianBlur(rng.random((height, width)).astype(np.float32), (0, 0), 1.6)
            core = base > np.quantile(base, 1.0 - density * 0.5)
        yield m.astype(np.uint8) * 255
