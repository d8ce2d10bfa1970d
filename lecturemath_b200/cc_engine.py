"""Device-side CC engine: thin Python over the C ABI (ctx + estimator handles, torch tensors for memory).

One CCEngine owns the scratch for a batch of frames of a fixed size; `label()` runs labeling + stats + crops,
`match()` runs temporal matching for the frames of the batch, the `read_*` calls bring tables to the host."""
import ctypes
import os

import numpy as np
import torch

from . import _lib

MIN_CC_PIXELS = 20        # R/AccessMath/preprocessing/content/labeler.py:22
LABEL_LAUNCHES = 3            # am_cc_label_batch: k_strip_label, k_resolve, k_crop_fill (+1 with a label image)
MATCH_LAUNCHES_PER_FRAME = 6  # am_est_add_frames: k_match_pairs, _overlap, _select, _refresh, _update, _copy


def match_launches(n_frames):
    """Kernel launches of one am_est_add_frames call: ONE cooperative k_match_fused for the whole batch (default), or the six
    per-frame kernels with AM_B200_MATCH=multi."""
    return MATCH_LAUNCHES_PER_FRAME * n_frames if os.environ.get("AM_B200_MATCH") == "multi" else 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _np(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CCEngine:
    def __init__(self, width, height, max_batch=1, min_pixels=MIN_CC_PIXELS, max_runs=0, max_labels=0, max_kept=0,
                 crop_words=0, device=None):
        self.lib = _lib.lib()
        self.width, self.height, self.max_batch = int(width), int(height), int(max_batch)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.wpr = self.lib.am_words_per_row(self.width)
        self._caps = [int(max_runs), int(max_labels), int(max_kept), int(crop_words), int(min_pixels)]
        self._grow = 0                   # how often the default capacities have been doubled (label(sync=True) retries on overflow)
        self._create()
        self.counts = None

    def _create(self):
        runs, labels, kept, crops, min_pixels = self._caps
        with torch.cuda.device(self.device):
            self.ctx = self.lib.am_cc_create(self.width, self.height, self.max_batch, runs, labels, kept, crops, min_pixels)
        if not self.ctx:
            raise _lib.AccessMathB200Error("am_cc_create failed")

    def _enlarge(self):
        """Double the kept-CC and crop capacities (the reference has no such limits: many overlapping large bounding boxes -- long
        diagonal strokes, a board frame around everything -- need more crop words than the 4 x frame default)."""
        P = self.width * self.height
        self._grow += 1
        runs, labels, kept, crops, min_pixels = self._caps
        kept = min(P // 2 + 64, 2 * (kept if kept > 0 else P // max(1, min_pixels) + 64))
        crops = 2 * (crops if crops > 0 else 4 * self.wpr * self.height + 1024)
        self._caps = [runs, labels, kept, crops, min_pixels]
        self.close()
        self._create()

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.am_cc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- masks ---------------------------------------------------------------------------------
    def pack(self, masks_u8):
        """uint8 (B,H,W) device tensor (ink != 0) -> uint32-as-int32 (B,H,WPR) bit-packed tensor."""
        assert masks_u8.is_cuda and masks_u8.dtype == torch.uint8 and masks_u8.is_contiguous()
        b = masks_u8.shape[0]
        bits = torch.empty((b, self.height, self.wpr), dtype=torch.int32, device=masks_u8.device)
        _lib.check(self.lib.am_pack_mask_u8(_p(masks_u8), self.width, self.height, b, _p(bits), _stream()), "am_pack_mask_u8")
        return bits

    def unpack(self, bits):
        b = bits.shape[0]
        out = torch.empty((b, self.height, self.width), dtype=torch.uint8, device=bits.device)
        _lib.check(self.lib.am_unpack_mask_u8(_p(bits), self.width, self.height, b, _p(out), _stream()), "am_unpack_mask_u8")
        return out

    # ---- labeling + stats + crops ------------------------------------------------------------------
    def label(self, bits, want_labels=False, sync=True, out=None):
        b = bits.shape[0]
        assert b <= self.max_batch and bits.is_contiguous() and bits.shape[1:] == (self.height, self.wpr)
        labels = None
        if want_labels:
            labels = out if out is not None else torch.empty((b, self.height, self.width), dtype=torch.int32, device=bits.device)
            assert labels.is_contiguous() and labels.dtype == torch.int32 and labels.shape == (b, self.height, self.width)
        _lib.check(self.lib.am_cc_label_batch(self.ctx, _p(bits), b, _p(labels) if want_labels else None, _stream()),
                   "am_cc_label_batch")
        self.batch = b
        if sync:
            try:
                self.read_counts()
            except _lib.AccessMathB200Error as e:
                if "code 3" not in str(e) or self._grow >= 4:
                    raise
                self._enlarge()                                          # capacity: grow and label the same masks again
                return self.label(bits, want_labels, sync, out)
        return labels

    def read_counts(self):
        c = np.zeros((self.batch, 4), dtype=np.int32)
        _lib.check(self.lib.am_cc_counts(self.ctx, self.batch, _np(c), _stream()), "am_cc_counts")
        self.counts = c                  # n_runs, n_labels, n_kept, crop_words
        return c

    def label_table(self, f):
        n = int(self.counts[f, 1])
        t = [np.zeros(n, dtype=np.int32) for _ in range(5)]
        _lib.check(self.lib.am_cc_read_label_table(self.ctx, f, n, *[_np(a) for a in t], _stream()), "am_cc_read_label_table")
        return t                         # min_y, max_y, min_x, max_x, count

    def kept_rows(self, f):
        n = int(self.counts[f, 2])
        rows = np.zeros((n, 8), dtype=np.int32)
        _lib.check(self.lib.am_cc_read_kept(self.ctx, f, n, _np(rows), _stream()), "am_cc_read_kept")
        return rows                      # unique, label, min_x, max_x, min_y, max_y, size, crop_off

    def crops(self, f):
        n = int(self.counts[f, 3])
        w = np.zeros(n, dtype=np.uint32)
        _lib.check(self.lib.am_cc_read_crops(self.ctx, f, n, _np(w), _stream()), "am_cc_read_crops")
        return w

    def packed_rows(self, batch=None):
        """All kept rows of the batch in one device tensor + offsets (one D2H for the whole batch)."""
        b = batch or self.batch
        total_cap = int(self.counts[:b, 2].sum()) if self.counts is not None else 0
        rows = torch.empty((max(total_cap, 1), 8), dtype=torch.int32, device=self.device)
        offs = torch.empty((b + 1,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.am_cc_pack_rows(self.ctx, b, _p(rows), max(total_cap, 1), _p(offs), _stream()), "am_cc_pack_rows")
        return rows[:total_cap], offs


    def pack_rows_into(self, rows, offs, batch=None):
        """Asynchronous variant: kept rows of the batch into the preallocated `rows` (cap, 8) / `offs` (batch+1,) CUDA
        tensors; rows beyond the capacity are dropped (offs[batch] still holds the true total)."""
        b = batch or self.batch
        _lib.check(self.lib.am_cc_pack_rows(self.ctx, b, _p(rows), rows.shape[0], _p(offs), _stream()), "am_cc_pack_rows")


class Estimator:
    """Handle on the device-side temporal matcher (am_est_*)."""

    def __init__(self, width, height, min_recall, min_precision, max_gap, max_uniques=0, max_active=0, arena_words=0, device=None):
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self.h = self.lib.am_est_create(int(width), int(height), float(min_recall), float(min_precision), int(max_gap),
                                            int(max_uniques), int(max_active), int(arena_words))
        if not self.h:
            raise _lib.AccessMathB200Error("am_est_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.am_est_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_frames(self, engine, first, n):
        _lib.check(self.lib.am_est_add_frames(self.h, engine.ctx, first, n, _stream()), "am_est_add_frames")

    def state(self):
        s = np.zeros(6, dtype=np.int32)
        _lib.check(self.lib.am_est_state(self.h, _np(s), _stream()), "am_est_state")
        tempo = (int(s[4]) & 0xffffffff) | ((int(s[5]) & 0xffffffff) << 32)
        return {"n_unique": int(s[0]), "n_active": int(s[1]), "img_idx": int(s[2]), "tempo_count": tempo}

    def uniques(self, first, n):
        rows = np.zeros((n, 8), dtype=np.int32)
        _lib.check(self.lib.am_est_read_uniques(self.h, first, n, _np(rows), _stream()), "am_est_read_uniques")
        return rows                      # first_frame, first_label, min_x, max_x, min_y, max_y, size, last_seen

    def unique_crop(self, u, words):
        w = np.zeros(words, dtype=np.uint32)
        _lib.check(self.lib.am_est_read_unique_crop(self.h, u, words, _np(w), _stream()), "am_est_read_unique_crop")
        return w

    # ---- frame-shard hand-off ----------------------------------------------------------------------
    def export_state(self):
        """-> (header int64[6], meta int32 (n_act,10) device, crops int32 (words,) device)."""
        sizes = np.zeros(2, dtype=np.int64)
        _lib.check(self.lib.am_est_export_sizes(self.h, _np(sizes), _stream()), "am_est_export_sizes")
        n_act, words = int(sizes[0]), int(sizes[1])
        meta = torch.zeros((max(n_act, 1), 10), dtype=torch.int32, device=self.device)
        crops = torch.zeros((max(words, 1),), dtype=torch.int32, device=self.device)
        if n_act:
            _lib.check(self.lib.am_est_export(self.h, _p(meta), _p(crops), _stream()), "am_est_export")
        st = self.state()
        header = torch.tensor([n_act, words, st["n_unique"], st["img_idx"], st["tempo_count"], 0], dtype=torch.int64)
        return header, meta, crops

    def export_dev(self, buf):
        """Asynchronous: pack the active set into the fixed-capacity int32 CUDA tensor `buf` (no host sync)."""
        _lib.check(self.lib.am_est_export_dev(self.h, _p(buf), buf.numel(), _stream()), "am_est_export_dev")

    def import_dev(self, buf):
        """Asynchronous: replace this estimator's state by the active set packed in `buf`."""
        _lib.check(self.lib.am_est_import_dev(self.h, _p(buf), _stream()), "am_est_import_dev")

    def export_dev_ptr(self, ptr, words):
        """export_dev into raw device memory (e.g. the ring successor's mailbox mapped through CUDA IPC)."""
        _lib.check(self.lib.am_est_export_dev(self.h, ctypes.c_void_p(ptr), int(words), _stream()), "am_est_export_dev")

    def import_dev_ptr(self, ptr):
        _lib.check(self.lib.am_est_import_dev(self.h, ctypes.c_void_p(ptr), _stream()), "am_est_import_dev")

    def import_state(self, header, meta, crops):
        n_act, words, n_unique, img_idx, tempo = [int(v) for v in header[:5]]
        _lib.check(self.lib.am_est_import(self.h, n_act, n_unique, img_idx, tempo, _p(meta) if n_act else None,
                                          _p(crops) if n_act else None, words, _stream()), "am_est_import")
