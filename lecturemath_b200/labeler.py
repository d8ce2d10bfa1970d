"""Labeler drop-in (R/AccessMath/preprocessing/content/labeler.py:20-191) running on the B200.

Same static-method surface: extractConnectedComponents(content, filter_small, is_labeled) and
extractSpatioTemporalContent(content, ages, filter_small, is_labeled) -> list[ConnectedComponent] in ascending
label order, cc_id = raw label - 1 (gaps of filtered CCs kept, labeler.py:171-187)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .cc_engine import CCEngine, MIN_CC_PIXELS as _MIN
from .connected_component import ConnectedComponent


class Labeler:
    MIN_CC_PIXELS = _MIN
    _engines = {}

    @staticmethod
    def _engine(width, height, filter_small):
        key = (width, height, bool(filter_small), torch.cuda.current_device())
        eng = Labeler._engines.get(key)
        if eng is None:
            eng = CCEngine(width, height, 1, Labeler.MIN_CC_PIXELS if filter_small else 1)
            Labeler._engines[key] = eng
        return eng

    @staticmethod
    def extractConnectedComponents(content, filter_small=True, is_labeled=False):
        return Labeler.extractSpatioTemporalContent(content, None, filter_small, is_labeled)   # labeler.py:107-111

    @staticmethod
    def extractSpatioTemporalContent(content, ages, filter_small=True, is_labeled=False):
        assert len(content.shape) == 2
        if is_labeled:
            raise NotImplementedError("is_labeled=True (pre-labeled input) is not on the stage-01/02 hot path")
        height, width = content.shape
        eng = Labeler._engine(width, height, filter_small)
        mask = torch.from_numpy(np.ascontiguousarray(content != 0).view(np.uint8)).cuda(non_blocking=True)
        bits = eng.pack(mask[None])
        need_ages = ages is not None and bool(np.any(ages))
        labels = eng.label(bits, want_labels=need_ages)
        n_labels, n_kept = int(eng.counts[0, 1]), int(eng.counts[0, 2])
        if n_labels == 0:
            return []                                                   # labeler.py:133-135
        rows, crops = eng.kept_rows(0), eng.crops(0)
        min_age = None
        if need_ages:                                                   # real ages: CC_AgeBoundaries on the resident label image
            n = n_labels
            d_ages = torch.from_numpy(np.ascontiguousarray(ages, dtype=np.float32)).cuda(non_blocking=True)
            d_out = torch.empty((6, n), dtype=torch.int32, device=labels.device)
            d_tmp = torch.empty(10 * n + 4, dtype=torch.int32, device=labels.device)
            _lib.check(eng.lib.am_cc_age_boundaries_dev(labels[0].data_ptr(), d_ages.data_ptr(), width, height, n, d_out.data_ptr(),
                                                        d_tmp.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "am_cc_age_boundaries_dev")
            min_age = d_out[5].cpu().numpy().view(np.float32)
        comps = []
        for i in range(n_kept):
            _, lab, x0, x1, y0, y1, size, off = (int(v) for v in rows[i])
            words = ((x1 >> 5) - (x0 >> 5) + 1) * (y1 - y0 + 1)
            cc = ConnectedComponent(lab - 1, np.int32(x0), np.int32(x1), np.int32(y0), np.int32(y1), np.int32(size),
                                    packed=crops[off:off + words])
            cc.start_time = cc.end_time = (np.float32(0.0) if min_age is None else min_age[lab - 1])
            comps.append(cc)
        return comps
