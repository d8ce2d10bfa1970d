"""oracle/cc_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for the CC stage).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (lecturemath_b200) never does: it fails loudly when the CUDA library is missing.

CPU restatement of the reference's stage 02 (R/ = /root/reference/ACCESS2021_release/):
  * label4            <- scipy.ndimage.label as called at R/AccessMath/preprocessing/content/labeler.py:126
  * age_boundaries    <- CC_AgeBoundaries, R/accessmath_lib.c:357-413
  * extract_components<- Labeler.extractSpatioTemporalContent, labeler.py:117-191
  * overlap_measure   <- ConnectedComponent.getOverlapFMeasure, R/AM_CommonTools/data/connected_component.py:202-250
  * bbox_pairs        <- IntervalIndex.find_matches x2 + set intersection,
                         R/AccessMath/preprocessing/tools/interval_index.py:42-99, cc_stability_estimator.py:73-84
  * StabilityOracle   <- CCStabilityEstimator.__init__/add_frame/finish_processing,
                         R/AccessMath/preprocessing/content/cc_stability_estimator.py:11-31,41-155,158-164

Parity status: the reference has no tests or golden vectors.  This restatement is pinned against the
reference itself, imported in the build container by oracle/gen_golden.py (outputs committed under
tests/golden/), and against oracle/_ref/accessmath_lib_ref.so (the reference C file compiled unmodified).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MIN_CC_PIXELS = 20                      # labeler.py:22

_I32P = ctypes.POINTER(ctypes.c_int32)
_F32P = ctypes.POINTER(ctypes.c_float)
_U8P = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    so = os.path.join(_HERE, "_build", "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "cc_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def ref_lib():
    """The reference's own accessmath_lib.c compiled unmodified (None when never built)."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "accessmath_lib_ref.so")
        if not os.path.exists(p):
            build()
        if os.path.exists(p):
            _ref = ctypes.CDLL(p)
    return _ref


def label4(content):
    """4-connected labeling, raster-order numbering (labeler.py:126). Returns (int32 labels, n)."""
    content = np.ascontiguousarray(content)
    assert content.ndim == 2
    if content.dtype != np.uint8:
        content = (content != 0).astype(np.uint8)
    h, w = content.shape
    labels = np.empty((h, w), dtype=np.int32)
    n = lib().orc_label4(content.ctypes.data_as(_U8P), ctypes.c_int(w), ctypes.c_int(h), labels.ctypes.data_as(_I32P))
    return labels, int(n)


def _age_boundaries(fn, labels, ages, n):
    h, w = labels.shape
    outs = [np.zeros(n, dtype=np.int32) for _ in range(5)] + [np.zeros(n, dtype=np.float32)]
    fn(labels.ctypes.data_as(_I32P), ages.ctypes.data_as(_F32P), ctypes.c_int(w), ctypes.c_int(h), ctypes.c_int(n),
       *[o.ctypes.data_as(_I32P) for o in outs[:5]], outs[5].ctypes.data_as(_F32P))
    return tuple(outs)   # mins_y, maxs_y, mins_x, maxs_x, counts, ages


def age_boundaries(labels, ages, n):
    """Our C restatement of CC_AgeBoundaries (accessmath_lib.c:357-413)."""
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    ages = np.ascontiguousarray(ages, dtype=np.float32)
    return _age_boundaries(lib().orc_age_boundaries, labels, ages, n)


def age_boundaries_ref(labels, ages, n):
    """The reference's own compiled CC_AgeBoundaries (oracle/_ref)."""
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    ages = np.ascontiguousarray(ages, dtype=np.float32)
    return _age_boundaries(ref_lib().CC_AgeBoundaries, labels, ages, n)


class OracleCC:
    """Value type restating ConnectedComponent (connected_component.py:21-41), hot-path fields only."""
    __slots__ = ("cc_id", "min_x", "max_x", "min_y", "max_y", "size", "img", "start_time", "end_time")

    def __init__(self, cc_id, min_x, max_x, min_y, max_y, size, img):
        self.cc_id, self.min_x, self.max_x, self.min_y, self.max_y = cc_id, min_x, max_x, min_y, max_y
        self.size, self.img = size, img
        self.start_time = self.end_time = None


def extract_components(content, ages=None, filter_small=True, use_ref_lib=False):
    """Labeler.extractSpatioTemporalContent (labeler.py:117-191): label, stats, crops of CCs >= 20 px.

    cc_id is the raw label - 1, gaps included (labeler.py:171-187)."""
    labels, n = label4(content)
    if n == 0:
        return [], labels, n                                     # labeler.py:133-135
    if ages is None:
        ages = np.zeros(labels.shape, dtype=np.float32)            # cc_stability_estimator.py:22
    fn = age_boundaries_ref if use_ref_lib else age_boundaries
    mins_y, maxs_y, mins_x, maxs_x, counts, out_ages = fn(labels, ages, n)
    comps = []
    for cc_id in range(n):
        if not filter_small or counts[cc_id] >= MIN_CC_PIXELS:   # labeler.py:177
            x0, x1, y0, y1 = mins_x[cc_id], maxs_x[cc_id], mins_y[cc_id], maxs_y[cc_id]
            img = (labels[y0:y1 + 1, x0:x1 + 1] == cc_id + 1).astype(np.uint8) * 255   # labeler.py:183
            cc = OracleCC(cc_id, int(x0), int(x1), int(y0), int(y1), int(counts[cc_id]), img)
            cc.start_time = cc.end_time = out_ages[cc_id]
            comps.append(cc)
    return comps, labels, n


def overlap_measure(a, b):
    """getOverlapFMeasure(other, False, False) (connected_component.py:202-250) -> (recall, precision)."""
    if not (a.max_y >= b.min_y and b.max_y >= a.min_y and a.max_x >= b.min_x and b.max_x >= a.min_x):
        return 0.0, 0.0
    match = lib().orc_overlap_count(
        a.img.ctypes.data_as(_U8P), a.min_x, a.max_x, a.min_y, a.max_y,
        b.img.ctypes.data_as(_U8P), b.min_x, b.max_x, b.min_y, b.max_y)
    return match / float(a.size), match / float(b.size)          # IEEE fp64 (:239-240)


def bbox_pairs(cur, uniq, active):
    """Candidate (cur_idx, unique_idx) pairs = inclusive bbox overlap on both axes, sorted.

    Equals sorted(set_x & set_y) of cc_stability_estimator.py:73-84 (IntervalIndex with add(min, max+1)
    is exactly inclusive interval overlap, interval_index.py:42-99)."""
    if not cur or not active:
        return []
    c = np.array([(cc.min_x, cc.max_x, cc.min_y, cc.max_y) for cc in cur], dtype=np.int64)
    act = np.asarray(active, dtype=np.int64)
    u = np.array([(uniq[i].min_x, uniq[i].max_x, uniq[i].min_y, uniq[i].max_y) for i in active], dtype=np.int64)
    pairs = []
    # chunk over current CCs to bound the c x a boolean matrix
    step = max(1, 4_000_000 // max(1, len(active)))
    for s in range(0, len(cur), step):
        cc = c[s:s + step]
        ov = ((cc[:, None, 0] <= u[None, :, 1]) & (u[None, :, 0] <= cc[:, None, 1]) &
              (cc[:, None, 2] <= u[None, :, 3]) & (u[None, :, 2] <= cc[:, None, 3]))
        ci, ui = np.nonzero(ov)                                   # row-major => sorted by (cur, active position)
        pairs.extend(zip((ci + s).tolist(), act[ui].tolist()))
    return pairs            # active is ascending => sorted like `merged`


class StabilityOracle:
    """CCStabilityEstimator restated (cc_stability_estimator.py:11-31, 41-155, 158-164).

    State names follow the reference so that tests read alike: unique_cc_objects, unique_cc_frames
    [(frame, raw_label)], cc_idx_per_frame [(unique_idx, cc)], cc_last_frame, cc_active, img_idx, tempo_count."""

    def __init__(self, width, height, min_recall, min_precision, max_gap, use_ref_lib=False):
        self.width, self.height = width, height
        self.min_recall, self.min_precision, self.max_gap = min_recall, min_precision, max_gap
        self.unique_cc_objects, self.unique_cc_frames, self.cc_idx_per_frame = [], [], []
        self.cc_last_frame, self.cc_active = [], []
        self.img_idx = 0
        self.tempo_count = 0
        self.use_ref_lib = use_ref_lib

    def get_raw_cc_count(self):                                   # :33-39
        return sum(len(f) for f in self.cc_idx_per_frame)

    def add_frame(self, binary):
        current_cc, _, _ = extract_components(binary, None, True, self.use_ref_lib)   # :50
        current = []
        if self.img_idx == 0:                                     # :52-69
            for cc in current_cc:
                self._new_unique(cc, current)
        else:
            merged = bbox_pairs(current_cc, self.unique_cc_objects, self.cc_active)   # :73-84
            self.tempo_count += len(merged)                       # :85
            nxt = 0
            for cc_idx, cc in enumerate(current_cc):              # :90-124
                found = False
                while nxt < len(merged) and merged[nxt][0] == cc_idx:
                    if not found:
                        prev_idx = merged[nxt][1]
                        recall, precision = overlap_measure(cc, self.unique_cc_objects[prev_idx])
                        if recall >= self.min_recall and precision >= self.min_precision:
                            found = True
                            self.unique_cc_frames[prev_idx].append((self.img_idx, cc.cc_id + 1))
                            current.append((prev_idx, cc))
                            self.cc_last_frame[prev_idx] = self.img_idx
                    nxt += 1
                if not found:
                    self._new_unique(cc, current)
            # expiry (:127-145)
            self.cc_active = [u for u in self.cc_active if self.img_idx - self.cc_last_frame[u] < self.max_gap]
        self.cc_idx_per_frame.append(current)                     # :150
        self.img_idx += 1

    def _new_unique(self, cc, current):
        self.unique_cc_objects.append(cc)
        self.unique_cc_frames.append([(self.img_idx, cc.cc_id + 1)])
        idx = len(self.unique_cc_objects) - 1
        current.append((idx, cc))
        self.cc_last_frame.append(self.img_idx)
        self.cc_active.append(idx)

    def finish_processing(self):                                  # :158-164
        return self.tempo_count

    # canonical tables used by the parity tests ------------------------------------------------
    def frame_table(self, t):
        """[(unique_idx, raw_label, min_x, max_x, min_y, max_y, size)] of frame t."""
        return [(u, cc.cc_id + 1, cc.min_x, cc.max_x, cc.min_y, cc.max_y, cc.size) for u, cc in self.cc_idx_per_frame[t]]
