"""Pickle interchange with the reference's stage CLIs ("Python surface -- CC").

The reference writes `tempo_stability_*.dat` = pickle of the estimator object graph (R/pre_ST3D_v3.0_02_cc_analaysis.py:43) and
stage 03 unpickles it (R/pre_ST3D_v3.0_03_cc_grouping.py:22-23).  pickle stores classes by module path:
    AccessMath.preprocessing.content.cc_stability_estimator.CCStabilityEstimator
    AM_CommonTools.data.connected_component.ConnectedComponent
    AccessMath.preprocessing.tools.interval_index.IntervalIndex
Two directions:

* dump_reference_pickle(est, file): writes this package's estimator under THOSE paths with exactly the attributes the reference's
  __init__ / add_frame leave behind (cc_stability_estimator.py:11-31; ConnectedComponent connected_component.py:25-41 with `img`
  materialised), so an UNMODIFIED reference stage 03 / 04 process loads it as its own class and calls its own methods on it.
* install_aliases(): registers this package's classes under the reference's module paths when the reference itself is not
  importable, so that `pickle.load` of a file written by the reference's stage 02 gives objects with this package's (device)
  stage-03 methods.  Never shadows a genuine `AccessMath` / `AM_CommonTools` installation.
"""
import copyreg
import importlib
import io
import pickle
import sys
import types

import numpy as np

EST_PATH = ("AccessMath.preprocessing.content.cc_stability_estimator", "CCStabilityEstimator")
CC_PATH = ("AM_CommonTools.data.connected_component", "ConnectedComponent")
IDX_PATH = ("AccessMath.preprocessing.tools.interval_index", "IntervalIndex")


def _genuine(path):
    """The reference's own class when its package is importable in this process (and is not one of our aliases), else None."""
    mod, name = path
    try:
        m = importlib.import_module(mod)
    except Exception:
        return None
    if getattr(m, "__lecturemath_b200_alias__", False):
        return None
    return getattr(m, name, None)


def _ensure_module(name):
    """sys.modules entry (and parent chain) for a dotted module name; created modules are marked as aliases."""
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        full = ".".join(parts[:i])
        if full not in sys.modules:
            m = types.ModuleType(full)
            m.__lecturemath_b200_alias__ = True
            m.__path__ = []
            sys.modules[full] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
    return sys.modules[name]


def install_aliases():
    """Make reference-written pickles loadable with this package's classes (see module docstring).  Returns the list of module
    paths that were aliased (empty when the genuine reference is importable)."""
    from .cc_stability_estimator import CCStabilityEstimator
    from .connected_component import ConnectedComponent
    done = []
    for (mod, name), cls in ((EST_PATH, CCStabilityEstimator), (CC_PATH, ConnectedComponent), (IDX_PATH, IntervalIndexState)):
        if _genuine((mod, name)) is not None:
            continue
        setattr(_ensure_module(mod), name, cls)
        done.append(mod)
    return done


class IntervalIndexState:
    """Data-only stand-in for the reference's IntervalIndex(only_data=True) (interval_index.py:15-40): `intervals` is
    {start: {end: [data, ...]}} with one (possibly empty) dict per integer position 0 .. max end."""

    def __init__(self, only_data=True):
        self.intervals = {}
        self.only_data = only_data

    def add(self, start, end, data):                       # interval_index.py:20-32
        while len(self.intervals) < end + 1:
            self.intervals[len(self.intervals)] = {}
        self.intervals[start].setdefault(end, []).append(data)


def active_uniques(est):
    """(cc_last_frame, cc_active) as the reference's add_frame leaves them (cc_stability_estimator.py:62-63, 104, 118-119, 127-143):
    last frame each unique was matched on; uniques not yet expired after the last frame's expiry pass, ascending."""
    last = [int(f[-1][0]) for f in est.unique_cc_frames]
    t = est.img_idx - 1
    if t <= 0:                                             # no expiry pass on frame 0
        return last, list(range(len(last)))
    # a unique expires in the pass of the first frame t' with t' - last >= max_gap, i.e. once t >= last + max_gap
    return last, [u for u, l in enumerate(last) if t - l < est.max_gap]


def reference_state(est):
    """The attribute dictionary of a reference CCStabilityEstimator in the state `est` is in (objects still this package's)."""
    last, active = active_uniques(est)
    ix, iy = IntervalIndexState(True), IntervalIndexState(True)
    for u in active:                                       # what stays in the two indices: the active uniques, in ascending order
        cc = est.unique_cc_objects[u]                       # (the empty per-position lists expired uniques leave behind are not kept)
        ix.add(int(cc.min_x), int(cc.max_x) + 1, u)
        iy.add(int(cc.min_y), int(cc.max_y) + 1, u)
    return {"width": est.width, "height": est.height, "min_recall": est.min_recall, "min_precision": est.min_precision,
            "max_gap": est.max_gap, "unique_cc_objects": est.unique_cc_objects, "unique_cc_frames": est.unique_cc_frames,
            "cc_idx_per_frame": list(est.cc_idx_per_frame), "cc_int_index_x": ix, "cc_int_index_y": iy,
            "fake_age": None if getattr(est, "fake_age", None) is None else np.zeros((est.height, est.width), dtype=np.float32),
            "img_idx": est.img_idx, "tempo_count": est.tempo_count, "cc_last_frame": last, "cc_active": active,
            "verbose": getattr(est, "verbose", False)}


def _cc_state(cc):
    return {"cc_id": cc.cc_id, "min_x": cc.min_x, "min_y": cc.min_y, "max_x": cc.max_x, "max_y": cc.max_y, "size": cc.size,
            "img": cc.img, "normalized": cc.normalized, "start_time": cc.start_time, "end_time": cc.end_time,
            "next_cc": cc.next_cc, "prev_cc": cc.prev_cc}


class _ReferencePickler(pickle.Pickler):
    """Writes this package's estimator / ConnectedComponent / index objects as instances of the reference's classes."""

    def __init__(self, file, targets, protocol=pickle.HIGHEST_PROTOCOL):
        super().__init__(file, protocol)
        self._targets = targets

    def reducer_override(self, obj):
        from .cc_stability_estimator import CCStabilityEstimator
        from .connected_component import ConnectedComponent
        if isinstance(obj, CCStabilityEstimator):
            return self._as(EST_PATH, reference_state(obj))
        if isinstance(obj, ConnectedComponent):
            return self._as(CC_PATH, _cc_state(obj))
        if isinstance(obj, IntervalIndexState):
            return self._as(IDX_PATH, dict(obj.__dict__))
        return NotImplemented

    def _as(self, path, state):
        # copyreg._reconstructor(cls, object, None) = object.__new__(cls): the stdlib's own reduction of a plain class (pickle
        # refuses copyreg.__newobj__ with a class other than the object's own)
        return copyreg._reconstructor, (self._targets[path], object, None), state


def dump_reference_pickle(est, file, protocol=pickle.HIGHEST_PROTOCOL):
    """pickle `est` -- this package's CCStabilityEstimator, or any object graph holding one, e.g. stage 02's hand-off tuple
    (frame_times, frame_indices, estimator) of R/pre_ST3D_v3.0_02_cc_analaysis.py:43 -- into `file` the way the reference would have
    (R/AM_CommonTools/util/misc_helper.py:157-163 uses HIGHEST_PROTOCOL), loadable by the unmodified reference."""
    targets, temp = {}, []
    for path in (EST_PATH, CC_PATH, IDX_PATH):
        cls = _genuine(path)
        if cls is None:                                    # no reference here: a name-only placeholder class under that path
            mod, name = path
            m = _ensure_module(mod)
            prev = getattr(m, name, None)
            cls = type(name, (), {"__module__": mod})
            setattr(m, name, cls)
            temp.append((m, name, prev))
        targets[path] = cls
    try:
        _ReferencePickler(file, targets, protocol).dump(est)
    finally:
        for m, name, prev in temp:                         # restore whatever install_aliases() had put there
            if prev is None:
                delattr(m, name)
            else:
                setattr(m, name, prev)


def dumps_reference_pickle(est, protocol=pickle.HIGHEST_PROTOCOL):
    buf = io.BytesIO()
    dump_reference_pickle(est, buf, protocol)
    return buf.getvalue()
