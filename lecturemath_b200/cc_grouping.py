"""Stage 03 (CC grouping) behind the reference's estimator methods.

R/pre_ST3D_v3.0_03_cc_grouping.py:41-101 calls, on the estimator that stage 02 produced:
    rebuilt_binary_images, split_stable_cc_by_gaps, get_stable_cc_idxs, compute_overlapping_stable_cc, compute_groups,
    compute_groups_temporal_information, compute_conflicting_groups, compute_group_images, frames_from_groups
(R/AccessMath/preprocessing/content/cc_stability_estimator.py:166-681).  GroupingMixin gives CCStabilityEstimator the same
methods, same arguments, same result shapes.  All pixel work runs on the B200 over the bit-packed unique-CC crops that
stage 02 left in HBM (csrc/grouping.cu: am_group_overlaps, am_group_images, am_paint_frames); what stays here is the
order-dependent list / dictionary bookkeeping whose ORDER is part of the reference's result (group numbering, list order).
There is no CPU path for the pixel work: without the library / a device these methods raise."""
import bisect
import ctypes
import functools
import gc

import numpy as np
import torch

from . import _lib
from .connected_component import pack_crop


class UniqueView(ctypes.Structure):
    """am_unique_view (include/accessmath_b200.h)."""
    _fields_ = [("min_x", ctypes.c_void_p), ("max_x", ctypes.c_void_p), ("min_y", ctypes.c_void_p), ("max_y", ctypes.c_void_p),
                ("size", ctypes.c_void_p), ("crop_off", ctypes.c_void_p), ("arena", ctypes.c_void_p), ("n", ctypes.c_int)]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gc_paused(fn):
    """These methods build hundreds of thousands of long-lived lists / tuples / objects (the reference's result shapes) and no reference
    cycles; with the cyclic collector running, its full passes over the growing heap were 100 of the 262 ms of stage 03 on the 32-frame
    dense workload (tools/grouping_bench.py --gc-report).  Paused for the duration of the call, restored afterwards."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        was = gc.isenabled()
        gc.disable()
        try:
            return fn(*args, **kwargs)
        finally:
            if was:
                gc.enable()
    return wrapper


def _dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def _packed(cc):
    if getattr(cc, "_packed", None) is None:
        cc._packed = pack_crop(cc.img, int(cc.min_x), int(cc.max_x), int(cc.min_y), int(cc.max_y))
    return cc._packed


def _crop_words(cc):
    return ((int(cc.max_x) >> 5) - (int(cc.min_x) >> 5) + 1) * (int(cc.max_y) - int(cc.min_y) + 1)


def view_from_components(ccs):
    """am_unique_view over host objects (min_x, max_x, min_y, max_y, size, img / packed crop): their boxes and bit-packed crops are
    uploaded once.  -> (UniqueView, tensors that must stay alive as long as the view is used)."""
    words = np.array([_crop_words(c) for c in ccs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(words)]).astype(np.uint64)
    arena = np.concatenate([_packed(c) for c in ccs]) if ccs else np.zeros(1, np.uint32)
    cols = [_dev([int(getattr(c, k)) for c in ccs], np.int32) for k in ("min_x", "max_x", "min_y", "max_y", "size")]
    keep = cols + [_dev(offs[:-1], np.uint64), _dev(arena.view(np.int32), np.int32)]
    view = UniqueView()
    for name, t in zip(("min_x", "max_x", "min_y", "max_y", "size", "crop_off", "arena"), keep):
        setattr(view, name, t.data_ptr())
    view.n = len(ccs)
    return view, keep


def overlap_pairs(view, d_ids, n, timed=None):
    """int64 [n_pairs][3] = (a, b, matched pixels) for all positions a < b of the n listed view rows whose inclusive bounding boxes
    intersect, ascending (am_group_overlaps)."""
    lib = _lib.lib()
    cap = max(1024, 8 * n)
    while True:
        pairs = torch.empty((cap, 3), dtype=torch.int32, device="cuda")
        cnt = ctypes.c_longlong(0)
        args = (ctypes.byref(view), d_ids.data_ptr(), n, pairs.data_ptr(), cap, ctypes.byref(cnt), _stream())
        rc = timed("am_group_overlaps", lib.am_group_overlaps, *args) if timed else lib.am_group_overlaps(*args)
        if rc == 3:
            cap = int(cnt.value)
            continue
        _lib.check(rc, "am_group_overlaps")
        return pairs[:cnt.value].cpu().numpy().astype(np.int64)


def frame_segment_items(groups_per_frame, group_ages, seg_index, n_groups):
    """int32 [n][2] = (frame, segment row) of every group alive in every frame, in (frame, list position) order.  Segment s of group g
    at frame t = the first s with marks[s + 1] >= t (the reference advances a per-group cursor frame by frame, :655-660); found for
    all pairs at once in the concatenated, (g, mark)-sorted age marks.  seg_index: {(g, s): row}, rows of one group consecutive."""
    import itertools
    alive = np.fromiter((len(p) for p in groups_per_frame), dtype=np.int64, count=len(groups_per_frame))
    n_items = int(alive.sum())
    items = np.zeros((n_items, 2), dtype=np.int32)
    if not n_items:
        return items
    g_of = np.fromiter(itertools.chain.from_iterable(groups_per_frame), dtype=np.int64, count=n_items)
    t_of = np.repeat(np.arange(len(groups_per_frame), dtype=np.int64), alive)
    gs = sorted(set(g_of.tolist()))
    n_marks = np.zeros(n_groups, dtype=np.int64)
    n_marks[gs] = [len(group_ages[g]) for g in gs]
    m_start = np.cumsum(n_marks) - n_marks
    marks = np.fromiter(itertools.chain.from_iterable(group_ages[g] for g in gs), dtype=np.int64, count=int(n_marks.sum()))
    big = int(max(int(marks.max()), len(groups_per_frame))) + 2
    keys = np.repeat(np.array(gs, dtype=np.int64), n_marks[gs]) * big + marks
    j = np.searchsorted(keys, g_of * big + t_of, side="left") - m_start[g_of]      # first mark >= t inside the group's marks
    seg = np.maximum(j - 1, 0)
    if int((seg + 1 >= n_marks[g_of]).sum()):
        raise IndexError("a frame lies beyond the last age mark of one of its groups")     # the reference's cursor would overrun too
    base = np.full(n_groups, -1, dtype=np.int64)
    for (g, s0), row in seg_index.items():
        if s0 == 0:
            base[g] = row
    items[:, 0], items[:, 1] = t_of, base[g_of] + seg
    return items


class GroupingMixin:
    def _timed(self, name, fn, *args):
        """Run one C-ABI call; when self.device_ms is a dict (tools/grouping_bench.py) accumulate its CUDA-event time there."""
        acc = getattr(self, "device_ms", None)
        if acc is None:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        e1.synchronize()
        acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
        return rc

    # ---- device view of the unique CCs ------------------------------------------------------------------------
    def _unique_view(self):
        """(UniqueView, obj_index): obj_index[u] = row of the view that holds unique u's box / crop.  Zero copy while the
        stage-02 estimator is alive on this device; otherwise (unpickled object) the packed crops are uploaded once."""
        n = len(self.unique_cc_objects)
        cached = getattr(self, "_view_cache", None)
        if cached is not None and cached[2] == n:
            return cached[0], cached[1]
        first, obj_index = {}, np.zeros(n, dtype=np.int32)
        for u, cc in enumerate(self.unique_cc_objects):               # split_stable_cc_by_gaps appends aliases of the same object
            obj_index[u] = first.setdefault(id(cc), u)
        lib = _lib.lib()
        view, keep = UniqueView(), None
        est = getattr(self, "_est", None)
        if est is not None and getattr(est, "h", None):
            _lib.check(lib.am_est_unique_view(est.h, ctypes.byref(view), _stream()), "am_est_unique_view")
            if view.n < int(obj_index.max(initial=-1)) + 1:
                raise _lib.AccessMathB200Error("device estimator holds %d uniques, the host state %d" % (view.n, n))
        else:
            ccs = [self.unique_cc_objects[u] for u in sorted(first.values())]
            rows = sorted(first.values())
            remap = {u: i for i, u in enumerate(rows)}
            obj_index = np.array([remap[int(v)] for v in obj_index], dtype=np.int32)
            view, keep = view_from_components(ccs)
        self._view_cache = (view, obj_index, n, keep)
        return view, obj_index

    # ---- :166-179 ---------------------------------------------------------------------------------------------
    @gc_paused
    def rebuilt_binary_images(self, chunk=64):
        out = []
        n_frames = len(self.cc_idx_per_frame)                         # (a live estimator builds its host view -- and the tables -- here)
        tables = getattr(self, "_frame_tables", None)
        fast = tables is not None and len(tables) == n_frames
        for f0 in range(0, n_frames, chunk):
            if fast:                                                  # the per-frame tables exactly as stage 02 read them back
                tb = tables[f0:f0 + chunk]
                base = np.concatenate([[0], np.cumsum([len(t[2]) for t in tb])])
                out.extend(self._paint_arrays(len(tb), np.repeat(np.arange(len(tb)), [len(t[0]) for t in tb]),
                                              np.concatenate([t[0] for t in tb]).reshape(-1, 4),
                                              np.concatenate([t[1] + np.uint64(b) for t, b in zip(tb, base[:-1])]),
                                              np.concatenate([t[2] for t in tb])))
            else:
                rows = self.cc_idx_per_frame[f0:f0 + chunk]
                ccs = [(t, cc) for t, row in enumerate(rows) for _, cc in row]
                out.extend(self._paint(len(rows), [t for t, _ in ccs], [c for _, c in ccs]))
        return out

    def rebuilt_binary_frame(self, frame_ccs):
        return self._paint(1, [0] * len(frame_ccs), [cc for _, cc in frame_ccs])[0]

    def _paint(self, n_frames, item_frame, ccs):
        """uint8 frames with `+= 255` at every set pixel of every (frame, crop) item -- am_paint_frames."""
        if not ccs:
            return self._paint_arrays(n_frames, np.zeros(0, np.int32), np.zeros((0, 4), np.int32), np.zeros(0, np.uint64), np.zeros(1, np.uint32))
        words = np.array([_crop_words(c) for c in ccs], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(words)])[:-1].astype(np.uint64)
        boxes = np.array([(int(c.min_x), int(c.max_x), int(c.min_y), int(c.max_y)) for c in ccs], dtype=np.int32)
        return self._paint_arrays(n_frames, item_frame, boxes, offs, np.concatenate([_packed(c) for c in ccs]))

    def _paint_arrays(self, n_frames, item_frame, boxes, offs, imgs):
        lib = _lib.lib()
        out = torch.empty(n_frames * self.height * self.width + 8, dtype=torch.uint8, device="cuda")
        n = len(boxes)
        if n:
            d_imgs = _dev(np.ascontiguousarray(imgs).view(np.int32), np.int32)
            d_boxes, d_frame = _dev(boxes, np.int32), _dev(item_frame, np.int32)
            d_img, d_off = _dev(np.arange(n), np.int32), _dev(offs, np.uint64)
            args = (n, d_frame.data_ptr(), d_img.data_ptr(), d_boxes.data_ptr(), d_off.data_ptr(), d_imgs.data_ptr())
        else:
            args = (0, None, None, None, None, None)
        _lib.check(self._timed("am_paint_frames", lib.am_paint_frames, *args, 0, n_frames, self.height, self.width, out.data_ptr(), _stream()),
                   "am_paint_frames")
        host = out[:n_frames * self.height * self.width].cpu().numpy().reshape(n_frames, self.height, self.width)
        return [host[t] for t in range(n_frames)]

    # ---- :181-228 (list bookkeeping; no arithmetic) ----------------------------------------------------------------
    @gc_paused
    def split_stable_cc_by_gaps(self, max_gap, stable_min_frames):
        split = 0
        objects, uframes, per_frame = self.unique_cc_objects, self.unique_cc_frames, self.cc_idx_per_frame
        for u in range(len(objects)):
            frames = uframes[u]
            if len(frames) < stable_min_frames or frames[-1][0] - frames[0][0] <= max_gap:
                continue                                                  # too short to matter / too compact to hold a gap > max_gap
            cuts = [i for i in range(1, len(frames)) if frames[i][0] - frames[i - 1][0] > max_gap]
            if not cuts:
                continue
            edges = [0] + cuts + [len(frames)]
            uframes[u] = frames[:cuts[0]]
            for a, b in zip(edges[1:-1], edges[2:]):
                new_u = len(objects)
                objects.append(objects[u])
                uframes.append(frames[a:b])
                for t, _ in frames[a:b]:
                    row = per_frame[t]
                    k = next((i for i, (uk, _) in enumerate(row) if uk == u), None)      # the FIRST instance with the old index
                    if k is not None:
                        row[k] = (new_u, row[k][1])
            split += 1
        return split

    def get_stable_cc_idxs(self, min_stable_frames):                         # :230-236
        return [u for u, f in enumerate(self.unique_cc_frames) if len(f) >= min_stable_frames]

    def get_temporal_index(self):                                            # :238-243
        pf = self.cc_idx_per_frame
        if hasattr(pf, "unique_indices"):                                        # live estimator: straight from the result rows
            return [pf.unique_indices(t) for t in range(len(pf))]
        return [[u for u, _ in row] for row in pf]

    # ---- :245-306 ---------------------------------------------------------------------------------------------
    def stable_overlaps_device(self, stable_idxs):
        """int64 array [n_pairs][3] = (idx_cc1, idx_cc2, matched pixels), idx_cc1 < idx_cc2 ascending: every pair of stable
        uniques whose inclusive bounding boxes intersect (am_group_overlaps)."""
        lib = _lib.lib()
        view, obj_index = self._unique_view()
        ids = np.asarray(stable_idxs, dtype=np.int64)
        if len(ids) < 2:
            return np.zeros((0, 3), dtype=np.int64)
        assert np.all(np.diff(ids) > 0), "stable_idxs must be ascending (get_stable_cc_idxs)"
        d_ids = _dev(obj_index[ids], np.int32)
        p = overlap_pairs(view, d_ids, len(ids), self._timed)
        p[:, 0], p[:, 1] = ids[p[:, 0]], ids[p[:, 1]]
        return p

    @gc_paused
    def compute_overlapping_stable_cc(self, stable_idxs, temporal_window):
        n = len(self.unique_cc_objects)
        all_ov, time_ov, total = [[] for _ in range(n)], [[] for _ in range(n)], 0
        p = self.stable_overlaps_device(stable_idxs)
        p = p[p[:, 2] > 0]                                                   # recall == precision == 0.0 otherwise (:283)
        if len(p) == 0:
            return time_ov, total, all_ov
        size = np.fromiter((int(c.size) for c in self.unique_cc_objects), dtype=np.int64, count=n)
        first = np.fromiter((f[0][0] for f in self.unique_cc_frames), dtype=np.int64, count=n)
        last = np.fromiter((f[-1][0] for f in self.unique_cc_frames), dtype=np.int64, count=n)
        s1, s2 = size[p[:, 0]], size[p[:, 1]]
        recall = p[:, 2] / s1.astype(np.float64)                             # connected_component.py:239-240 (IEEE fp64 divide)
        precision = p[:, 2] / s2.astype(np.float64)
        matched = (s1 * recall).astype(np.int64)                             # :284  int(cc1.size * recall): fp64 product, truncated
        in_time = (last[p[:, 0]] + temporal_window >= first[p[:, 1]]) & (last[p[:, 1]] >= first[p[:, 0]] - temporal_window)
        for u1, u2, m, a, b, r, q, t in zip(p[:, 0].tolist(), p[:, 1].tolist(), matched.tolist(), s1.tolist(), s2.tolist(),
                                            recall.tolist(), precision.tolist(), in_time.tolist()):
            all_ov[u1].append((u2, m, b, a))
            all_ov[u2].append((u1, m, a, b))
            if t:
                time_ov[u1].append((u2, r, q))
                time_ov[u2].append((u1, q, r))
                total += 1
        return time_ov, total, all_ov

    # ---- :308-413 (sequential merge; the numbering of the groups is part of the result) ---------------------------------
    @gc_paused
    def compute_groups(self, stable_idxs, overlapping_cc, min_recall, t_fmeasure, t_time_IOU):
        groups, owner = [], {}
        for u1 in stable_idxs:
            g = owner.get(u1)
            if g is None:
                g = owner[u1] = len(groups)
                groups.append([u1])
            for u2, recall, _ in overlapping_cc[u1]:
                if recall < min_recall:
                    continue
                other = owner.get(u2)
                if other is None:
                    owner[u2] = g
                    groups[g].append(u2)
                elif other != g:
                    for m in groups[other]:
                        owner[m] = g
                    groups[g].extend(groups[other])
                    groups[other] = []
        final = [grp for grp in groups if grp]
        return final, {m: g for g, grp in enumerate(final) for m in grp}

    # ---- :415-444 ---------------------------------------------------------------------------------------------
    @gc_paused
    def compute_groups_temporal_information(self, cc_groups):
        n_frames = len(self.cc_idx_per_frame)
        uframes = self.unique_cc_frames                                # (the property guards the lazy host view: read it once)
        ages, gids, first, stop = {}, [], [], []
        for g, grp in enumerate(cc_groups):
            if not grp:
                continue
            marks = set()
            for u in grp:
                fr = uframes[u]
                marks.add(fr[0][0]); marks.add(fr[-1][0])
            marks = sorted(marks)
            ages[g] = marks
            gids.append(g); first.append(marks[0]); stop.append(min(marks[-1] + 1, n_frames))
        # per_frame[t] = groups alive at t, ascending (the reference appends g to every frame of its span, group by group)
        per_frame = [[] for _ in range(n_frames)]
        if gids:
            first, span = np.array(first, dtype=np.int64), np.maximum(np.array(stop, dtype=np.int64) - np.array(first, dtype=np.int64), 0)
            total = int(span.sum())
            if total:
                g_rep = np.repeat(np.array(gids, dtype=np.int64), span)
                t_rep = np.repeat(first, span) + (np.arange(total) - np.repeat(np.cumsum(span) - span, span))
                order = np.argsort(t_rep, kind="stable")               # stable: ascending g inside every frame
                cuts = np.cumsum(np.bincount(t_rep, minlength=n_frames))[:-1]
                per_frame = [part.tolist() for part in np.split(g_rep[order], cuts)]
        return ages, per_frame

    # ---- :446-500 ---------------------------------------------------------------------------------------------
    @gc_paused
    def compute_conflicting_groups(self, stable_idxs, all_overlapping_cc, n_groups, group_idx_per_cc):
        """Same dictionary as the reference's loop (:446-500); the per-pair arithmetic (box areas, intersection, unmatched pixels) runs
        on arrays, only the dictionary is filled pair by pair, in the reference's order."""
        conflicts = {g: {} for g in range(n_groups)}
        pairs = [(u1, u2, m, s2, s1) for u1 in stable_idxs for (u2, m, s2, s1) in all_overlapping_cc[u1] if u1 < u2]
        if not pairs:
            return conflicts
        p = np.array(pairs, dtype=np.int64).reshape(-1, 5)
        n = len(self.unique_cc_objects)
        gi = np.full(n, -1, dtype=np.int64)
        gi[np.fromiter(group_idx_per_cc.keys(), dtype=np.int64, count=len(group_idx_per_cc))] = \
            np.fromiter(group_idx_per_cc.values(), dtype=np.int64, count=len(group_idx_per_cc))
        g1, g2 = gi[p[:, 0]], gi[p[:, 1]]
        keep = g1 != g2
        p, g1, g2 = p[keep], g1[keep], g2[keep]
        box = self._box_table()                                              # int64 [n][4] = min_x, max_x, min_y, max_y
        a, b = box[p[:, 0]], box[p[:, 1]]
        w = np.minimum(a[:, 1], b[:, 1]) - np.maximum(a[:, 0], b[:, 0]) + 1  # connected_component.py getOverlapArea
        h = np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 2], b[:, 2]) + 1
        inter = np.where((w > 0) & (h > 0), w * h, 0)
        area = lambda q: (q[:, 1] - q[:, 0] + 1) * (q[:, 3] - q[:, 2] + 1)
        vals = np.stack([p[:, 2], p[:, 4] + p[:, 3] - 2 * p[:, 2], area(a) + area(b) - inter, inter], axis=1)
        keys = ("matched", "unmatched", "area_union", "area_intersection")
        for ga, gb, v in zip(g1.tolist(), g2.tolist(), vals.tolist()):
            for x, y in ((ga, gb), (gb, ga)):
                slot = conflicts[x].get(y)
                if slot is None:
                    conflicts[x][y] = dict(zip(keys, v))
                else:
                    for k, q in zip(keys, v):
                        slot[k] += q
        return conflicts

    def _box_table(self):
        """int64 [n_uniques][4] = min_x, max_x, min_y, max_y (aliases appended by split_stable_cc_by_gaps included)."""
        objs = self.unique_cc_objects
        cached = getattr(self, "_box_cache", None)
        if cached is not None and len(cached) == len(objs):
            return cached
        dev_boxes = getattr(self, "_uboxes", None)
        if dev_boxes is not None and len(dev_boxes) and getattr(self, "_est", None) is not None:
            _, obj_index = self._unique_view()                            # live estimator: the boxes as the device reported them
            t = dev_boxes[obj_index]
        else:
            t = np.array([(c.min_x, c.max_x, c.min_y, c.max_y) for c in objs], dtype=np.int64).reshape(-1, 4)
        self._box_cache = t
        return t

    # ---- :575-636 ---------------------------------------------------------------------------------------------
    @gc_paused
    def compute_group_images(self, cc_groups, group_ages, segment_threshold):
        lib = _lib.lib()
        view, obj_index = self._unique_view()
        bounds, seg_rows, members, seg_key = {}, [], [], []
        objs, uframes, rows_of = self.unique_cc_objects, self.unique_cc_frames, obj_index.tolist()    # (lazy-view properties: read once)
        b_left, b_right = bisect.bisect_left, bisect.bisect_right
        for g, grp in enumerate(cc_groups):
            if not grp:
                continue
            c = objs[grp[0]]
            x0, x1, y0, y1 = c.min_x, c.max_x, c.min_y, c.max_y
            for u in grp[1:]:
                c = objs[u]
                if c.min_x < x0: x0 = c.min_x
                if c.max_x > x1: x1 = c.max_x
                if c.min_y < y0: y0 = c.min_y
                if c.max_y > y1: y1 = c.max_y
            bounds[g] = (x0, x1, y0, y1)
            box = (int(x0), int(x1), int(y0), int(y1))
            marks = group_ages[g]
            if len(grp) == 1 and len(marks) == 2:                            # one CC, one segment that spans all of its frames
                fr = uframes[grp[0]]
                if marks[0] <= fr[0][0] and fr[-1][0] <= marks[1]:
                    members.append((rows_of[grp[0]], len(fr)))
                    seg_rows.append(box + (len(members) - 1, len(members)))
                    seg_key.append((g, 0))
                    continue
            times = [[t for t, _ in uframes[u]] for u in grp]                # ascending frame indices
            for s, (t0, t1) in enumerate(zip(marks[:-1], marks[1:])):
                begin = len(members)
                for u, ts in zip(grp, times):
                    seen = b_right(ts, t1) - b_left(ts, t0)                  # frames of u inside [t0, t1] (:607)
                    if seen:
                        members.append((rows_of[u], seen))
                seg_rows.append(box + (begin, len(members)))
                seg_key.append((g, s))
        images = {g: [] for g in bounds}
        self._group_device = None
        if not seg_rows:
            return images, bounds
        seg = np.array(seg_rows, dtype=np.int32).reshape(-1, 6)
        words = ((seg[:, 1] >> 5) - (seg[:, 0] >> 5) + 1).astype(np.int64) * (seg[:, 3] - seg[:, 2] + 1)
        offs = np.concatenate([[0], np.cumsum(words)]).astype(np.uint64)
        d_seg, d_mem, d_off = _dev(seg, np.int32), _dev(np.array(members, dtype=np.int32).reshape(-1, 2), np.int32), _dev(offs[:-1], np.uint64)
        d_out = torch.empty(int(offs[-1]) + 1, dtype=torch.int32, device="cuda")
        _lib.check(self._timed("am_group_images", lib.am_group_images, ctypes.byref(view), len(seg), d_seg.data_ptr(), d_mem.data_ptr(),
                               float(segment_threshold), d_off.data_ptr(), d_out.data_ptr(), _stream()), "am_group_images")
        # one unpack for all segments (format conversion only: bit-packed -> the reference's uint8 0/255 arrays)
        px = np.unpackbits(d_out[:int(offs[-1])].cpu().numpy().view(np.uint8), bitorder="little")
        px *= np.uint8(255)
        cws = (((seg[:, 1] >> 5) - (seg[:, 0] >> 5) + 1) * 32).tolist()
        lead, width = (seg[:, 0] & 31).tolist(), (seg[:, 1] - seg[:, 0] + 1).tolist()
        bit0 = (offs.astype(np.int64) * 32).tolist()
        contiguous = np.ascontiguousarray
        for i, (g, s) in enumerate(seg_key):
            x0 = lead[i]
            images[g].append(contiguous(px[bit0[i]:bit0[i + 1]].reshape(-1, cws[i])[:, x0:x0 + width[i]]))
        # kept on the device for frames_from_groups: segment boxes, word offsets, bit-packed images
        self._group_device = ({k: i for i, k in enumerate(seg_key)}, _dev(seg[:, :4], np.int32), d_off, d_out)
        return images, bounds

    # ---- :638-681 ---------------------------------------------------------------------------------------------
    @gc_paused
    def frames_from_groups(self, cc_groups, group_boundaries, groups_per_frame, group_ages, group_images, save_prefix=None,
                           stable_min_frames=3, show_unstable=True, chunk=64):
        """-> list of PNG-encoded clean binary frames (channel 0 of the reference's canvas).  The stable groups are painted on
        the device from the images compute_group_images left there; `save_prefix` debugging dumps are not supported."""
        if save_prefix is not None:
            raise NotImplementedError("frames_from_groups(save_prefix=...) writes debugging PNGs; only the returned frames are produced here")
        lib = _lib.lib()
        dev = getattr(self, "_group_device", None)
        if dev is None:
            dev = self._upload_group_images(group_images, group_boundaries)
        seg_index, d_boxes, d_off, d_imgs = dev
        items = frame_segment_items(groups_per_frame, group_ages, seg_index, len(cc_groups))
        clean, n_frames = [], len(groups_per_frame)
        from .wire import PngEncoder
        chunk = min(chunk, max(n_frames, 1))
        enc = PngEncoder(self.width, self.height, chunk, torch.device("cuda", torch.cuda.current_device()), compress=True)
        wpr = lib.am_words_per_row(self.width)
        d_bits = torch.empty((chunk, self.height, wpr), dtype=torch.int32, device="cuda")
        d_flags = torch.zeros((chunk,), dtype=torch.int32, device="cuda")
        out = torch.empty(chunk * self.height * self.width + 8, dtype=torch.uint8, device="cuda")
        enc8 = None
        for f0 in range(0, n_frames, chunk):
            nf = min(chunk, n_frames - f0)
            sel = items[(items[:, 0] >= f0) & (items[:, 0] < f0 + nf)]
            if len(sel):
                d_f, d_i = _dev(sel[:, 0], np.int32), _dev(sel[:, 1], np.int32)
                args = (len(sel), d_f.data_ptr(), d_i.data_ptr(), d_boxes.data_ptr(), d_off.data_ptr(), d_imgs.data_ptr())
            else:
                args = (0, None, None, None, None, None)
            _lib.check(self._timed("am_paint_frames", lib.am_paint_frames, *args, f0, nf, self.height, self.width, out.data_ptr(), _stream()),
                       "am_paint_frames")
            # the 03 -> 04 wire format (`cv2.imencode(".png", reconstructed[:, :, 0])`, :677-678) written on the device: a clean frame
            # whose pixels are all 0 / 255 is exactly a 1-bit image (csrc/png.cu, decodes to the same array); a frame where two groups
            # overlap wraps to 254 (uint8 `+=`) and is written as an 8-bit grayscale PNG by the same writer
            _lib.check(lib.am_pack_mask_u8_exact(out.data_ptr(), self.width, self.height, nf, d_bits.data_ptr(), d_flags.data_ptr(), _stream()),
                       "am_pack_mask_u8_exact")
            files = enc.encode(d_bits, nf)
            flags = d_flags[:nf].cpu().numpy()
            for t in np.nonzero(flags)[0].tolist():                       # the 8-bit form of the same writer keeps the 254s
                if enc8 is None:
                    enc8 = PngEncoder(self.width, self.height, 1, enc.device, compress=True, depth=8)
                files[t] = enc8.encode(out[t * self.height * self.width:(t + 1) * self.height * self.width].view(1, self.height, self.width), 1)[0]
            clean.extend(files)
        return clean

    def _upload_group_images(self, group_images, group_boundaries):
        """Group images supplied by the caller (e.g. after unpickling): pack and upload them."""
        keys, boxes, chunks = [], [], []
        for g in sorted(group_images):
            x0, x1, y0, y1 = (int(v) for v in group_boundaries[g])
            for s, im in enumerate(group_images[g]):
                keys.append((g, s)); boxes.append((x0, x1, y0, y1)); chunks.append(pack_crop(im, x0, x1, y0, y1))
        words = np.array([len(c) for c in chunks], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(words)]).astype(np.uint64)
        imgs = np.concatenate(chunks) if chunks else np.zeros(1, np.uint32)
        self._group_device = ({k: i for i, k in enumerate(keys)}, _dev(np.array(boxes, dtype=np.int32).reshape(-1, 4), np.int32),
                              _dev(offs[:-1], np.uint64), _dev(imgs.view(np.int32), np.int32))
        return self._group_device
