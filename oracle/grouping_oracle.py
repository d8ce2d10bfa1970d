"""oracle/grouping_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for stage 03).

Restates, on top of a finished StabilityOracle (oracle/cc_oracle.py), the estimator methods that
R/pre_ST3D_v3.0_03_cc_grouping.py:41-101 calls (R/AccessMath/preprocessing/content/cc_stability_estimator.py):
  rebuilt_binary_images / rebuilt_binary_frame   <- :166-179
  split_stable_cc_by_gaps                        <- :181-228
  get_stable_cc_idxs                             <- :230-236
  compute_overlapping_stable_cc                  <- :245-306  (IntervalIndex sweeps = inclusive bbox overlap, interval_index.py:42-99;
                                                               getOverlapFMeasure, connected_component.py:202-250)
  compute_groups                                 <- :308-413
  compute_groups_temporal_information            <- :415-444
  compute_conflicting_groups                     <- :446-500  (getBoxArea / getOverlapArea, connected_component.py:46-68)
  compute_group_images                           <- :575-636
  frames_from_groups                             <- :638-681
Method names, argument order and result shapes are the reference's, so oracle/gen_golden_grouping.run_stage03 drives the
reference class, this oracle and the CUDA drop-in alike.

Pinned by tests/golden/cc_grouping.npz (outputs of the unmodified reference on four seeded videos, oracle/gen_golden_grouping.py).
Only tests/, smoke() and bench.py's reference legs may import this module.
"""
import numpy as np

from .cc_oracle import overlap_measure


def _box_area(cc):
    return (cc.max_x - cc.min_x + 1) * (cc.max_y - cc.min_y + 1)


def _overlap_area(a, b):
    if a.min_x <= b.max_x and b.min_x <= a.max_x and a.min_y <= b.max_y and b.min_y <= a.max_y:
        return (min(a.max_x, b.max_x) - max(a.min_x, b.min_x) + 1) * (min(a.max_y, b.max_y) - max(a.min_y, b.min_y) + 1)
    return 0


class GroupingOracle:
    def __init__(self, stab):
        self.width, self.height = stab.width, stab.height
        self.unique_cc_objects = list(stab.unique_cc_objects)
        self.unique_cc_frames = [list(f) for f in stab.unique_cc_frames]
        self.cc_idx_per_frame = [list(f) for f in stab.cc_idx_per_frame]

    # ---- :166-179 ---------------------------------------------------------------------------------------------
    def rebuilt_binary_frame(self, frame_ccs):
        binary = np.zeros((self.height, self.width), dtype=np.uint8)
        for _, cc in frame_ccs:
            binary[cc.min_y:cc.max_y + 1, cc.min_x:cc.max_x + 1] += cc.img
        return binary

    def rebuilt_binary_images(self):
        return [self.rebuilt_binary_frame(f) for f in self.cc_idx_per_frame]

    # ---- :181-228 ---------------------------------------------------------------------------------------------
    def split_stable_cc_by_gaps(self, max_gap, stable_min_frames):
        split = 0
        for u in range(len(self.unique_cc_objects)):                  # only the uniques that exist before splitting
            frames = self.unique_cc_frames[u]
            cuts = [i for i in range(1, len(frames)) if frames[i][0] - frames[i - 1][0] > max_gap]
            if not cuts or len(frames) < stable_min_frames:
                continue
            edges = [0] + cuts + [len(frames)]
            runs = [frames[a:b] for a, b in zip(edges[:-1], edges[1:])]
            self.unique_cc_frames[u] = runs[0]
            for run in runs[1:]:
                new_u = len(self.unique_cc_objects)
                self.unique_cc_objects.append(self.unique_cc_objects[u])      # same object, new index
                self.unique_cc_frames.append(run)
                for t, _ in run:
                    row = self.cc_idx_per_frame[t]
                    for k, (uk, cc) in enumerate(row):                # the FIRST instance still carrying the old index
                        if uk == u:
                            row[k] = (new_u, cc)
                            break
            split += 1
        return split

    # ---- :230-236 ---------------------------------------------------------------------------------------------
    def get_stable_cc_idxs(self, min_stable_frames):
        return [u for u, f in enumerate(self.unique_cc_frames) if len(f) >= min_stable_frames]

    # ---- :245-306 ---------------------------------------------------------------------------------------------
    def compute_overlapping_stable_cc(self, stable_idxs, temporal_window):
        n = len(self.unique_cc_objects)
        all_ov, time_ov, total = [[] for _ in range(n)], [[] for _ in range(n)], 0
        ids = np.asarray(stable_idxs, dtype=np.int64)
        box = np.array([(self.unique_cc_objects[u].min_x, self.unique_cc_objects[u].max_x, self.unique_cc_objects[u].min_y,
                         self.unique_cc_objects[u].max_y) for u in stable_idxs], dtype=np.int64).reshape(-1, 4)
        for a in range(len(ids)):                                     # ascending (idx1, idx2), idx1 < idx2
            later = slice(a + 1, None)
            hit = ((box[a, 0] <= box[later, 1]) & (box[later, 0] <= box[a, 1]) &
                   (box[a, 2] <= box[later, 3]) & (box[later, 2] <= box[a, 3]))
            u1 = int(ids[a])
            cc1, f1 = self.unique_cc_objects[u1], self.unique_cc_frames[u1]
            for u2 in ids[later][hit].tolist():
                cc2, f2 = self.unique_cc_objects[u2], self.unique_cc_frames[u2]
                recall, precision = overlap_measure(cc1, cc2)
                if recall > 0.0 or precision > 0.0:
                    matched = int(cc1.size * recall)                  # :284 (fp64 product truncated, as the reference)
                    all_ov[u1].append((u2, matched, cc2.size, cc1.size))
                    all_ov[u2].append((u1, matched, cc1.size, cc2.size))
                    if f1[-1][0] + temporal_window >= f2[0][0] and f2[-1][0] >= f1[0][0] - temporal_window:
                        time_ov[u1].append((u2, recall, precision))
                        time_ov[u2].append((u1, precision, recall))
                        total += 1
        return time_ov, total, all_ov

    # ---- :308-413 ---------------------------------------------------------------------------------------------
    def compute_groups(self, stable_idxs, overlapping_cc, min_recall, t_fmeasure, t_time_IOU):
        groups, owner = [], {}
        for u1 in stable_idxs:
            if u1 not in owner:
                owner[u1] = len(groups)
                groups.append([u1])
            g = owner[u1]
            for u2, recall, _ in overlapping_cc[u1]:
                if recall < min_recall:
                    continue
                if u2 not in owner:
                    owner[u2] = g
                    groups[g].append(u2)
                elif owner[u2] != g:                                  # absorb the other group, keep its slot empty
                    other = owner[u2]
                    for m in groups[other]:
                        owner[m] = g
                        groups[g].append(m)
                    groups[other] = []
        final, final_owner = [], {}
        for grp in groups:
            if grp:
                for m in grp:
                    final_owner[m] = len(final)
                final.append(grp)
        return final, final_owner

    # ---- :415-444 ---------------------------------------------------------------------------------------------
    def compute_groups_temporal_information(self, cc_groups):
        n_frames = len(self.cc_idx_per_frame)
        ages, per_frame = {}, [[] for _ in range(n_frames)]
        for g, grp in enumerate(cc_groups):
            if not grp:
                continue
            marks = sorted({self.unique_cc_frames[u][0][0] for u in grp} | {self.unique_cc_frames[u][-1][0] for u in grp})
            ages[g] = marks
            for t in range(marks[0], min(marks[-1] + 1, n_frames)):
                per_frame[t].append(g)
        return ages, per_frame

    # ---- :446-500 ---------------------------------------------------------------------------------------------
    def compute_conflicting_groups(self, stable_idxs, all_overlapping_cc, n_groups, group_idx_per_cc):
        conflicts = {g: {} for g in range(n_groups)}
        keys = ("matched", "unmatched", "area_union", "area_intersection")
        for u1 in stable_idxs:
            cc1 = self.unique_cc_objects[u1]
            for u2, matched, size2, size1 in all_overlapping_cc[u1]:
                if not u1 < u2:
                    continue
                g1, g2 = group_idx_per_cc[u1], group_idx_per_cc[u2]
                if g1 == g2:
                    continue
                cc2 = self.unique_cc_objects[u2]
                inter = _overlap_area(cc1, cc2)
                vals = (matched, size1 + size2 - matched * 2, _box_area(cc1) + _box_area(cc2) - inter, inter)
                for a, b in ((g1, g2), (g2, g1)):
                    slot = conflicts[a].setdefault(b, dict.fromkeys(keys, 0))
                    for k, v in zip(keys, vals):
                        slot[k] += v
        return conflicts

    # ---- :575-636 ---------------------------------------------------------------------------------------------
    def compute_group_images(self, cc_groups, group_ages, segment_threshold):
        images, bounds = {}, {}
        for g, grp in enumerate(cc_groups):
            if not grp:
                continue
            ccs = [self.unique_cc_objects[u] for u in grp]
            x0, x1 = min(c.min_x for c in ccs), max(c.max_x for c in ccs)
            y0, y1 = min(c.min_y for c in ccs), max(c.max_y for c in ccs)
            bounds[g] = (x0, x1, y0, y1)
            marks, segs = group_ages[g], []
            for t0, t1 in zip(marks[:-1], marks[1:]):
                votes = np.zeros((y1 - y0 + 1, x1 - x0 + 1), dtype=np.int32)
                for u, cc in zip(grp, ccs):
                    seen = sum(1 for t, _ in self.unique_cc_frames[u] if t0 <= t <= t1)
                    if seen:
                        votes[cc.min_y - y0:cc.max_y - y0 + 1, cc.min_x - x0:cc.max_x - x0 + 1] += (cc.img // 255) * seen
                with np.errstate(all="ignore"):
                    segs.append(((votes.astype(np.float64) / votes.max()) >= segment_threshold).astype(np.uint8) * 255)
            images[g] = segs
        return images, bounds

    # ---- :638-681 ---------------------------------------------------------------------------------------------
    def frames_from_groups(self, cc_groups, group_boundaries, groups_per_frame, group_ages, group_images, save_prefix=None,
                           stable_min_frames=3, show_unstable=True):
        import cv2
        seg_of = [0] * len(cc_groups)
        clean = []
        for t, present in enumerate(groups_per_frame):
            frame = np.zeros((self.height, self.width), dtype=np.uint8)        # channel 0 of the reference's BGR canvas
            for g in present:
                marks = group_ages[g]
                while marks[seg_of[g] + 1] < t:
                    seg_of[g] += 1
                x0, x1, y0, y1 = group_boundaries[g]
                frame[y0:y1 + 1, x0:x1 + 1] += group_images[g][seg_of[g]]      # uint8 wrap: two groups -> 254
            clean.append(cv2.imencode(".png", frame)[1])
        return clean
