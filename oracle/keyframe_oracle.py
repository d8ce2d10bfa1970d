"""oracle/keyframe_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for the key-frame overlap tests).

CPU restatement of (R/ = /root/reference/ACCESS2021_release/):
  * overlapping_cc_groups      <- CCStabilityEstimator.compute_overlapping_CC_groups,
                                  R/AccessMath/preprocessing/content/cc_stability_estimator.py:696-749
  * keyframes_for_intervals    <- KeyframeExtractor.GenerateFromST3DForIntervals,
                                  R/AccessMath/preprocessing/content/keyframe_extractor.py:12-150
over the SpaceTimeStruct fields (R/AccessMath/data/space_time_struct.py:5-16).  The order of Python's set / dict iteration is part of
the reference's result (which member of a conflict wins a tie depends on its position in `list(set)`), so the same containers are used.

Pinned by tests/golden/keyframes.npz (outputs of the unmodified reference, oracle/gen_golden_keyframes.py).
Only tests/ may import this module."""
import numpy as np

from oracle.cc_oracle import overlap_measure


class GroupCC:
    def __init__(self, cc_id, min_x, max_x, min_y, max_y, size, img):
        self.cc_id, self.min_x, self.max_x, self.min_y, self.max_y, self.size, self.img = cc_id, min_x, max_x, min_y, max_y, size, img


def overlapping_cc_groups(ccs):
    """-> (overlapping_groups: list of lists of positions, no_overlaps: list of positions)  (:696-749)."""
    n = len(ccs)
    adj = [[i] for i in range(n)]
    for a in range(n):                                            # :703-714
        for b in range(a + 1, n):
            with np.errstate(all="ignore"):
                recall, precision = overlap_measure(ccs[a], ccs[b])
            if recall > 0.0 or precision > 0.0:
                adj[a].append(b)
                adj[b].append(a)
    owner = list(range(n))                                        # :717-736 (transitive merge; sets as in the reference)
    merged = {i: {i} for i in range(n)}
    for i in range(n):
        g1 = owner[i]
        for j in adj[i][1:]:
            g2 = owner[j]
            if g1 != g2:
                merged[g1] = merged[g1].union(merged[g2])
                for k in merged[g2]:
                    owner[k] = g1
                del merged[g2]
    groups, singles = [], []                                      # :739-746
    for g in merged:
        members = list(merged[g])
        if len(members) == 1:
            singles.append(members[0])
        else:
            groups.append(members)
    return groups, singles


def keyframes_for_intervals(frame_times, height, width, group_ages, group_images, group_boundaries, video_segments):
    """-> (keyframes: list of uint8 (H, W, 3), keyframe_times: list of sorted [(start_time, min_x, max_x, min_y, max_y)])  (:12-150)."""
    keyframes, times = [], []
    for start, end in video_segments:
        local, as_cc = [], {}
        for g in group_ages:                                      # :29-48: groups alive in the segment, image of the last overlapping interval
            ages = group_ages[g]
            if start <= ages[-1] and ages[0] <= end:
                last = 0
                while last + 2 < len(ages) and ages[last + 2] <= end:
                    last += 1
                x0, x1, y0, y1 = group_boundaries[g]
                img = group_images[g][last]
                as_cc[g] = GroupCC(g, x0, x1, y0, y1, img.sum() // 255, img)
        ccs = list(as_cc.values())
        groups, singles = overlapping_cc_groups(ccs)              # :51-52
        mask = np.zeros((height, width), dtype=np.int32)

        def paint(cc):                                            # :60-64 / :126-131
            mask[cc.min_y:cc.max_y + 1, cc.min_x:cc.max_x + 1] += cc.img // 255
            local.append((frame_times[group_ages[cc.cc_id][0]], cc.min_x, cc.max_x, cc.min_y, cc.max_y))

        for pos in singles:
            paint(ccs[pos])
        for group in groups:                                      # :67-131: inside a conflict keep the most recent compatible members
            n = len(group)
            incompatible = np.zeros((n, n), dtype=bool)
            by_age = []
            for i, pos in enumerate(group):
                by_age.append((group_ages[ccs[pos].cc_id][0], i))
                for j in range(i + 1, n):
                    with np.errstate(all="ignore"):
                        recall, _ = overlap_measure(ccs[pos], ccs[group[j]])
                    if recall > 0.0:
                        incompatible[i, j] = incompatible[j, i] = True
            accepted = []
            for _, i in sorted(by_age, reverse=True):
                if not any(incompatible[k, i] for k in accepted):
                    accepted.append(i)
            for i in accepted:
                paint(ccs[group[i]])
        img = np.zeros((height, width, 3), dtype=np.uint8)        # :133-138: conflicts end up white in every channel too
        img[mask >= 1, :] = 255
        keyframes.append(255 - img)                               # :146
        times.append(sorted(local))
    return keyframes, times
