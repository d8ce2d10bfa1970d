"""KeyframeExtractor.GenerateFromST3DForIntervals (R/AccessMath/preprocessing/content/keyframe_extractor.py:12-150) -- the overlap
tests of the key-frame builder on the device.

For every video segment the reference wraps the image of each CC group alive in it in a ConnectedComponent and calls
getOverlapFMeasure on ALL pairs (CCStabilityEstimator.compute_overlapping_CC_groups, cc_stability_estimator.py:696-749: a Python
double loop, each test slicing two arrays), once more inside every conflict, then paints the winners.  Here the pixel work of a
segment is two launches: am_group_overlaps over the bit-packed group images (all pairs with intersecting boxes -> matched pixels;
`recall > 0 or precision > 0` is `matched > 0`) and am_paint_frames for the key-frame.  What stays on the host is the bookkeeping
whose iteration ORDER is part of the reference's result (set / dict based transitive merge, most-recent-first acceptance inside a
conflict), written with the same containers.  `st3D` is duck-typed: frame_times, height, width, cc_group_ages, cc_group_images,
cc_group_boundaries (R/AccessMath/data/space_time_struct.py:5-16)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .cc_grouping import _dev, _stream, overlap_pairs, view_from_components
from .connected_component import ConnectedComponent


class KeyframeExtractor:
    @staticmethod
    def compute_overlapping_CC_groups(cc_objects, return_matches=False):
        """(overlapping_groups, no_overlaps) as CCStabilityEstimator.compute_overlapping_CC_groups (:696-749); with return_matches also
        the set of overlapping position pairs (a < b)."""
        n = len(cc_objects)
        adjacency = [[i] for i in range(n)]
        matches = set()
        if n >= 2:
            view, keep = view_from_components(cc_objects)
            pairs = overlap_pairs(view, _dev(np.arange(n), np.int32), n)
            for a, b, m in pairs.tolist():                            # ascending (a, b): the reference's append order (:703-714)
                if m > 0:
                    adjacency[a].append(b)
                    adjacency[b].append(a)
                    matches.add((a, b))
            del keep
        owner = list(range(n))                                        # :717-736
        merged = {i: {i} for i in range(n)}
        for i in range(n):
            g1 = owner[i]
            for j in adjacency[i][1:]:
                g2 = owner[j]
                if g1 != g2:
                    merged[g1] = merged[g1].union(merged[g2])
                    for k in merged[g2]:
                        owner[k] = g1
                    del merged[g2]
        groups, singles = [], []
        for g in merged:                                              # :739-746
            members = list(merged[g])
            (singles if len(members) == 1 else groups).append(members[0] if len(members) == 1 else members)
        return (groups, singles, matches) if return_matches else (groups, singles)

    @staticmethod
    def GenerateFromST3DForIntervals(st3D, video_segments, verbose=True):
        lib = _lib.lib()
        ages_of, images_of, bounds_of = st3D.cc_group_ages, st3D.cc_group_images, st3D.cc_group_boundaries
        H, W = st3D.height, st3D.width
        final_keyframes, keyframes_times = [], []
        if verbose:
            print("Total CC Groups Given: " + str(len(bounds_of)))
            print("Total Video Segments: " + str(len(video_segments)))
        for segment_idx, (start_int, end_int) in enumerate(video_segments):
            if verbose:
                print("Processing segment #{0:d} ({1:d} - {2:d})".format(segment_idx + 1, start_int, end_int))
            ccs = []
            for g in ages_of:                                         # :29-48
                ages = ages_of[g]
                if start_int <= ages[-1] and ages[0] <= end_int:
                    last = 0
                    while last + 2 < len(ages) and ages[last + 2] <= end_int:
                        last += 1
                    x0, x1, y0, y1 = bounds_of[g]
                    img = images_of[g][last]
                    ccs.append(ConnectedComponent(g, x0, x1, y0, y1, img.sum() // 255, img))
            groups, singles, matches = KeyframeExtractor.compute_overlapping_CC_groups(ccs, True)
            chosen = list(singles)                                    # positions painted into the key-frame, in the reference's order
            in_conflict = 0
            for group in groups:                                      # :67-131
                in_conflict += len(group)
                by_age = sorted(((ages_of[ccs[pos].cc_id][0], i) for i, pos in enumerate(group)), reverse=True)
                accepted = []
                for _, i in by_age:                                   # incompatible <=> the two images share a pixel (recall > 0, :85-88)
                    a = group[i]
                    if not any((min(a, group[k]), max(a, group[k])) in matches for k in accepted):
                        accepted.append(i)
                chosen.extend(group[i] for i in accepted)
            local_times = [(st3D.frame_times[ages_of[ccs[pos].cc_id][0]], ccs[pos].min_x, ccs[pos].max_x, ccs[pos].min_y, ccs[pos].max_y)
                           for pos in chosen]
            final_keyframes.append(KeyframeExtractor._paint(lib, [ccs[pos] for pos in chosen], H, W))
            if verbose:
                print("-> Total Groups contained: " + str(len(ccs)))
                print("-> Total Groups without Conflicts: " + str(len(singles)))
                print("-> Total Groups with Conflicts: " + str(in_conflict))
            keyframes_times.append(sorted(local_times))
        return final_keyframes, keyframes_times

    @staticmethod
    def _paint(lib, ccs, H, W):
        """255 - (white where any chosen group has a pixel), three equal channels (:133-146) -- am_paint_frames + one comparison."""
        out = torch.empty(H * W + 8, dtype=torch.uint8, device="cuda")
        if ccs:
            view, keep = view_from_components(ccs)
            n = len(ccs)
            boxes = _dev(np.array([(int(c.min_x), int(c.max_x), int(c.min_y), int(c.max_y)) for c in ccs], dtype=np.int32).reshape(-1, 4), np.int32)
            d_frame, d_img = _dev(np.zeros(n), np.int32), _dev(np.arange(n), np.int32)
            args = (n, d_frame.data_ptr(), d_img.data_ptr(), boxes.data_ptr(), ctypes.c_void_p(view.crop_off), ctypes.c_void_p(view.arena))
        else:
            args = (0, None, None, None, None, None)
        _lib.check(lib.am_paint_frames(*args, 0, 1, H, W, out.data_ptr(), _stream()), "am_paint_frames")
        gray = torch.where(out[:H * W] != 0, 0, 255).to(torch.uint8).view(H, W)          # `+= 255` per group: non-zero wherever one painted
        return gray.unsqueeze(-1).expand(H, W, 3).contiguous().cpu().numpy()
