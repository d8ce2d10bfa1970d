"""CCStabilityEstimator drop-in (R/AccessMath/preprocessing/content/cc_stability_estimator.py:9-164), hot-path
part: __init__, add_frame(img, input_binary=True), finish_processing, get_raw_cc_count -- labeling, statistics
and temporal matching all run on the B200 (libaccessmath_b200.so); this class only keeps the Python-visible state
the reference exposes (unique_cc_objects, unique_cc_frames, cc_idx_per_frame, img_idx, tempo_count ...).

`add_frames(masks)` is the batched fast path (many frames per launch sequence, one read-back per batch).
The stage-03 methods (rebuilt_binary_images ... frames_from_groups, :166-681) come from cc_grouping.GroupingMixin."""
import numpy as np
import torch

from .cc_engine import CCEngine, Estimator
from .cc_grouping import GroupingMixin
from .connected_component import ConnectedComponent


class CCStabilityEstimator(GroupingMixin):
    def __init__(self, width, height, min_recall, min_precision, max_gap, verbose=False, max_batch=16):
        self.width, self.height = width, height
        self.min_recall, self.min_precision, self.max_gap = min_recall, min_precision, max_gap
        self.unique_cc_objects = []
        self.unique_cc_frames = []
        self.cc_idx_per_frame = []
        self.fake_age = None
        self.img_idx = 0
        self.tempo_count = 0
        self.verbose = verbose
        self._engine = CCEngine(width, height, max_batch)
        self._est = Estimator(width, height, min_recall, min_precision, max_gap)
        self._frame_tables = []          # per frame: (boxes int32 [n][4], crop word offsets, the frame's packed crops) as the GPU returned them

    def __getstate__(self):
        """Pickled like the reference object (tempo_stability_*.dat, pre_ST3D_v3.0_02_cc_analaysis.py:43): the Python-visible
        state only; device handles stay behind and stage 03 re-uploads the packed crops when it runs in another process."""
        drop = ("_engine", "_est", "_view_cache", "_group_device", "device_ms", "_frame_tables")
        return {k: v for k, v in self.__dict__.items() if k not in drop}

    def get_raw_cc_count(self):                                          # :33-39
        return sum(len(f) for f in self.cc_idx_per_frame)

    # ---- reference-compatible per-frame entry point -------------------------------------------------
    def add_frame(self, img, input_binary=False):
        if not input_binary:
            raise NotImplementedError("add_frame(input_binary=False) uses the legacy background-subtraction binarizer "
                                      "(cc_stability_estimator.py:48), which is outside the FCN hot path")
        self.add_frames(np.asarray(img)[None])

    def add_frames(self, masks):
        """masks: uint8 (n, H, W) numpy array or CUDA tensor, ink != 0."""
        if isinstance(masks, np.ndarray):
            masks = torch.from_numpy(np.ascontiguousarray(masks != 0).view(np.uint8)).cuda(non_blocking=True)
        eng = self._engine
        for s in range(0, masks.shape[0], eng.max_batch):
            chunk = masks[s:s + eng.max_batch].contiguous()
            self.add_packed(eng.pack(chunk))

    def add_packed(self, bits):
        """bits: bit-packed CUDA tensor (n <= max_batch, H, WPR) -- what the FCN epilogue emits."""
        eng, est = self._engine, self._est
        n = bits.shape[0]
        eng.label(bits, want_labels=False, sync=False)
        est.add_frames(eng, 0, n)
        eng.read_counts()
        rows, offs = eng.packed_rows(n)
        rows_h, offs_h = rows.cpu().numpy(), offs.cpu().numpy()
        st = est.state()                                                 # raises on a capacity / hand-off flag BEFORE rows are absorbed
        crops = [eng.crops(f) if eng.counts[f, 2] else None for f in range(n)]
        for f in range(n):
            self._absorb(rows_h[offs_h[f]:offs_h[f + 1]], crops[f])
        self.tempo_count = st["tempo_count"]
        if st["img_idx"] != self.img_idx or st["n_unique"] != len(self.unique_cc_objects):
            raise RuntimeError("device estimator state (img_idx %d, %d uniques) does not match the host view (%d, %d)"
                               % (st["img_idx"], st["n_unique"], self.img_idx, len(self.unique_cc_objects)))

    def _absorb(self, rows, crops):
        current = []
        self._frame_tables.append((np.ascontiguousarray(rows[:, 2:6]), rows[:, 7].astype(np.uint64),
                                   crops if crops is not None else np.zeros(1, np.uint32)))
        for r in rows:
            u, lab, x0, x1, y0, y1, size, off = (int(v) for v in r)
            words = ((x1 >> 5) - (x0 >> 5) + 1) * (y1 - y0 + 1)
            cc = ConnectedComponent(lab - 1, np.int32(x0), np.int32(x1), np.int32(y0), np.int32(y1), np.int32(size),
                                    packed=crops[off:off + words])
            cc.start_time = cc.end_time = np.float32(0.0)
            if u == len(self.unique_cc_objects):                         # new unique CC (:111-124)
                self.unique_cc_objects.append(cc)
                self.unique_cc_frames.append([(self.img_idx, lab)])
            else:                                                        # matched (:99-104)
                self.unique_cc_frames[u].append((self.img_idx, lab))
            current.append((u, cc))
        self.cc_idx_per_frame.append(current)
        self.img_idx += 1
        if self.verbose:
            print("[" + str(self.img_idx) + " (" + str(len(rows)) + ", " + str(len(self.unique_cc_objects)) + ")]", end="\r")

    def finish_processing(self):                                         # :158-164
        if self.verbose:
            print(".")
        print("Total CC merges tested: " + str(self.tempo_count))
        self.fake_age = None
