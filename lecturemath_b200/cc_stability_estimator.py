"""CCStabilityEstimator drop-in (R/AccessMath/preprocessing/content/cc_stability_estimator.py:9-164), hot-path
part: __init__, add_frame(img, input_binary=True), finish_processing, get_raw_cc_count -- labeling, statistics
and temporal matching all run on the B200 (libaccessmath_b200.so); this class only keeps the Python-visible state
the reference exposes (unique_cc_objects, unique_cc_frames, cc_idx_per_frame, img_idx, tempo_count ...).

The reference calls add_frame once per frame (R/pre_ST3D_v3.0_02_cc_analaysis.py:37-38) and reads the state after
finish_processing().  Here add_frame only STAGES the frame (pinned host memory; bit-packed when it comes from this package's
Helper.decompress_binary_images); every `max_batch` frames one upload + one label / match launch sequence run on the GPU, and the
Python objects of the reference's state (one ConnectedComponent per CC and frame) are built lazily, the first time
unique_cc_objects / unique_cc_frames / cc_idx_per_frame are read -- results are identical, the per-frame host cost is a memcpy.
`add_frames(masks)` / `add_packed(bits)` are the explicit batched entry points.
The stage-03 methods (rebuilt_binary_images ... frames_from_groups, :166-681) come from cc_grouping.GroupingMixin."""
import ctypes

import numpy as np
import torch

from . import _lib
from .cc_engine import CCEngine, Estimator
from .cc_grouping import GroupingMixin, gc_paused
from .connected_component import ConnectedComponent
from .packed_mask import PackedMask



class _LazyFrames:
    """cc_idx_per_frame of a live estimator: per frame the list [(unique_idx, ConnectedComponent), ...] of the reference
    (cc_stability_estimator.py:57, 102, 115, 145), built from the frame's result rows the first time somebody looks at that frame.
    A long video holds millions of (frame, CC) instances, stage 03 touches a handful of frames (split_stable_cc_by_gaps) -- creating
    them all up front was the single largest cost of the drop-in (1 s per 160k instances)."""

    def __init__(self, est):
        self._est = est
        self._rows = []                  # per frame: the materialised list, or the raw (rows int32 [n][8], packed crops) pair

    def _push_raw(self, rows, crops):
        self._rows.append((rows, crops))

    def _get(self, t):
        r = self._rows[t]
        if isinstance(r, tuple):
            if t < 0:
                t += len(self._rows)
            r = self._rows[t] = self._est._build_row(t, *r)
        return r

    def row_len(self, t):
        r = self._rows[t]
        return len(r[0]) if isinstance(r, tuple) else len(r)

    def unique_indices(self, t):
        r = self._rows[t]
        return r[0][:, 0].tolist() if isinstance(r, tuple) else [u for u, _ in r]

    def __len__(self):
        return len(self._rows)

    def __getitem__(self, t):
        if isinstance(t, slice):
            return [self._get(i) for i in range(*t.indices(len(self._rows)))]
        return self._get(t)

    def __setitem__(self, t, value):
        self._rows[t] = value

    def __iter__(self):
        for t in range(len(self._rows)):
            yield self._get(t)

    def append(self, value):
        self._rows.append(value)

    def __eq__(self, other):
        return list(self) == list(other)

    def __reduce__(self):                # pickles as the plain list it stands for
        return (list, (list(self),))


class CCStabilityEstimator(GroupingMixin):
    def __init__(self, width, height, min_recall, min_precision, max_gap, verbose=False, max_batch=16):
        self.width, self.height = width, height
        self.min_recall, self.min_precision, self.max_gap = min_recall, min_precision, max_gap
        self._uniques, self._uframes, self._per_frame = [], [], _LazyFrames(self)
        self._ufirst = []                # per unique of the device numbering: (frame, raw label) of its first appearance
        self._uboxes = np.zeros((0, 4), dtype=np.int64)
        self.fake_age = None
        self.img_idx = 0                 # frames handed in (staged frames included)
        self._tempo = 0
        self.verbose = verbose
        self._init_device(max_batch)

    def _init_device(self, max_batch):
        self._max_batch = int(max_batch)
        self._engines = [CCEngine(self.width, self.height, self._max_batch) for _ in range(2)]
        self._engine = self._engines[0]
        self._est = Estimator(self.width, self.height, self.min_recall, self.min_precision, self.max_gap)
        self._frame_tables = []          # per frame: (boxes int32 [n][4], crop word offsets, the frame's packed crops) as the GPU returned them
        self._raw = []                   # per frame, not yet turned into Python objects: (rows int32 [n][8], packed crops)
        self._absorbed = 0               # frames whose objects exist
        self._slot = 0
        self._inflight = [None, None]    # frames of the batch each engine holds results for
        self._ev = [torch.cuda.Event() for _ in range(2)]
        self._staged = 0                 # frames waiting in the staging buffers
        self._stage_kind = None          # "scan" | "words" | "dense"
        self._h_stage = {}               # kind -> pinned staging tensor
        self._d_stage = {}
        self._d_bits = None
        self._copy_stream = None
        self._d2h_stream = None

    # ---- the reference's attributes, materialised on demand --------------------------------------------------------
    @property
    def unique_cc_objects(self):
        self._materialise()
        return self._uniques

    @unique_cc_objects.setter
    def unique_cc_objects(self, v):
        self._uniques = v

    @property
    def unique_cc_frames(self):
        self._materialise()
        return self._uframes

    @unique_cc_frames.setter
    def unique_cc_frames(self, v):
        self._uframes = v

    @property
    def cc_idx_per_frame(self):
        self._materialise()
        return self._per_frame

    @cc_idx_per_frame.setter
    def cc_idx_per_frame(self, v):
        self._per_frame = v

    @property
    def tempo_count(self):
        if getattr(self, "_engines", None) is not None and (self._staged or self._inflight != [None, None]):
            self.flush()
        return self._tempo

    @tempo_count.setter
    def tempo_count(self, v):
        self._tempo = v

    def __getstate__(self):
        """Pickled like the reference object (tempo_stability_*.dat, pre_ST3D_v3.0_02_cc_analaysis.py:43): the Python-visible
        state only; device handles stay behind and stage 03 re-uploads the packed crops when it runs in another process.
        (lecturemath_b200.compat.dump_reference_pickle writes it under the reference's own class paths.)"""
        self._materialise()
        keep = {k: v for k, v in self.__dict__.items() if not k.startswith("_") and k != "device_ms"}
        keep.update(unique_cc_objects=self._uniques, unique_cc_frames=self._uframes, cc_idx_per_frame=list(self._per_frame),
                    tempo_count=self._tempo)
        return keep

    def __setstate__(self, state):
        state = dict(state)
        self._uniques = state.pop("unique_cc_objects", [])
        self._uframes = state.pop("unique_cc_frames", [])
        self._per_frame = state.pop("cc_idx_per_frame", [])
        self._tempo = state.pop("tempo_count", 0)
        self.__dict__.update(state)

    def get_raw_cc_count(self):                                          # :33-39
        self.flush()
        pf = self._per_frame
        done = sum(pf.row_len(t) for t in range(len(pf))) if isinstance(pf, _LazyFrames) else sum(len(f) for f in pf)
        return done + sum(len(r[0]) for r in getattr(self, "_raw", []))

    # ---- reference-compatible per-frame entry point -------------------------------------------------
    def add_frame(self, img, input_binary=False):
        if not input_binary:
            raise NotImplementedError("add_frame(input_binary=False) uses the legacy background-subtraction binarizer "
                                      "(cc_stability_estimator.py:48), which is outside the FCN hot path")
        h, w = self.height, self.width
        if isinstance(img, PackedMask):
            kind = "scan" if img._words is None else "words"
            src = img.scan if kind == "scan" else img.words
        else:
            kind, src = "dense", np.asarray(img)
        if tuple(img.shape[:2]) != (h, w):
            raise ValueError("frame of shape %s handed to a %dx%d estimator" % (tuple(img.shape), w, h))
        if self._staged and kind != self._stage_kind:
            self._submit_staged()
        self._stage_kind = kind
        buf = self._h_stage.get(kind)
        if buf is None:
            shape = {"scan": (h, 1 + (w + 7) // 8), "words": (h, self._engine.wpr * 4), "dense": (h, w)}[kind]
            buf = self._h_stage[kind] = torch.empty((2, self._max_batch) + shape, dtype=torch.uint8).pin_memory()
            self._d_stage[kind] = torch.empty((2, self._max_batch) + shape, dtype=torch.uint8, device=self._engine.device)
        dst = buf[self._slot, self._staged].numpy()
        if kind == "dense":
            dst[...] = src                                               # am_pack_mask_u8 packs `!= 0` (labeler.py:126 labels `content != 0`)
        else:
            dst[...] = src.view(np.uint8).reshape(dst.shape)
        self._staged += 1
        self.img_idx += 1
        if self._staged == self._max_batch:
            self._submit_staged()
        if self.verbose:
            print("[" + str(self.img_idx) + "]", end="\r")

    def _submit_staged(self):
        """Upload the staged frames of the current slot and run them (asynchronously)."""
        n, kind, k = self._staged, self._stage_kind, self._slot
        if n == 0:
            return
        if self._inflight[k] is not None:
            self._collect(k)
        eng = self._engines[k]
        lib = _lib.lib()
        main = torch.cuda.current_stream(eng.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(eng.device)
            self._ev_in = [torch.cuda.Event() for _ in range(2)]
            self._d_bits = torch.zeros((2, self._max_batch, self.height, eng.wpr), dtype=torch.int32, device=eng.device)
        with torch.cuda.stream(self._copy_stream):
            self._d_stage[kind][k, :n].copy_(self._h_stage[kind][k, :n], non_blocking=True)
            self._ev_in[k].record(self._copy_stream)
        main.wait_event(self._ev_in[k])
        st = ctypes.c_void_p(main.cuda_stream)
        d = self._d_stage[kind][k]
        if kind == "scan":
            bits = self._d_bits[k, :n]
            _lib.check(lib.am_png1_scanlines_to_bits(d.data_ptr(), n, self.height, self.width, bits.data_ptr(), st), "am_png1_scanlines_to_bits")
        elif kind == "words":
            bits = d[:n].view(torch.int32).view(n, self.height, eng.wpr)
        else:
            bits = eng.pack(d[:n])
        self._staged = 0
        self._run(bits, k)
        self._slot = k ^ 1
        if self._inflight[k ^ 1] is not None:                             # the other slot's batch finished long ago: take its results now
            self._collect(k ^ 1)

    def add_frames(self, masks):
        """masks: uint8 (n, H, W) numpy array or CUDA tensor, ink != 0."""
        self._submit_staged()
        if isinstance(masks, np.ndarray):
            masks = torch.from_numpy(np.ascontiguousarray(masks != 0).view(np.uint8)).cuda(non_blocking=True)
        eng = self._engine
        for s in range(0, masks.shape[0], eng.max_batch):
            chunk = masks[s:s + eng.max_batch].contiguous()
            self.add_packed(eng.pack(chunk))

    def add_packed(self, bits, defer=False):
        """bits: bit-packed CUDA tensor (n <= max_batch, H, WPR) -- what the FCN epilogue emits.  defer=True: return as soon as the
        GPU work is enqueued (results are read back when the slot is reused, or by flush())."""
        self._submit_staged()
        k = self._slot
        if self._inflight[k] is not None:
            self._collect(k)
        self.img_idx += bits.shape[0]
        self._run(bits, k)
        self._slot = k ^ 1
        if not defer:
            self.flush()                                                 # (defer: the slot's results are read when it comes round again)

    def _run(self, bits, k):
        eng, est = self._engines[k], self._est
        n = bits.shape[0]
        eng.label(bits, want_labels=False, sync=False)
        est.add_frames(eng, 0, n)
        self._ev[k].record(torch.cuda.current_stream(eng.device))
        self._inflight[k] = n

    def _collect(self, k):
        """Results of the batch engine k holds: counts, rows and crops to the host (raw arrays; objects are built lazily)."""
        n = self._inflight[k]
        if n is None:
            return
        eng = self._engines[k]
        if self._d2h_stream is None:
            self._d2h_stream = torch.cuda.Stream(eng.device)
        # read-backs on their own stream: on the compute stream they would queue up behind whatever has been enqueued since
        # (the worker that feeds this estimator has the NEXT batch's FCN there) and stall the host for a whole step
        self._d2h_stream.wait_event(self._ev[k])
        with torch.cuda.stream(self._d2h_stream):
            eng.batch = n
            eng.read_counts()
            rows, offs = eng.packed_rows(n)
            rows_h, offs_h = rows.cpu().numpy(), offs.cpu().numpy()
            for f in range(n):
                crops = eng.crops(f) if eng.counts[f, 2] else None
                self._raw.append((rows_h[offs_h[f]:offs_h[f + 1]], crops))
        self._inflight[k] = None

    def flush(self):
        """Run whatever is staged and bring every outstanding result to the host; raises on a device-side capacity flag."""
        if getattr(self, "_engines", None) is None:                      # unpickled / host-only object: nothing lives on a device
            return
        self._submit_staged()
        first = self._slot                                               # the older batch sits in the slot that would be reused next
        for k in (first, first ^ 1):
            self._collect(k)
        st = self._est.state()                                           # raises on a capacity / hand-off flag
        self._tempo = st["tempo_count"]
        n_frames = self._absorbed + len(self._raw)
        if st["img_idx"] != n_frames:
            raise RuntimeError("device estimator has seen %d frames, the host view %d" % (st["img_idx"], n_frames))
        self._n_unique_dev = st["n_unique"]

    def _materialise(self):
        """Host view of everything the device has processed so far (cheap guard in front of every attribute read)."""
        if getattr(self, "_engines", None) is None or (not self._raw and not self._staged and self._inflight == [None, None]):
            return
        self._materialise_now()

    @gc_paused
    def _materialise_now(self):
        """Vectorised over all outstanding frames: one ConnectedComponent per NEW unique (its first-seen instance), the (frame, label)
        lists per unique, the per-frame tables stage 03 paints from; the per-frame instance lists stay raw (see _LazyFrames)."""
        self.flush()
        raw, self._raw = self._raw, []
        if not raw:
            return
        t0 = self._absorbed
        if len(self._uniques) != len(self._ufirst):
            raise RuntimeError("frames were added after split_stable_cc_by_gaps renumbered the unique CCs (stage 03 runs on a finished estimator)")
        lazy = isinstance(self._per_frame, _LazyFrames)
        for rows, crops in raw:
            self._frame_tables.append((np.ascontiguousarray(rows[:, 2:6]), rows[:, 7].astype(np.uint64),
                                       crops if crops is not None else np.zeros(1, np.uint32)))
            if lazy:
                self._per_frame._push_raw(rows, crops)
        counts = np.fromiter((len(r[0]) for r in raw), dtype=np.int64, count=len(raw))
        self._absorbed += len(raw)
        if counts.sum():
            rows = np.concatenate([r[0] for r in raw if len(r[0])])
            frame_of = np.repeat(np.arange(t0, t0 + len(raw)), counts)
            u, lab = rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64)
            n_old = len(self._ufirst)
            # new uniques are numbered in order of appearance: the first row carrying each index >= n_old is its first-seen instance
            new_u, first_row = np.unique(u, return_index=True)
            first_row = first_row[new_u >= n_old]
            founders = rows[first_row]
            self._uboxes = np.concatenate([self._uboxes, founders[:, 2:6].astype(np.int64)])      # min_x, max_x, min_y, max_y per unique
            zero = np.float32(0.0)
            for f, (_, lb, x0, x1, y0, y1, size, off), b32 in zip((frame_of[first_row] - t0).tolist(), founders.tolist(),
                                                                  founders[:, 2:7]):            # (:111-124)
                words = ((x1 >> 5) - (x0 >> 5) + 1) * (y1 - y0 + 1)
                cc = ConnectedComponent(lb - 1, b32[0], b32[1], b32[2], b32[3], b32[4], packed=raw[f][1][off:off + words])
                cc.start_time = cc.end_time = zero
                self._uniques.append(cc)
                self._uframes.append([])
                self._ufirst.append((f + t0, lb))
            order = np.argsort(u, kind="stable")                          # (frame, label) order is kept inside every unique
            us, fs, ls = u[order], frame_of[order].tolist(), lab[order].tolist()
            cuts = np.flatnonzero(np.diff(us)) + 1
            starts = np.concatenate([[0], cuts]).tolist()
            ends = np.concatenate([cuts, [len(us)]]).tolist()
            uframes = self._uframes
            for a, b in zip(starts, ends):                                # (:99-104) / (:113)
                uframes[int(us[a])].extend(zip(fs[a:b], ls[a:b]))
        if not lazy:                                                      # somebody replaced cc_idx_per_frame by a plain list
            for k, (rows_k, crops_k) in enumerate(raw):
                self._per_frame.append(self._build_row(t0 + k, rows_k, crops_k))
        if len(self._ufirst) != self._n_unique_dev:
            raise RuntimeError("device estimator holds %d uniques, the host view %d" % (self._n_unique_dev, len(self._ufirst)))

    def _build_row(self, t, rows, crops):
        """[(unique_idx, ConnectedComponent)] of frame t; the instance that founded a unique IS that unique's object (:113-115)."""
        current = []
        for r in rows:
            u, lab, x0, x1, y0, y1, size, off = (int(v) for v in r)
            if u < len(self._ufirst) and self._ufirst[u] == (t, lab):
                cc = self._uniques[u]
            else:
                words = ((x1 >> 5) - (x0 >> 5) + 1) * (y1 - y0 + 1)
                cc = ConnectedComponent(lab - 1, np.int32(x0), np.int32(x1), np.int32(y0), np.int32(y1), np.int32(size),
                                        packed=crops[off:off + words])
                cc.start_time = cc.end_time = np.float32(0.0)
            current.append((u, cc))
        return current

    def finish_processing(self):                                         # :158-164
        self.flush()
        if self.verbose:
            print(".")
        print("Total CC merges tested: " + str(self.tempo_count))
        self.fake_age = None
