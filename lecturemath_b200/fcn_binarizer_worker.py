"""FCN_LectureNet_Binarizer drop-in (R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:5-80):
the VideoProcessor worker protocol -- initialize(width, height), handleFrame(frame, last_frame, v_index, abs_time,
rel_time, abs_frame_idx), getWorkName(), finalize(), set_debug_mode(...) -- and the attributes stage 01 reads back
(frame_times, frame_indices, compressed_frames, lecture_net; pre_ST3D_v3.0_01_binarize.py:49-55).

The reference drives this one frame at a time (video_processor.py:167-170) and only reads the results after finalize()
(:196).  batch = 1 reproduces that literally (every call runs the network and returns with its PNG appended).  batch > 1 keeps
the SAME calls but makes them asynchronous: handleFrame copies the frame into a pinned staging slot and returns; every `batch`
frames one H2D copy + one FCN step + one device PNG encode are enqueued on the GPU, their files are appended while the next
batch is being staged, and finalize() drains (a partial last batch included).  The lists a caller reads after finalize() have
the per-frame protocol's entries in the same order; only their growth in between lags by up to two batches.  (The network's tiling
depends on the batch size, so a near-threshold pixel may come out differently than with batch = 1 -- both within the 0.1 % bar.)"""
import ctypes

import cv2
import numpy as np
import torch


class FCN_LectureNet_Binarizer:
    # VideoProcessor (this package's) hands such a worker the decoded frame as it is when a resolution is forced: frames whose size
    # differs from initialize(width, height) are resized on the device (am_resize_linear_u8 = cv2.resize's INTER_LINEAR algorithm,
    # R/AccessMath/preprocessing/video_processor/video_processor.py:164-165), whole batches at a time
    accepts_unresized_frames = True

    def __init__(self, lecture_net, keep_others=None, png="device", batch=1, estimator=None):
        """png: "device" = compressed_frames are written on the GPU from the bit-packed mask (csrc/png.cu: 1-bit grayscale, deflate;
                decodes to the same pixels), "device-stored" = the same with stored deflate blocks, "cv2" = cv2.imencode on the host
                as the reference does (:56; batch = 1 only).
        batch: frames per GPU step (see the module docstring).
        keep_others: keep last_text / last_rec (the diagnostic outputs of :58-60); default True for batch = 1, False otherwise
                (they cost two fp32 full-frame read-backs per frame).
        estimator: optional lecturemath_b200 CCStabilityEstimator that receives every batch's bit-packed masks on the device
                (stage 01 and stage 02 in one pass, no PNG round trip); compressed_frames is still filled."""
        self.png = png
        self.batch = int(batch)
        if self.batch > 1 and png == "cv2":
            raise ValueError("png='cv2' encodes on the host frame by frame: use batch=1")
        self.keep_others = (self.batch == 1) if keep_others is None else bool(keep_others)
        self.estimator = estimator
        self._png_encoder = None
        self.width = self.height = 0
        self.frame_count = 0
        self.lecture_net = lecture_net
        self.last_binary = self.last_text = self.last_rec = None
        self.frame_times = self.frame_indices = self.compressed_frames = None
        self.debug_mode = False
        self.debug_start = self.debug_end = 0.0
        self.debug_out_dir = None
        self.debug_video_name = ""
        self._pipe = None

    def initialize(self, width, height):
        self.width, self.height = width, height
        self.frame_count = 0
        self.frame_times, self.frame_indices, self.compressed_frames = [], [], []
        self._pipe = None

    def set_debug_mode(self, active, start_time, end_time, out_dir, video_name):
        self.debug_mode, self.debug_start, self.debug_end = active, start_time, end_time
        self.debug_out_dir, self.debug_video_name = out_dir, video_name

    # ---- the reference's per-frame call -----------------------------------------------------------------------
    def handleFrame(self, frame, last_frame, v_index, abs_time, rel_time, abs_frame_idx):
        """BGR frame -> ink mask (ink = 255) -> PNG bytes appended (the 01->02 wire format, :50-64)."""
        self.frame_count += 1
        if self.batch > 1:
            return self._stage(frame, abs_time, abs_frame_idx)
        net = self.lecture_net
        if self.width and self.height and frame.shape[:2] != (self.height, self.width):
            frame = _resize_on_device(net, frame, self.width, self.height)                   # CUDA tensor (1, H, W, 3)
        h, w = (self.height, self.width) if torch.is_tensor(frame) else frame.shape[:2]
        # frames above 2.5 MP: binarize_frames halves them (LANCZOS) and resizes the mask back (NEAREST) on the device
        plan = net.binarize_frames(frame if torch.is_tensor(frame) else np.ascontiguousarray(frame)[None], want_others=self.keep_others)
        binary, text_mask, rec_img = net.masks_from_plan(plan, 0, self.keep_others)
        if self.png != "cv2":
            if self._png_encoder is None or (self._png_encoder.width, self._png_encoder.height) != (w, h):
                from .wire import PngEncoder
                self._png_encoder = PngEncoder(w, h, 1, plan.bits.device, compress=self.png != "device-stored")
            raw_data = self._png_encoder.encode(plan.bits[:1])[0]
        else:
            flag, raw_data = cv2.imencode(".png", binary)
        if self.estimator is not None:
            self.estimator.add_packed(plan.bits[:1])
        self.last_binary, self.last_text, self.last_rec = binary, text_mask, rec_img
        self.compressed_frames.append(raw_data)
        self.frame_indices.append(abs_frame_idx)
        self.frame_times.append(abs_time)
        self._debug(binary, abs_time, self.frame_count)

    def _debug(self, binary, abs_time, count):
        if self.debug_mode and self.debug_start <= abs_time <= self.debug_end:           # :66-72
            cv2.imwrite(self.debug_out_dir + "/binary_" + self.debug_video_name + "_" + str(count) + ".png", np.asarray(binary))

    def getWorkName(self):
        return "FCN_LectureNet Frame Binarizer"

    def finalize(self):
        if self._pipe is not None:
            self._pipe.drain(self)

    # ---- batch > 1 ----------------------------------------------------------------------------------------------
    def _stage(self, frame, abs_time, abs_frame_idx):
        if self._pipe is not None and self._pipe.in_shape != frame.shape[:2]:      # next file of the list has another capture size
            self._pipe.drain(self)
            self._pipe = None
        if self._pipe is None:
            ih, iw = frame.shape[:2]
            w, h = (self.width, self.height) if (self.width and self.height) else (iw, ih)
            # staging buffers (100 MB of pinned memory at 1080p x 8) are kept with the network and reused by the next worker / video
            key = ("worker_pipe", w, h, iw, ih, self.batch, self.png != "device-stored", self.keep_others)
            pipe = self.lecture_net._plans.get(key)
            if pipe is None or pipe.busy:
                pipe = _BatchPipe(self.lecture_net, w, h, self.batch, self.png != "device-stored", self.keep_others, in_size=(iw, ih))
                self.lecture_net._plans.setdefault(key, pipe)
            pipe.busy = True
            self._pipe = pipe
        self._pipe.push(self, frame, abs_time, abs_frame_idx)


def _resize_on_device(net, frame, width, height):
    """One frame (H, W, 3) uint8 numpy -> CUDA tensor (1, height, width, 3): cv2.resize(frame, (width, height)) on the device."""
    from . import _lib
    lib = _lib.lib()
    if net._device is None:
        net.cuda()
    with torch.cuda.device(net._device):
        src = torch.from_numpy(np.ascontiguousarray(frame)).to(net._device, non_blocking=True)
        dst = torch.empty((1, height, width, 3), dtype=torch.uint8, device=net._device)
        _lib.check(lib.am_resize_linear_u8(src.data_ptr(), 1, frame.shape[0], frame.shape[1], 3, height, width, dst.data_ptr(),
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "am_resize_linear_u8")
    return dst


class _BatchPipe:
    """Two staging slots: while the GPU works on one batch the host fills the other.  Per batch, all on the compute stream unless
    noted: H2D (copy stream) -> [LANCZOS halving] -> FCN -> [mask upscale] -> [estimator] -> PNG encode + file read-back (own stream)."""

    def __init__(self, net, width, height, batch, compress, keep_others, in_size=None):
        from .wire import PngEncoder
        from . import _lib
        self.lib = _lib.lib()
        in_w, in_h = in_size or (width, height)
        self.in_shape = (in_h, in_w)                                      # frames as the caller hands them (before a forced resize)
        self.resize = (in_w, in_h) != (width, height)
        if net._device is None:
            net.cuda()
        self.net, self.width, self.height, self.batch, self.keep_others = net, width, height, batch, keep_others
        self.device = net._device
        self.large = net.large_adapter(batch, height, width)             # > 2.5 MP: FCN at the halved size (FCN_lecturenet.py:434-437)
        self.plan = net.plan(batch, self.large.fcn_height, self.large.fcn_width)
        shape = (batch, in_h, in_w, 3)
        self.h_in = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.d_in = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.d_sized = torch.empty((batch, height, width, 3), dtype=torch.uint8, device=self.device) if self.resize else None
        self.fcn_in = torch.empty_like(self.plan.frames) if self.large.active else None
        # the slot's own copy of the batch's masks: the PNG encoder reads it on its stream while the next FCN step rewrites plan.bits
        self.bits = [torch.zeros((batch, height, self.plan.lib.am_words_per_row(width)), dtype=torch.int32, device=self.device)
                     for _ in range(2)]
        self.enc = [PngEncoder(width, height, batch, self.device, compress) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.out_stream = torch.cuda.Stream(self.device)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.meta = [[], []]                                             # (abs_time, abs_frame_idx, frame_count) per staged frame
        self.inflight = [None, None]                                     # number of frames of the batch the slot's GPU work covers
        self.others = [None, None]
        self.slot = 0
        self.launches = 0
        self.busy = False                                                # owned by one worker between its first frame and finalize()

    def push(self, worker, frame, abs_time, abs_frame_idx):
        k = self.slot
        if self.inflight[k] is not None:                                  # the slot's previous batch: results out before it is reused
            self.collect(worker, k)
        n = len(self.meta[k])
        self.h_in[k][n].copy_(torch.from_numpy(np.ascontiguousarray(frame)))   # torch's CPU copy is multi-threaded: ~0.05 ms per 1080p frame
        self.meta[k].append((abs_time, abs_frame_idx, worker.frame_count))
        if n + 1 == self.batch:
            self.submit(worker, k)

    def submit(self, worker, k):
        n = len(self.meta[k])
        if n == 0:
            return
        plan, main = self.plan, torch.cuda.current_stream(self.device)
        dbg = getattr(self, "debug_timing", None)
        if dbg is not None:
            t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            t[0].record(self.copy_stream)
            t[2].record(main)
        with torch.cuda.stream(self.copy_stream):
            self.d_in[k].copy_(self.h_in[k], non_blocking=True)            # (a partial last batch carries stale frames behind n: ignored)
            self.ev_in[k].record(self.copy_stream)
        main.wait_event(self.ev_in[k])
        if dbg is not None:
            t[1].record(self.copy_stream)
            t[3].record(main)
            dbg.append(t)
        frames = self.d_in[k]
        if self.resize:                                                  # forced resolution (video_processor.py:164-165) on the device
            from . import _lib
            _lib.check(self.lib.am_resize_linear_u8(frames.data_ptr(), self.batch, self.in_shape[0], self.in_shape[1], 3, self.height, self.width,
                                                    self.d_sized.data_ptr(), ctypes.c_void_p(main.cuda_stream)), "am_resize_linear_u8")
            frames = self.d_sized
            self.launches += 1
        if self.large.active:
            self.large.downscale(self.fcn_in, main.cuda_stream, src=frames)
            plan.run(main.cuda_stream, self.keep_others, 128, frames=self.fcn_in, want_logits=False)
            bits = self.large.upscale_bits(plan.bits, main.cuda_stream, out=self.bits[k])
        else:
            plan.run(main.cuda_stream, self.keep_others, 128, frames=frames, want_logits=False)
            bits = self.bits[k]
            bits.copy_(plan.bits, non_blocking=True)
        self.launches += plan.launches_per_run + self.large.launches_per_run + 3
        if worker.estimator is not None:
            worker.estimator.add_packed(bits[:n], defer=True)
        if self.keep_others:                                             # diagnostics of the batch's LAST frame (:58-60)
            self.others[k] = (plan.text_logit[n - 1].clone(), plan.rec[n - 1].clone())
        # the encode kernels run on the compute stream (a side-stream kernel can only start at a conv-kernel boundary and then holds
        # an SM the next persistent conv kernel wants); only the read-back of the finished files is on its own stream
        self.enc[k].launch(bits, n, main, self.out_stream)
        self.inflight[k] = n
        self.slot = k ^ 1

    def collect(self, worker, k):
        n = self.inflight[k]
        if n is None:
            return
        files = self.enc[k].finish(n, self.out_stream)
        for f, (abs_time, abs_idx, count) in zip(files, self.meta[k][:n]):
            worker.compressed_frames.append(f)
            worker.frame_indices.append(abs_idx)
            worker.frame_times.append(abs_time)
            if worker.debug_mode and worker.debug_start <= abs_time <= worker.debug_end:
                from .packed_mask import parse_png1
                worker._debug(parse_png1(f), abs_time, count)
        from .packed_mask import parse_png1
        worker.last_binary = parse_png1(files[-1])                        # lazy 0 / 255 view of the newest mask
        if self.others[k] is not None:
            t, r = self.others[k]
            tm = (torch.sigmoid(t).cpu().numpy() * 255).astype(np.uint8)
            worker.last_text = np.where(tm >= 128, 255, 0).astype(np.uint8)
            rec = r.cpu().numpy() * 0.5 + 0.5
            worker.last_rec = np.clip(rec[:, :, ::-1] * 255, 0, 255).astype(np.uint8)
            self.others[k] = None
        self.meta[k] = []
        self.inflight[k] = None

    def drain(self, worker):
        """finalize(): flush the partial batch, then hand out everything still in flight, oldest first."""
        k = self.slot
        if self.meta[k] and self.inflight[k] is None:
            older = k ^ 1
            self.collect(worker, older)
            self.submit(worker, k)
            self.collect(worker, k)
        else:
            self.collect(worker, k)                                       # (slot k, when in flight, is the older one)
            self.collect(worker, k ^ 1)
        if worker.estimator is not None:
            worker.estimator.flush()
        self.busy = False
