"""VideoProcessor drop-in (R/AccessMath/preprocessing/video_processor/video_processor.py:21-199), the frame source of stage 01
(SURVEY.md 8f rank 3): decode + fps sampling + forced resolution + the worker protocol.

What is reproduced exactly (pinned by tests/golden/video_sampling.json, captured from the unmodified reference on seeded videos):
  * which frames reach the worker: every int(video_fps / frames_per_second)-th decoded frame (`frames_per_second` None or <= 0: all
    frames, :92-97), the FIRST sampled frame of the whole run is only remembered as `last_frame`, never handed over (:167-170);
  * their labels: frame_time = time of all previous files + CAP_PROP_POS_MSEC after the read, frame_idx = frames of all previous
    files + CAP_PROP_POS_FRAMES after the read (:160-169), file lengths taken after seeking to the end (:182-191);
  * the probe the reference makes between skipping by grab() and skipping by seeking (:106-142) including its side effect: the second
    iteration seeks to `position + jump - 1`, which with jump = 0 re-reads the frame just read;
  * `limit` (:105), `force_no_seek`, the resolution check across files (:62-87) and initialize() / finalize() of the worker.
Decoding stays cv2.VideoCapture on the host (this image has FFmpeg but no NVDEC headers).  What moves to the GPU is the forced
resize: a worker that declares `accepts_unresized_frames` (this package's FCN_LectureNet_Binarizer) receives the decoded frame as it
is and resizes whole batches on the device (am_resize_linear_u8 = OpenCV's INTER_LINEAR fixed-point algorithm); any other worker gets
`cv2.resize(frame, (forced_width, forced_height))` on the host exactly as the reference does (:164-165)."""
import time

import cv2


class VideoProcessor:
    def __init__(self, file_list, frames_per_second=1):
        self.file_list = file_list
        self.frames_per_second = frames_per_second
        self.forced_width = None
        self.forced_height = None

    def force_resolution(self, width, height):
        self.forced_width, self.forced_height = width, height

    def checkError(self):
        print("Works?")

    # ---- sampling -------------------------------------------------------------------------------------------------
    def _jump(self, capture):
        fps = capture.get(cv2.CAP_PROP_FPS)
        if self.frames_per_second is None or self.frames_per_second <= 0.0:
            return 0
        return int(fps / self.frames_per_second)

    def sampled_frames(self, limit=0, force_no_seek=False, on_open=None):
        """Generator over the frames the reference samples, in order: (frame, video_idx, abs_time, rel_time, abs_frame_idx, read_no),
        read_no = running count of sampled frames (0 = the one the reference skips).  on_open(video_idx, width, height) is called
        when a file is opened, before its first frame."""
        read_no = -1
        base_time, base_frames = 0.0, 0
        for video_idx, path in enumerate(self.file_list):
            capture = cv2.VideoCapture(path)
            if on_open is not None:
                on_open(video_idx, int(capture.get(cv2.CAP_PROP_FRAME_WIDTH)), int(capture.get(cv2.CAP_PROP_FRAME_HEIGHT)))
            jump = self._jump(capture)
            # how to skip `jump - 1` frames: "grab" them one by one or "seek"; the first two iterations try one each, the faster stays
            mode, spent = ("grab" if force_no_seek else "probe-grab"), {"grab": 0.0, "seek": 0.0}
            rel_time, rel_frame = 0.0, 0
            while limit == 0 or read_no < limit:
                ok = True
                if mode in ("probe-seek", "seek"):
                    t0 = time.perf_counter()
                    ok = capture.set(cv2.CAP_PROP_POS_FRAMES, capture.get(cv2.CAP_PROP_POS_FRAMES) + jump - 1)
                    spent["seek"] += time.perf_counter() - t0
                else:
                    t0 = time.perf_counter()
                    for _ in range(jump - 1):
                        ok = capture.grab()
                        if not ok:
                            break
                        rel_time, rel_frame = capture.get(cv2.CAP_PROP_POS_MSEC), capture.get(cv2.CAP_PROP_POS_FRAMES)
                    spent["grab"] += time.perf_counter() - t0
                if mode == "probe-grab":
                    mode = "probe-seek"
                elif mode == "probe-seek":
                    mode = "grab" if spent["grab"] < spent["seek"] else "seek"
                    print("Grabbing frames to jump" if mode == "grab" else "Jumping to frames directly")
                flag, frame = capture.read() if ok else (False, None)
                if not flag:
                    break                                                     # end of this file
                read_no += 1
                rel_time, rel_frame = capture.get(cv2.CAP_PROP_POS_MSEC), capture.get(cv2.CAP_PROP_POS_FRAMES)
                yield frame, video_idx, base_time + rel_time, rel_time, int(base_frames + rel_frame), read_no
            capture.set(cv2.CAP_PROP_POS_AVI_RATIO, 1.0)                      # length of the file as the container reports it ...
            length, frames = capture.get(cv2.CAP_PROP_POS_MSEC), capture.get(cv2.CAP_PROP_POS_FRAMES)
            if length < rel_time or frames < rel_frame:                       # ... unless the reads got further than that
                length, frames = rel_time, rel_frame
            base_time += length
            base_frames += frames
            capture.release()

    def doProcessing(self, video_worker, limit=0, verbose=False, force_no_seek=False):
        size = {}
        raw_ok = bool(getattr(video_worker, "accepts_unresized_frames", False))

        def on_open(video_idx, cap_w, cap_h):
            if not size:
                if self.forced_width is not None:
                    size["w"], size["h"] = self.forced_width, self.forced_height
                else:
                    size["w"], size["h"] = cap_w, cap_h
                video_worker.initialize(size["w"], size["h"])
            elif self.forced_width is None and (size["w"], size["h"]) != (cap_w, cap_h):
                raise Exception("All video files on the list must have the same resolution")
            size["resize"] = self.forced_width is not None and (cap_w, cap_h) != (self.forced_width, self.forced_height)

        if verbose:
            print("Video processing for " + video_worker.getWorkName() + " has begun")
        t_start = time.time()
        last_frame = None
        for frame, video_idx, abs_time, rel_time, abs_idx, read_no in self.sampled_frames(limit, force_no_seek, on_open):
            if size["resize"] and not raw_ok:
                frame = cv2.resize(frame, (self.forced_width, self.forced_height))
            if read_no > 0:
                video_worker.handleFrame(frame, last_frame, video_idx, abs_time, rel_time, abs_idx)
                if verbose and read_no % 50 == 0:
                    print("Frames Processed = " + str(read_no) + ", Video Time = " + _stamp(abs_time))
            last_frame = frame
        video_worker.finalize()
        if verbose:
            print("Video processing for " + video_worker.getWorkName() + " completed: " + _stamp((time.time() - t_start) * 1000.0))


def _stamp(ms):
    """hh:mm:ss.d like the reference's TimeHelper.stampToStr."""
    s = ms / 1000.0
    return "%02d:%02d:%04.1f" % (int(s // 3600), int((s % 3600) // 60), s % 60)
