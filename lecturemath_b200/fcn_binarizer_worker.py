"""FCN_LectureNet_Binarizer drop-in (R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:5-80):
the VideoProcessor worker protocol -- initialize(width, height), handleFrame(frame, last_frame, v_index, abs_time,
rel_time, abs_frame_idx), getWorkName(), finalize(), set_debug_mode(...) -- and the attributes stage 01 reads back
(frame_times, frame_indices, compressed_frames, lecture_net; pre_ST3D_v3.0_01_binarize.py:49-55)."""
import cv2
import numpy as np


class FCN_LectureNet_Binarizer:
    def __init__(self, lecture_net, keep_others=True, png="device"):
        """png: "device" = the PNG bytes of compressed_frames are written on the GPU from the bit-packed mask (csrc/png.cu: 1-bit
        grayscale, stored deflate; decodes to the same pixels), "cv2" = cv2.imencode on the host as the reference does (:56)."""
        self.png = png
        self._png_encoder = None
        self.width = self.height = 0
        self.frame_count = 0
        self.lecture_net = lecture_net
        self.keep_others = keep_others
        self.last_binary = self.last_text = self.last_rec = None
        self.frame_times = self.frame_indices = self.compressed_frames = None
        self.debug_mode = False
        self.debug_start = self.debug_end = 0.0
        self.debug_out_dir = None
        self.debug_video_name = ""

    def initialize(self, width, height):
        self.width, self.height = width, height
        self.frame_count = 0
        self.frame_times, self.frame_indices, self.compressed_frames = [], [], []

    def set_debug_mode(self, active, start_time, end_time, out_dir, video_name):
        self.debug_mode, self.debug_start, self.debug_end = active, start_time, end_time
        self.debug_out_dir, self.debug_video_name = out_dir, video_name

    def handleFrame(self, frame, last_frame, v_index, abs_time, rel_time, abs_frame_idx):
        """BGR frame -> ink mask (ink = 255) -> PNG bytes appended (the 01->02 wire format, :50-64)."""
        self.frame_count += 1
        net = self.lecture_net
        h, w = frame.shape[:2]
        # frames above 2.5 MP: binarize_frames halves them (LANCZOS) and resizes the mask back (NEAREST) on the device
        plan = net.binarize_frames(np.ascontiguousarray(frame)[None], want_others=self.keep_others)
        binary, text_mask, rec_img = net.masks_from_plan(plan, 0, self.keep_others)
        if self.png == "device":
            if self._png_encoder is None or (self._png_encoder.width, self._png_encoder.height) != (w, h):
                from .wire import PngEncoder
                self._png_encoder = PngEncoder(w, h, 1, plan.bits.device)
            raw_data = self._png_encoder.encode(plan.bits[:1])[0]
        else:
            flag, raw_data = cv2.imencode(".png", binary)
        self.last_binary, self.last_text, self.last_rec = binary, text_mask, rec_img
        self.compressed_frames.append(raw_data)
        self.frame_indices.append(abs_frame_idx)
        self.frame_times.append(abs_time)
        if self.debug_mode and self.debug_start <= abs_time <= self.debug_end:
            cv2.imwrite(self.debug_out_dir + "/binary_" + self.debug_video_name + "_" + str(self.frame_count) + ".png", binary)

    def getWorkName(self):
        return "FCN_LectureNet Frame Binarizer"

    def finalize(self):
        pass
