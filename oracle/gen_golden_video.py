"""oracle/gen_golden_video.py -- regenerates tests/golden/video_sampling.json by running the UNMODIFIED reference VideoProcessor
(R/AccessMath/preprocessing/video_processor/video_processor.py) over seeded synthetic videos with a recording worker.

Run in the build container only (needs /root/reference):   python oracle/gen_golden_video.py
The videos themselves are not committed: make_video() rewrites them (cv2.VideoWriter, MJPG) wherever the test runs, and every frame
carries its own index as a block pattern that survives JPEG compression, so a log entry names the decoded frame independent of
the codec's exact bytes."""
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/ACCESS2021_release"
sys.path.insert(0, REPO)

VIDEOS = {"a": (100, 320, 180, 30.0), "b": (47, 320, 180, 30.0), "c": (40, 256, 144, 25.0)}      # frames, width, height, fps
# (file list, frames_per_second, forced resolution, limit, force_no_seek)
CASES = [
    (["a", "b"], 10, None, 0, True),
    (["a", "b"], 10, None, 0, False),
    (["a", "b"], None, None, 0, True),
    (["a"], 0, None, 12, False),              # "all frames" + the seek probe: the second iteration re-reads frame 0
    (["a", "b"], 7, None, 0, True),
    (["a"], 45, None, 0, True),               # more samples per second than the file has frames: int(30 / 45) = 0 -> every frame
    (["a", "b"], 10, None, 5, True),
    (["a", "c"], 5, (320, 180), 0, True),     # second file has another size and frame rate: forced resolution
    (["c"], 12.5, (480, 270), 0, True),
]


def make_video(path, n, w, h, fps):
    import cv2
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
    if not vw.isOpened():
        raise RuntimeError("cv2.VideoWriter cannot write MJPG/avi here")
    for i in range(n):
        fr = np.full((h, w, 3), 30, np.uint8)
        for b in range(8):
            if (i >> b) & 1:
                fr[h // 4:3 * h // 4, (2 + 3 * b) * w // 28:(4 + 3 * b) * w // 28] = 220
        vw.write(fr)
    vw.release()


def frame_number(frame):
    """The index a frame of make_video carries (works at any resolution the frame was resized to)."""
    h, w = frame.shape[:2]
    return sum(1 << b for b in range(8) if frame[h // 2, (3 + 3 * b) * w // 28, 0] > 128)


def write_videos(directory):
    paths = {}
    for name, (n, w, h, fps) in VIDEOS.items():
        paths[name] = os.path.join(directory, name + ".avi")
        make_video(paths[name], n, w, h, fps)
    return paths


class Recorder:
    """Worker that logs what VideoProcessor hands it."""

    def initialize(self, width, height):
        self.size, self.log, self.finalized = [width, height], [], False

    def handleFrame(self, frame, last_frame, v_index, abs_time, rel_time, abs_frame_idx):
        self.log.append([v_index, round(float(abs_time), 3), round(float(rel_time), 3), int(abs_frame_idx), frame_number(frame),
                         -1 if last_frame is None else frame_number(last_frame), list(frame.shape)])

    def getWorkName(self):
        return "recorder"

    def finalize(self):
        self.finalized = True


def run_case(VideoProcessor, paths, case):
    files, fps, forced, limit, no_seek = case
    vp = VideoProcessor([paths[f] for f in files], fps)
    if forced is not None:
        vp.force_resolution(*forced)
    rec = Recorder()
    vp.doProcessing(rec, limit, False, no_seek)
    return {"size": rec.size, "finalized": rec.finalized, "log": rec.log}


def main():
    sys.path.insert(0, REF)
    from AccessMath.preprocessing.video_processor.video_processor import VideoProcessor
    import cv2
    out = {"opencv": cv2.__version__, "cases": []}
    with tempfile.TemporaryDirectory() as d:
        paths = write_videos(d)
        for case in CASES:
            out["cases"].append(run_case(VideoProcessor, paths, case))
    with open(os.path.join(REPO, "tests", "golden", "video_sampling.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases;", sum(len(c["log"]) for c in out["cases"]), "log entries")


if __name__ == "__main__":
    main()
