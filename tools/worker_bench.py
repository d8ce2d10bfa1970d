"""Throughput of the reference's own per-frame worker protocol (VideoProcessor -> worker.handleFrame, one 1080p BGR frame per
call, PNG bytes appended to compressed_frames) through the drop-in FCN_LectureNet_Binarizer, with the PNG written on the device
(csrc/png.cu) and with cv2.imencode on the host as the reference does.   python tools/worker_bench.py [--frames 48]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=48)
    args = ap.parse_args()
    import torch
    from bench import make_net
    from lecturemath_b200 import synth
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    from lecturemath_b200.helper import Helper
    net = make_net().cuda()
    frames = list(synth.whiteboard_frames(16, 1080, 1920, seed=1234))
    line = {"workload": "FCN_LectureNet_Binarizer.handleFrame, 1080p frames one at a time (the reference's worker protocol)"}
    decoded = {}
    for mode in ("device", "cv2"):
        w = FCN_LectureNet_Binarizer(net, keep_others=False, png=mode)
        w.initialize(1920, 1080)
        for i in range(4):
            w.handleFrame(frames[i % 16], None, 0, 0.0, 0.0, i)
        w.initialize(1920, 1080)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.frames):
            w.handleFrame(frames[i % 16], None, 0, 33.3 * i, 33.3 * i, i)
        dt = time.perf_counter() - t0
        line[mode + "_frames_per_s"] = round(args.frames / dt, 1)
        line[mode + "_png_kb_per_frame"] = round(float(np.mean([len(r) for r in w.compressed_frames])) / 1e3, 1)
        decoded[mode] = Helper.decompress_binary_images(w.compressed_frames[:4])
    line["decoded_identical"] = all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(decoded["device"], decoded["cv2"]))
    # the same calls through the batching worker (handleFrame stages, one GPU step per 8 frames, finalize drains)
    for rep in range(2):
        w = FCN_LectureNet_Binarizer(net, batch=8)
        w.initialize(1920, 1080)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.frames):
            w.handleFrame(frames[i % 16], None, 0, 33.3 * i, 33.3 * i, i)
        w.finalize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    line["device_batch8_frames_per_s"] = round(args.frames / dt, 1)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
