"""oracle/gen_golden_resize.py -- regenerates tests/golden/resize.npz (build container only: needs /root/reference).

Pins the > 2.5 MP branch of FCN_LectureNet.binarize (R/AccessMath/lecturenet_v1/FCN_lecturenet.py:434-437, :481-494):
  * small known-answer cases of the two third-party operators it calls -- PIL.Image.resize(LANCZOS) and
    cv2.resize(INTER_NEAREST) -- at even / odd / tiny sizes (inputs + outputs stored);
  * the UNMODIFIED reference worker (FCN_LectureNet_Binarizer.handleFrame) on a seeded 2000x1300 (2.6 MP) whiteboard frame
    with the tiny golden network: the LANCZOS-halved image it feeds the FCN (sha256 + a sample row) and the full-size ink
    mask it returns (bit-packed).
   python oracle/gen_golden_resize.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import GOLD, REF, import_reference      # noqa: E402

LANCZOS_CASES = [(47, 61, 3), (48, 60, 3), (101, 33, 3), (16, 16, 1), (13, 14, 4), (3, 5, 3), (66, 130, 3)]
NEAREST_CASES = [(23, 30, 47, 61), (24, 30, 48, 60), (50, 16, 101, 33), (8, 8, 16, 16), (33, 65, 66, 130), (33, 65, 67, 131)]
BIG = (1300, 2000)                                              # H, W: 2.6 MP -> FCN at 1000 x 650


def main():
    Labeler, CCStabilityEstimator, FCN_LectureNet, FCN_LectureNet_Binarizer, Configuration = import_reference()
    import cv2
    import PIL.Image
    import torch
    from lecturemath_b200 import synth
    rng = np.random.default_rng(2024)
    out = {}
    for i, (h, w, c) in enumerate(LANCZOS_CASES):
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        if i == 1:
            img[::2] = 255; img[1::2] = 0                       # saturating stripes: exercises both clip8 ends
        pil = PIL.Image.fromarray(img[:, :, 0] if c == 1 else img, {1: "L", 3: "RGB", 4: "CMYK"}[c])
        res = np.asarray(pil.resize((int(w / 2), int(h / 2)), PIL.Image.LANCZOS)).reshape(int(h / 2), int(w / 2), c)
        out["lanczos_in_%d" % i], out["lanczos_out_%d" % i] = img, res
    for i, (sh, sw, dh, dw) in enumerate(NEAREST_CASES):
        m = (rng.random((sh, sw)) < 0.4).astype(np.uint8) * 255
        out["nearest_in_%d" % i] = m
        out["nearest_out_%d" % i] = cv2.resize(m, (dw, dh), interpolation=cv2.INTER_NEAREST)
    # the reference itself on a 2.6 MP frame
    z = np.load(os.path.join(GOLD, "fcn_forward.npz"))
    cfg = Configuration.from_file(os.path.join(GOLD, "fcn_tiny.conf"))
    net = FCN_LectureNet.CreateFromConfig(cfg, 3, False)
    net.load_state_dict({k[len("tiny_sd/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("tiny_sd/")})
    net.eval()
    frame = next(iter(synth.whiteboard_frames(1, BIG[0], BIG[1], seed=11)))
    seen = {}
    orig = FCN_LectureNet.prepare_image

    def spy(pil):
        seen["img"] = np.asarray(pil).copy()
        return orig(pil)
    FCN_LectureNet.prepare_image = staticmethod(spy)
    worker = FCN_LectureNet_Binarizer(net)
    worker.initialize(BIG[1], BIG[0])
    worker.handleFrame(frame, None, 0, 0.0, 0.0, 0)
    FCN_LectureNet.prepare_image = staticmethod(orig)
    small = seen["img"]                                          # RGB, 650 x 1000
    assert small.shape == (650, 1000, 3) and worker.last_binary.shape == BIG
    out["big_shape"] = np.array(BIG)
    out["big_seed"] = np.array(11)
    out["big_frame_sha256"] = np.frombuffer(hashlib.sha256(frame.tobytes()).digest(), dtype=np.uint8)
    out["big_small_rgb_sha256"] = np.frombuffer(hashlib.sha256(small.tobytes()).digest(), dtype=np.uint8)
    out["big_small_rgb_rows"] = small[[0, 1, 324, 648, 649]]
    out["big_ink_bits"] = np.packbits(worker.last_binary > 0, axis=-1)
    print("2.6 MP frame: ink %.2f %%" % (100.0 * (worker.last_binary > 0).mean()))
    np.savez_compressed(os.path.join(GOLD, "resize.npz"), **out)
    print("wrote", os.path.join(GOLD, "resize.npz"), os.path.getsize(os.path.join(GOLD, "resize.npz")), "bytes")


if __name__ == "__main__":
    main()
