"""oracle/gen_golden.py -- regenerates tests/golden/*.npz by running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py
The reference has no tests or golden vectors of its own, so these captured outputs of
the reference -- scipy.ndimage.label + CC_AgeBoundaries + Labeler + CCStabilityEstimator + FCN_LectureNet
-- are what pins the oracle (oracle/cc_oracle.py, oracle/fcn_oracle.py) and, through it, the CUDA path.
/root/reference cannot travel to the GPU box; the fixtures can.
"""
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/ACCESS2021_release"
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)


def import_reference():
    """The reference CDLL-loads './accessmath_lib.so' relative to the CWD at import time (labeler.py:24);
    the shipped .so is a Windows DLL, so give it the ELF build of its own C file (oracle/Makefile)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "--no-print-directory"], stdout=subprocess.DEVNULL)
    tmp = tempfile.mkdtemp(prefix="amref_")
    shutil.copy(os.path.join(REPO, "oracle", "_ref", "accessmath_lib_ref.so"), os.path.join(tmp, "accessmath_lib.so"))
    os.chdir(tmp)
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from AccessMath.preprocessing.content.labeler import Labeler
    from AccessMath.preprocessing.content.cc_stability_estimator import CCStabilityEstimator
    from AccessMath.lecturenet_v1.FCN_lecturenet import FCN_LectureNet
    from AccessMath.preprocessing.video_worker.FCN_lecturenet_binarizer import FCN_LectureNet_Binarizer
    from AM_CommonTools.configuration.configuration import Configuration
    return Labeler, CCStabilityEstimator, FCN_LectureNet, FCN_LectureNet_Binarizer, Configuration


def estimator_tables(est):
    """Flatten the reference estimator's state into arrays."""
    per_frame = []      # rows: frame, unique_idx, raw_label, min_x, max_x, min_y, max_y, size
    for t, frame in enumerate(est.cc_idx_per_frame):
        for u, cc in frame:
            per_frame.append((t, u, cc.cc_id + 1, cc.min_x, cc.max_x, cc.min_y, cc.max_y, cc.size))
    uframes = []        # rows: unique_idx, frame, raw_label (in list order)
    for u, lst in enumerate(est.unique_cc_frames):
        for t, lab in lst:
            uframes.append((u, t, lab))
    uniq = [(cc.cc_id + 1, cc.min_x, cc.max_x, cc.min_y, cc.max_y, cc.size) for cc in est.unique_cc_objects]
    return (np.array(per_frame, dtype=np.int64).reshape(-1, 8), np.array(uframes, dtype=np.int64).reshape(-1, 3),
            np.array(uniq, dtype=np.int64).reshape(-1, 6), int(est.tempo_count))


def gen_cc(Labeler, CCStabilityEstimator):
    import scipy.ndimage
    from lecturemath_b200 import synth
    # 1. known-answer masks: diagonal pixels are separate CCs; U-shape keeps first raster label
    diag = np.zeros((6, 8), np.uint8)
    for i in range(6):
        diag[i, i] = 255
    diag[2, 5:8] = 255
    ushape = np.zeros((5, 9), np.uint8)
    ushape[0:4, 1] = 255; ushape[0:4, 7] = 255; ushape[3, 1:8] = 255; ushape[0, 4] = 255
    kat = {}
    for name, m in (("diag", diag), ("ushape", ushape)):
        lab, n = scipy.ndimage.label(m)
        kat[name + "_mask"] = m; kat[name + "_labels"] = lab.astype(np.int32); kat[name + "_n"] = n
    np.savez_compressed(os.path.join(GOLD, "cc_known_answer.npz"), **kat)

    # 2. label + stats + crops on seeded masks (incl. ragged widths not multiple of 32, empty, full)
    out = {}
    cases = {"blob_96x128": next(iter(synth.random_blob_masks(1, 96, 128, seed=5))),
             "blob_67x121": next(iter(synth.random_blob_masks(1, 67, 121, seed=6))),
             "glyph_180x320": next(iter(synth.glyph_masks(1, 180, 320, seed=0, occluder_w=0))),
             "empty_40x70": np.zeros((40, 70), np.uint8),
             "full_33x65": np.full((33, 65), 255, np.uint8)}
    rng = np.random.default_rng(11)
    cases["noise_50x97"] = (rng.random((50, 97)) < 0.55).astype(np.uint8) * 255
    for name, m in cases.items():
        lab, n = scipy.ndimage.label(m)
        ages = rng.random(m.shape).astype(np.float32) * 100
        ccs = Labeler.extractSpatioTemporalContent(m, ages)
        tab = np.array([(c.cc_id + 1, c.min_x, c.max_x, c.min_y, c.max_y, c.size) for c in ccs], dtype=np.int64).reshape(-1, 6)
        out[name + "_mask"] = m; out[name + "_labels"] = lab.astype(np.int32); out[name + "_n"] = n
        out[name + "_ages"] = ages; out[name + "_table"] = tab
        out[name + "_minage"] = np.array([c.start_time for c in ccs], dtype=np.float32)
        out[name + "_crops"] = (np.concatenate([c.img.ravel() for c in ccs]) if ccs else np.zeros(0, np.uint8))
    np.savez_compressed(os.path.join(GOLD, "cc_label_stats.npz"), **out)

    # 3. temporal matching (the whole of stage 02) on small seeded videos
    runs = {"blobs_gap6": (list(synth.random_blob_masks(60, 96, 128, seed=3)), 0.85, 0.85, 6),
            "blobs_gap85": (list(synth.random_blob_masks(40, 72, 100, seed=4, jitter=0.01)), 0.85, 0.85, 85),
            "glyphs": (list(synth.glyph_masks(30, 180, 320, seed=1, churn=0.05, occluder_w=60, occluder_step=25)), 0.85, 0.85, 4),
            "loose": (list(synth.random_blob_masks(30, 64, 96, seed=9, jitter=0.05)), 0.5, 0.4, 3)}
    out = {}
    for name, (masks, r, p, gap) in runs.items():
        h, w = masks[0].shape
        est = CCStabilityEstimator(w, h, r, p, gap, False)
        for m in masks:
            est.add_frame(m, True)
        pf, uf, uq, tc = estimator_tables(est)
        out[name + "_masks"] = np.packbits(np.stack(masks) > 0, axis=-1)
        out[name + "_shape"] = np.array([len(masks), h, w]); out[name + "_params"] = np.array([r, p, gap], dtype=np.float64)
        out[name + "_per_frame"] = pf; out[name + "_uframes"] = uf; out[name + "_uniq"] = uq; out[name + "_tempo"] = tc
        print(name, "frames", len(masks), "raw ccs", len(pf), "uniques", len(uq), "tested", tc)
    np.savez_compressed(os.path.join(GOLD, "cc_stability.npz"), **out)


TINY_CONF = """
FCN_BINARIZER_NET_DOWN_CONV_FILTERS_1 = 16
FCN_BINARIZER_NET_DOWN_CONV_FILTERS_2 = 16
FCN_BINARIZER_NET_DOWN_CONV_FILTERS_3 = 32
FCN_BINARIZER_NET_DOWN_CONV_FILTERS_4 = 32
FCN_BINARIZER_NET_DOWN_CONV_FILTERS_5 = 48
FCN_BINARIZER_NET_MIDDLE_CONV_FILTERS_MIDDLE = 48
FCN_BINARIZER_NET_UPSAMPLE_FILTERS_5 = 32
FCN_BINARIZER_NET_UP_CONV_FILTERS_5 = 32
FCN_BINARIZER_NET_UPSAMPLE_FILTERS_4 = 16
FCN_BINARIZER_NET_UP_CONV_FILTERS_4 = 16
FCN_BINARIZER_NET_UPSAMPLE_FILTERS_3 = 16
FCN_BINARIZER_NET_UP_CONV_FILTERS_3 = 16
FCN_BINARIZER_NET_UPSAMPLE_FILTERS_2 = 16
FCN_BINARIZER_NET_UP_CONV_FILTERS_2 = 16
FCN_BINARIZER_NET_UPSAMPLE_FILTERS_1 = 32
FCN_BINARIZER_NET_UP_CONV_FILTERS_1 = 32
FCN_BINARIZER_NET_PIXEL_FEATURES_1 = 32
FCN_BINARIZER_NET_PIXEL_FEATURES_2 = 16
FCN_BINARIZER_NET_PIXEL_KERNEL_SIZE = 7
FCN_BINARIZER_NET_KERNEL_SIZE = 3
"""


def state_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode()); h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def gen_fcn(FCN_LectureNet, FCN_LectureNet_Binarizer, Configuration):
    import torch
    from lecturemath_b200 import synth
    torch.set_num_threads(8)
    conf_path = os.path.join(GOLD, "fcn_tiny.conf")
    with open(conf_path, "w") as f:
        f.write(TINY_CONF)
    frame = next(iter(synth.whiteboard_frames(1, 90, 112, seed=7)))
    out = {"frame_bgr": frame}
    for tag, path in (("tiny", conf_path), ("full", os.path.join(REF, "configs", "FCN_LectureNet.conf"))):
        cfg = Configuration.from_file(path)
        torch.manual_seed(0)
        net = FCN_LectureNet.CreateFromConfig(cfg, 3, False)
        # give BatchNorm non-trivial running stats / affine so that folding is really exercised
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for name, mod in net.named_modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
                    mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                    mod.weight.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                    mod.bias.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
        net.eval()
        sd = net.state_dict()
        if tag == "tiny":
            for k, v in sd.items():
                out["tiny_sd/" + k] = v.numpy()
        out[tag + "_sd_hash"] = state_hash(sd)
        worker = FCN_LectureNet_Binarizer(net)
        worker.initialize(112, 90)
        worker.handleFrame(frame, None, 0, 0.0, 0.0, 0)
        out[tag + "_binary"] = worker.last_binary
        out[tag + "_text"] = worker.last_text
        out[tag + "_rec"] = worker.last_rec
        import cv2
        from PIL import Image
        pil = Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_RGB2BGR))
        with torch.no_grad():
            logit, text, rec = net.forward(FCN_LectureNet.prepare_image(pil))
        out[tag + "_logit"] = logit[0, 0].numpy(); out[tag + "_text_logit"] = text[0, 0].numpy(); out[tag + "_rec_raw"] = rec[0].numpy()
        print(tag, "params", sum(v.numel() for v in sd.values()), "ink%", 100.0 * (worker.last_binary > 0).mean())
    np.savez_compressed(os.path.join(GOLD, "fcn_forward.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    Labeler, CCStabilityEstimator, FCN_LectureNet, FCN_LectureNet_Binarizer, Configuration = import_reference()
    which = sys.argv[1:] or ["cc", "fcn"]
    if "cc" in which:
        gen_cc(Labeler, CCStabilityEstimator)
    if "fcn" in which:
        gen_fcn(FCN_LectureNet, FCN_LectureNet_Binarizer, Configuration)
    print("golden fixtures written to", GOLD)
