"""CC stage alone (BASELINE configs[3]): label + stats + crops (and temporal matching) on dense-glyph 1080p masks resident
in HBM, timed with CUDA events on the launching stream, L2 flushed between iterations.  One JSON line per batch size.

    python tools/cc_bench.py [--batches 32,148,296] [--iters 5] [--hw 1080x1920] [--labels] [--no-match] [--max-labels 65536]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="32,148,296")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--hw", default="1080x1920")
    ap.add_argument("--labels", action="store_true", help="also materialise the int32 label image (the SciPy operator boundary)")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--max-labels", type=int, default=65536)
    ap.add_argument("--pool", type=int, default=32, help="distinct masks generated on the CPU (cycled to fill the batch)")
    args = ap.parse_args()
    h, w = (int(v) for v in args.hw.split("x"))
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_engine import CCEngine, Estimator
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    peak = 6547.5
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    pool = np.stack(list(synth.glyph_masks(args.pool, h, w, seed=0)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for batch in [int(b) for b in args.batches.split(",")]:
        masks = pool[np.arange(batch) % args.pool]
        eng = CCEngine(w, h, batch, max_labels=args.max_labels, max_kept=args.max_labels, device=dev)
        bits = eng.pack(torch.from_numpy(masks).to(dev))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_label = t_match = 0.0
        for it in range(args.iters + 2):
            est = None if args.no_match else Estimator(w, h, 0.85, 0.85, 85, device=dev)
            flush.fill_(it)
            ev[0].record()
            eng.label(bits, want_labels=args.labels, sync=False)
            ev[1].record()
            if est is not None:
                est.add_frames(eng, 0, min(batch, 32))
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                t_label += ev[0].elapsed_time(ev[1])
                t_match += ev[1].elapsed_time(ev[2])
        counts = eng.read_counts()
        n_labels = float(counts[:, 1].mean())
        P = h * w
        canon, strict = 12.125 * P + 24 * n_labels, 4.125 * P + 24 * n_labels
        fps = batch * args.iters / (t_label / 1000.0)
        out = {"batch": batch, "label_image": bool(args.labels), "us_per_frame": 1e6 / fps, "label_frames_per_s": fps,
               "canonical_GBps": fps * canon / 1e9, "frac_canonical": fps * canon / 1e9 / peak,
               "strict_floor_GBps": fps * strict / 1e9, "frac_strict_floor": fps * strict / 1e9 / peak,
               "runs_per_frame": float(counts[:, 0].mean()), "labels_per_frame": n_labels, "ccs_per_frame": float(counts[:, 2].mean())}
        if est is not None:
            out["match_frames_per_s"] = min(batch, 32) * args.iters / (t_match / 1000.0)
            out["tempo_count"] = est.state()["tempo_count"]
        print(json.dumps(out), flush=True)
        eng.close()
        del eng, bits


if __name__ == "__main__":
    main()
