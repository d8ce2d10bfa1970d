"""Host-side cost breakdown of the per-frame drop-in calls (handleFrame staging copy, PNG parse, add_frame staging) on this box."""
import concurrent.futures
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def t(fn, n=20):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "loadavg", os.getloadavg(), "torch threads", torch.get_num_threads())
    H, W = 1080, 1920
    frame = np.random.default_rng(0).integers(0, 255, (H, W, 3), dtype=np.uint8)
    pinned = torch.empty((8, H, W, 3), dtype=torch.uint8).pin_memory()
    pn = pinned.numpy()
    pageable = np.empty((H, W, 3), np.uint8)
    print("np.copyto -> pageable      %.3f ms" % t(lambda: np.copyto(pageable, frame)))
    print("np.copyto -> pinned        %.3f ms" % t(lambda: np.copyto(pn[0], frame)))
    print("torch copy_ -> pinned      %.3f ms" % t(lambda: pinned[0].copy_(torch.from_numpy(frame))))
    for parts in (2, 4, 8):
        pool = concurrent.futures.ThreadPoolExecutor(parts - 1)
        step = (H + parts - 1) // parts

        def par():
            futs = [pool.submit(np.copyto, pn[0][i * step:(i + 1) * step], frame[i * step:(i + 1) * step]) for i in range(1, parts)]
            np.copyto(pn[0][:step], frame[:step])
            for f in futs:
                f.result()
        print("parallel copy x%d -> pinned %.3f ms" % (parts, t(par)))
    dev = torch.empty((8, H, W, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def h2d_pageable():
        dev[0].copy_(torch.from_numpy(frame), non_blocking=True)
        torch.cuda.synchronize()
    print("H2D straight from pageable %.3f ms" % t(h2d_pageable))

    def h2d_pinned():
        dev[0].copy_(pinned[0], non_blocking=True)
        torch.cuda.synchronize()
    print("H2D from pinned            %.3f ms" % t(h2d_pinned))
    from oracle import png_oracle as PO
    from lecturemath_b200 import synth
    from lecturemath_b200.packed_mask import parse_png1
    import zlib
    m = (next(iter(synth.whiteboard_frames(1, H, W, seed=21, chalk=True, strokes_per_frame=40))).max(axis=2) > 128).astype(np.uint8) * 255
    png = np.frombuffer(PO.png1_deflate(m), np.uint8)
    print("png bytes", len(png))
    print("parse_png1 (deflate)       %.3f ms" % t(lambda: parse_png1(png)))
    b = png.tobytes()
    n = int.from_bytes(b[33:37], "big")
    print("  zlib.decompress only     %.3f ms" % t(lambda: zlib.decompress(b[41:41 + n])))
    pm = parse_png1(png)
    stage = torch.empty((8,) + pm.scan.shape, dtype=torch.uint8).pin_memory().numpy()

    def st():
        stage[0][...] = pm.scan
    print("add_frame staging copy     %.3f ms" % t(st))


if __name__ == "__main__":
    main()
