"""Where the batching worker's wall time goes: host staging copy / submit / collect (blocking) per batch.  python tools/worker_profile.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from bench import make_net, frame_pool
    from lecturemath_b200 import fcn_binarizer_worker as FW
    net = make_net().cuda()
    pool = frame_pool(16, 1234)
    acc = {"copy": 0.0, "submit": 0.0, "collect": 0.0}
    P = FW._BatchPipe
    orig_submit, orig_collect = P.submit, P.collect

    def submit(self, worker, k):
        t0 = time.perf_counter(); orig_submit(self, worker, k); acc["submit"] += time.perf_counter() - t0

    def collect(self, worker, k):
        t0 = time.perf_counter(); orig_collect(self, worker, k); acc["collect"] += time.perf_counter() - t0
    P.submit, P.collect = submit, collect
    from lecturemath_b200 import wire
    evs = []
    orig_launch = wire.PngEncoder.launch

    def launch(self, bits, n, stream, copy_stream=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        orig_launch(self, bits, n, stream, copy_stream)
        e1.record(stream)
        evs.append((e0, e1))
    wire.PngEncoder.launch = launch
    from lecturemath_b200 import fcn_lecturenet as FL
    fevs = []
    orig_run = FL.FCNPlan.run

    def run(self, stream, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_run(self, stream, *a, **k)
        e1.record()
        fevs.append((e0, e1))
    FL.FCNPlan.run = run
    for rep in range(2):
        for k in acc:
            acc[k] = 0.0
        w = FW.FCN_LectureNet_Binarizer(net, batch=8)
        w.initialize(1920, 1080)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(160):
            w.handleFrame(pool[i % 16], None, 0, 0.0, 0.0, i)
            if i == 0:
                w._pipe.debug_timing = []
        w.finalize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dbg = w._pipe.debug_timing
        print("H2D of a batch: %.3f ms; compute stream stalled on it: %.3f ms; H2D start relative to the stream's arrival at the wait: %.3f ms" %
              (np.mean([t[0].elapsed_time(t[1]) for t in dbg]), np.mean([t[2].elapsed_time(t[3]) for t in dbg]),
               np.mean([t[2].elapsed_time(t[0]) for t in dbg])))
        w._pipe.debug_timing = None
        print("png encode on the compute stream: %.3f ms per batch" % (sum(a.elapsed_time(b) for a, b in evs) / max(len(evs), 1)))
        evs.clear()
        print("FCN plan.run on the compute stream: %.3f ms per batch; gaps between consecutive runs: %.3f ms" %
              (sum(a.elapsed_time(b) for a, b in fevs) / max(len(fevs), 1),
               sum(fevs[i][1].elapsed_time(fevs[i + 1][0]) for i in range(len(fevs) - 1)) / max(len(fevs) - 1, 1)))
        fevs.clear()
        print("rep %d: %.1f fps, total %.1f ms; submit %.1f ms, collect(blocking) %.1f ms, rest (staging copies + python) %.1f ms" %
              (rep, 160 / dt, dt * 1e3, acc["submit"] * 1e3, acc["collect"] * 1e3, (dt - acc["submit"] - acc["collect"]) * 1e3))
    # the raw staging copy, cold memory
    pin = torch.empty((16, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    for i in range(160):
        pin[i % 16].copy_(torch.from_numpy(pool[(i * 7) % 16]))
    print("staging copy alone: %.3f ms per frame" % ((time.perf_counter() - t0) / 160 * 1e3))


if __name__ == "__main__":
    main()
