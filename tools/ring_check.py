"""Multi-GPU parity check of the frame-sharded pipeline (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ring_check.py

Every rank runs its round-robin chunks through StreamingExtractor (NCCL ring hand-off of the active unique-CC set); rank 0
also runs the WHOLE video alone (world = 1) and compares all per-frame result rows, the unique count and tempo_count:
they must be identical (the temporal matching stays one ordered scan)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", default="360x640")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--full", action="store_true", help="full-size network (FCN_LectureNet.conf widths) instead of the tiny test config")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h, w = (int(v) for v in args.hw.split("x"))
    from lecturemath_b200 import synth
    from lecturemath_b200.configuration import Configuration
    from lecturemath_b200.fcn_lecturenet import FCN_LectureNet
    from lecturemath_b200.pipeline import StreamingExtractor
    conf = os.path.join(REPO, "tests", "golden", "fcn_full.conf" if args.full else "fcn_tiny.conf")
    torch.manual_seed(0)
    net = FCN_LectureNet.CreateFromConfig(Configuration.from_file(conf), 3, False).eval()
    b, n_chunks = args.batch, args.rounds * world
    frames = np.stack(list(synth.whiteboard_frames(b * n_chunks, h, w, seed=7)))
    pinned = torch.from_numpy(frames).pin_memory()
    sx = StreamingExtractor(net, w, h, 0.85, 0.85, 85, batch=b, rank=rank, world=world, device="cuda:%d" % local)
    mine = {}
    for s in range(args.rounds):
        c = s * world + rank
        sx.submit(pinned[c * b:(c + 1) * b], last=(s == args.rounds - 1))
        if s >= 1:
            mine[(s - 1) * world + rank] = sx.collect(s - 1)
    mine[(args.rounds - 1) * world + rank] = sx.collect(args.rounds - 1)
    state = sx.finish()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, state))
    ok = True
    if rank == 0:
        rows = {}
        for m, _ in gathered:
            rows.update(m)
        final = gathered[world - 1][1]                   # the last rank holds the final temporal state
        ref = StreamingExtractor(net, w, h, 0.85, 0.85, 85, batch=b, rank=0, world=1, device="cuda:0")
        ref_rows = []
        for c in range(n_chunks):
            ref.submit(pinned[c * b:(c + 1) * b], last=(c == n_chunks - 1))
            if c >= 1:
                ref_rows += ref.collect(c - 1)
        ref_rows += ref.collect(n_chunks - 1)
        ref_state = ref.finish()
        n_rows = 0
        for c in range(n_chunks):
            for f in range(b):
                a, r = rows[c][f], ref_rows[c * b + f]
                n_rows += len(r)
                if a.shape != r.shape or not np.array_equal(a, r):
                    ok = False
                    print("MISMATCH chunk %d frame %d: %s vs %s" % (c, f, a.shape, r.shape))
        for key in ("n_unique", "img_idx", "tempo_count"):
            if final[key] != ref_state[key]:
                ok = False
                print("MISMATCH state", key, final[key], ref_state[key])
        print("ring_check world=%d frames=%d rows=%d uniques=%d tempo_count=%d : %s" %
              (world, b * n_chunks, n_rows, ref_state["n_unique"], ref_state["tempo_count"], "IDENTICAL" if ok else "DIFFERENT"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
