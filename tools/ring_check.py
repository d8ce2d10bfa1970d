"""Multi-rank parity check of the frame-sharded pipeline: N ranks must give the 1-rank answer.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ring_check.py
    (or start the ranks directly with RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT set, as tests/test_ring_gpu.py does)

Every rank runs its round-robin chunks through StreamingExtractor (hand-off of the active unique-CC set around the ring);
rank 0 also runs the WHOLE video alone (world = 1) and compares all per-frame result rows, the unique count and tempo_count:
they must be identical (the temporal matching stays one ordered scan).

  --same-device     every rank uses cuda:0 (two processes sharing ONE GPU: CUDA-IPC mailboxes and stream memory operations work
                    across processes on one device too, so the peer-memory ring is testable on a 1-GPU box); implies gloo
  --handoff p2p|nccl
  --masks glyph     inject dense-handwriting masks into the CC stage (the FCN still runs): thousands of uniques per hand-off"""
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
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--handoff", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--masks", default="fcn", choices=["fcn", "glyph"])
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    if args.same_device:
        local = 0
        if args.handoff == "nccl":
            raise SystemExit("NCCL refuses two ranks on one device; use --handoff p2p with --same-device")
    torch.cuda.set_device(local)
    if args.same_device:
        dist.init_process_group("gloo")                  # control plane only (exchange of the IPC handles, result gather)
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h, w = (int(v) for v in args.hw.split("x"))
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_engine import CCEngine
    from lecturemath_b200.configuration import Configuration
    from lecturemath_b200.fcn_lecturenet import FCN_LectureNet
    from lecturemath_b200.pipeline import StreamingExtractor
    conf = os.path.join(REPO, "tests", "golden", "fcn_full.conf" if args.full else "fcn_tiny.conf")
    torch.manual_seed(0)
    net = FCN_LectureNet.CreateFromConfig(Configuration.from_file(conf), 3, False).eval()
    b, n_chunks = args.batch, args.rounds * world
    dev = "cuda:%d" % local
    frames = np.stack(list(synth.whiteboard_frames(b * n_chunks, h, w, seed=7)))
    pinned = torch.from_numpy(frames).pin_memory()
    bits = None
    if args.masks == "glyph":
        masks = np.stack(list(synth.glyph_masks(b * n_chunks, h, w, seed=11)))
        bits = CCEngine(w, h, b * n_chunks, device=dev).pack(torch.from_numpy(masks).to(dev))
    inj = (lambda c: bits[c * b:(c + 1) * b].contiguous()) if bits is not None else (lambda c: None)
    sx = StreamingExtractor(net, w, h, 0.85, 0.85, 85, batch=b, rank=rank, world=world, device=dev, handoff=args.handoff)
    mine = {}
    lag = sx.lag
    for s in range(args.rounds):
        c = s * world + rank
        sx.submit(pinned[c * b:(c + 1) * b], last=(s == args.rounds - 1), inject_bits=inj(c))
        if s >= lag:
            mine[(s - lag) * world + rank] = sx.collect(s - lag)
    sx.flush()
    for s in range(max(0, args.rounds - lag), args.rounds):
        mine[s * world + rank] = sx.collect(s)
    state = sx.finish()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, state))
    ok = True
    if rank == 0:
        rows = {}
        for m, _ in gathered:
            rows.update(m)
        final = gathered[world - 1][1]                   # the last rank holds the final temporal state
        ref = StreamingExtractor(net, w, h, 0.85, 0.85, 85, batch=b, rank=0, world=1, device=dev)
        ref_rows = []
        for c in range(n_chunks):
            ref.submit(pinned[c * b:(c + 1) * b], last=(c == n_chunks - 1), inject_bits=inj(c))
            if c >= ref.lag:
                ref_rows += ref.collect(c - ref.lag)
        ref.flush()
        for c in range(max(0, n_chunks - ref.lag), n_chunks):
            ref_rows += ref.collect(c)
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
        print("ring_check world=%d handoff=%s%s masks=%s frames=%d rows=%d uniques=%d tempo_count=%d : %s" %
              (world, args.handoff, " (one device)" if args.same_device else "", args.masks, b * n_chunks, n_rows, ref_state["n_unique"],
               ref_state["tempo_count"], "IDENTICAL" if ok else "DIFFERENT"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
