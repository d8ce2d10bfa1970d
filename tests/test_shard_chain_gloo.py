"""Frame-shard chain on CPU: world_size-2 `gloo` run of the multi-GPU protocol (pipeline.run_chain + the active-set
wire format of pipeline.send_active_set / recv_active_set) with the oracle standing in for the CUDA estimator.
The 2-shard result must equal the single-process run (bit-exact hand-off of the ACTIVE set)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cc_oracle as O
from lecturemath_b200 import synth
from lecturemath_b200.connected_component import pack_crop, unpack_crop
from lecturemath_b200.pipeline import recv_active_set, run_chain, send_active_set, shard_ranges

W, H, N = 96, 64, 30
PARAMS = (0.85, 0.85, 6)


def export_oracle_state(est):
    """Oracle estimator -> (header, meta, crops) in the product's wire format."""
    meta, crops = [], []
    for u in est.cc_active:
        cc = est.unique_cc_objects[u]
        words = pack_crop(cc.img, cc.min_x, cc.max_x, cc.min_y, cc.max_y)
        first_frame, first_label = est.unique_cc_frames[u][0] if est.unique_cc_frames[u] else (-1, cc.cc_id + 1)
        meta.append([u, cc.min_x, cc.max_x, cc.min_y, cc.max_y, cc.size, est.cc_last_frame[u], first_frame, first_label, len(words)])
        crops.append(words)
    crops = np.concatenate(crops) if crops else np.zeros(0, np.uint32)
    header = torch.tensor([len(meta), len(crops), len(est.unique_cc_objects), est.img_idx, est.tempo_count, 0], dtype=torch.int64)
    meta_t = torch.tensor(meta, dtype=torch.int32).reshape(-1, 10) if meta else torch.zeros((1, 10), dtype=torch.int32)
    crops_t = torch.from_numpy(crops.view(np.int32).copy()) if len(crops) else torch.zeros(1, dtype=torch.int32)
    return header, meta_t, crops_t


def import_oracle_state(est, header, meta, crops):
    n_act, words, n_unique, img_idx, tempo = [int(v) for v in header[:5]]
    est.unique_cc_objects = [None] * n_unique
    est.unique_cc_frames = [[] for _ in range(n_unique)]
    est.cc_last_frame = [-(10 ** 9)] * n_unique
    est.cc_active = []
    cw = crops.numpy().view(np.uint32)
    off = 0
    for row in meta[:n_act].tolist():
        u, x0, x1, y0, y1, size, last, _, label, nw = row
        cc = O.OracleCC(label - 1, x0, x1, y0, y1, size, unpack_crop(cw[off:off + nw], x0, x1, y0, y1))
        off += nw
        est.unique_cc_objects[u] = cc
        est.cc_last_frame[u] = last
        est.cc_active.append(u)
    est.img_idx, est.tempo_count = img_idx, tempo


def _table(est, t0, t1, frame_offset):
    return [(t + frame_offset,) + row for t in range(t0, t1) for row in est.frame_table(t)]


def _worker(rank, world, port, masks, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_ranges(len(masks), world)[rank]
    est = O.StabilityOracle(W, H, *PARAMS)

    def recv_state(src):
        import_oracle_state(est, *recv_active_set(src))

    def send_state(dst):
        send_active_set(*export_oracle_state(est), dst)

    def match_shard():
        first = len(est.cc_idx_per_frame)                 # 0: cc_idx_per_frame only holds this shard's frames
        base = est.img_idx
        for m in masks[lo:hi]:
            est.add_frame(m)
        return _table(est, first, len(est.cc_idx_per_frame), base - first)

    rows = run_chain(rank, world, match_shard, recv_state, send_state)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rows, est.tempo_count, len(est.unique_cc_objects)))
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_shard_chain_equals_single_run():
    masks = np.stack(list(synth.random_blob_masks(N, H, W, seed=11)))
    ref = O.StabilityOracle(W, H, *PARAMS)
    for m in masks:
        ref.add_frame(m)
    ref_rows = _table(ref, 0, N, 0)
    assert len(ref.unique_cc_objects) > 10 and ref.tempo_count > 0
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world = 2
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, masks, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = [r for part in gathered for r in part[0]]
    assert rows == ref_rows
    assert gathered[-1][1] == ref.tempo_count and gathered[-1][2] == len(ref.unique_cc_objects)


def test_shard_ranges_cover_all_frames():
    for n, g in [(10, 3), (108000, 8), (7, 8), (0, 2)]:
        r = shard_ranges(n, g)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def test_crop_pack_roundtrip():
    rng = np.random.default_rng(0)
    for x0, x1, y0, y1 in [(0, 0, 0, 0), (31, 32, 5, 9), (5, 70, 0, 3), (64, 95, 2, 2)]:
        img = (rng.random((y1 - y0 + 1, x1 - x0 + 1)) < 0.5).astype(np.uint8) * 255
        np.testing.assert_array_equal(unpack_crop(pack_crop(img, x0, x1, y0, y1), x0, x1, y0, y1), img)
