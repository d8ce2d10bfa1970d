"""Fused per-frame content extraction on one B200, and its frame-sharded multi-GPU form.

ContentExtractor = stage 01 + stage 02 of the reference without the PNG/pickle hand-off
(R/pre_ST3D_v3.0_01_binarize.py:49-55 -> R/pre_ST3D_v3.0_02_cc_analaysis.py:19-43):
    uint8 BGR frames --H2D--> FCN binarizer (tcgen05) --> bit-packed ink mask --> CC label/stats/crops
    --> temporal matching --D2H--> per-frame rows (unique_idx, raw_label, min_x, max_x, min_y, max_y, size)

Multi-GPU: contiguous frame ranges per rank, no data-path collective in phase 1 (FCN + masks);
temporal matching is a sequential scan whose carried state is the ACTIVE unique-CC set, so ranks form a chain:
rank r receives (active set, next unique index, img_idx, tempo_count) from r-1, matches its shard, sends to r+1.
"""
import numpy as np
import torch

from .cc_engine import CCEngine, Estimator


class ContentExtractor:
    def __init__(self, net, width, height, min_recall=0.85, min_precision=0.85, max_gap=85, batch=8, device=None,
                 max_uniques=0, max_active=0, arena_words=0):
        self.net, self.width, self.height, self.batch = net, width, height, batch
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        net.cuda(self.device.index)
        self.plan = net.plan(batch, height, width)
        self.engine = CCEngine(width, height, batch, device=self.device)
        self.est = Estimator(width, height, min_recall, min_precision, max_gap, max_uniques, max_active, arena_words, device=self.device)
        self.pinned_in = torch.empty((batch, height, width, 3), dtype=torch.uint8).pin_memory()
        self.launches = 0

    # ---- device-resident step (inputs already in HBM) ------------------------------------------------
    def step_device(self, frames_dev=None, match=True, timing=None):
        """One pass of the hot path over one batch.  Returns (rows, offsets) device tensors when match=True."""
        plan, eng = self.plan, self.engine
        if frames_dev is not None:
            plan.frames.copy_(frames_dev, non_blocking=True)
        plan.run(torch.cuda.current_stream().cuda_stream, False, 128, timing)
        self.launches += plan.launches_per_run
        eng.label(plan.bits, want_labels=False, sync=False)
        self.launches += 12
        if not match:
            return None
        self.est.add_frames(eng, 0, self.batch)
        self.launches += 4 * self.batch
        return None

    def read_rows(self):
        """D2H of the batch's result rows: list (per frame) of int32 arrays [n_cc][7]."""
        eng = self.engine
        eng.read_counts()
        rows, offs = eng.packed_rows(self.batch)
        self.launches += 2
        rows_h, offs_h = rows.cpu().numpy(), offs.cpu().numpy()
        return [rows_h[offs_h[f]:offs_h[f + 1], :7] for f in range(self.batch)]

    # ---- public end-to-end call (host buffers) -----------------------------------------------------------
    def process_batch(self, frames_host):
        """frames_host: uint8 (batch, H, W, 3) BGR numpy array / CPU tensor -> per-frame result rows (host)."""
        t = torch.from_numpy(frames_host) if isinstance(frames_host, np.ndarray) else frames_host
        if not t.is_pinned():                                            # pageable input: stage through pinned memory
            self.pinned_in.copy_(t)
            t = self.pinned_in
        self.plan.frames.copy_(t, non_blocking=True)
        self.step_device(None, match=True)
        return self.read_rows()

    def masks_host(self):
        """uint8 (batch, H, W) ink masks (255 = ink) of the last batch, in the reference's format."""
        return self.engine.unpack(self.plan.bits).cpu().numpy()

    # ---- frame-shard chain ------------------------------------------------------------------------------
    def recv_state(self, src):
        h, meta, crops = recv_active_set(src, self.device)
        self.est.import_state(h, meta, crops)

    def send_state(self, dst):
        header, meta, crops = self.est.export_state()
        send_active_set(header, meta, crops, dst, self.device)


# Wire format of the hand-off between frame shards (the only data-path exchange):
#   header int64[6] = n_active, crop_words, n_unique, img_idx, tempo_count, 0
#   meta   int32[n_active][10] = unique_idx, min_x, max_x, min_y, max_y, size, last_seen, first_frame, first_label, crop_words
#   crops  int32[crop_words]   = the first-seen bit-packed crops of the active uniques, concatenated in meta order
def send_active_set(header, meta, crops, dst, device=None):
    """Point-to-point send of the active unique-CC set (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cpu")
    dist.send(header.to(dev), dst=dst)
    dist.send(meta.to(dev).contiguous(), dst=dst)
    dist.send(crops.to(dev).contiguous(), dst=dst)


def recv_active_set(src, device=None):
    import torch.distributed as dist
    dev = device if device is not None else torch.device("cpu")
    header = torch.zeros(6, dtype=torch.int64, device=dev)
    dist.recv(header, src=src)
    h = header.cpu()
    n_act, words = int(h[0]), int(h[1])
    meta = torch.zeros((max(n_act, 1), 10), dtype=torch.int32, device=dev)
    crops = torch.zeros((max(words, 1),), dtype=torch.int32, device=dev)
    dist.recv(meta, src=src)
    dist.recv(crops, src=src)
    return h, meta, crops


def shard_ranges(n_frames, world):
    """Contiguous frame ranges [r*F/G, (r+1)*F/G) per rank."""
    return [(r * n_frames // world, (r + 1) * n_frames // world) for r in range(world)]


def run_chain(rank, world, match_shard, recv_state, send_state):
    """Phase-2 protocol: rank r waits for the active set of rank r-1, matches its own frames, hands the set on.
    The three callables are injected so that the protocol is testable on CPU (gloo) with the oracle as engine."""
    if rank > 0:
        recv_state(rank - 1)
    out = match_shard()
    if rank + 1 < world:
        send_state(rank + 1)
    return out
