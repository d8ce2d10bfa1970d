#!/usr/bin/env python
"""bench.py -- headline benchmark: 1080p frames/s for (FCN binarize + CC label/stats + temporal match).

  python bench.py --gpus N --steps K --warmup W            this repo on N B200s (torchrun launches N>1)
  python bench.py --impl reference --steps K --warmup W     the reference's CPU algorithm (oracle port) on host cores

A "step" = one pass of the hot path over one batch of synthetic 1080p whiteboard frames (BASELINE.json
configs[1]).  `value` = whole-job frames/s with the frames already resident in HBM; `e2e` = the same metric
through StreamingExtractor.submit/collect with HOST buffers (H2D of the frames and D2H of the result rows inside the
timed region); `e2e_dropin` = the same through the reference's own per-frame surface (FCN_LectureNet_Binarizer.handleFrame ->
CCStabilityEstimator.add_frame).  Multi-GPU: frame chunks dealt round-robin to the ranks (weak scaling), temporal matching
is ONE ordered scan whose active unique-CC set travels rank to rank over NVLink peer memory (lecturemath_b200/pipeline.py);
`ring_parity` = the N-rank result rows / uniques / tempo_count compared with a 1-rank replay of the same chunks.
`gpu_reference` = the reference's production arm (lecture_net.cuda(): torch/cuDNN) timed on the same GPU in the same run.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

H, W = 1080, 1920
CONF = os.path.join(REPO, "tests", "golden", "fcn_full.conf")
METRIC = "1080p frames/s (FCN binarize + CC label/stats/match)"


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML every 5 ms when pynvml is
    importable (a 10-step run lasts ~100 ms, one nvidia-smi call about as long), else nvidia-smi every 0.2 s."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False     # samples: [sm_mhz, sm_max_mhz, hw, hw_thermal, sw_thermal, sw_power]
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.phys = phys
        except Exception:
            self.nvml, self.phys = None, index

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown, n.nvmlClocksEventReasonSwThermalSlowdown,
                n.nvmlClocksEventReasonSwPowerCap]
        return [str(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), str(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))] + \
               ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.005)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.phys), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                if self.nvml is not None:
                    self.nvml = None                   # fall back to nvidia-smi
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_net():
    import torch
    from lecturemath_b200.configuration import Configuration
    from lecturemath_b200.fcn_lecturenet import FCN_LectureNet
    torch.manual_seed(0)                                  # the reference's random init under seed 0 (no trained weights on the box)
    return FCN_LectureNet.CreateFromConfig(Configuration.from_file(CONF), 3, False).eval()


CHALK = False


def frame_pool(n, seed):
    from lecturemath_b200 import synth
    return np.stack(list(synth.whiteboard_frames(n, H, W, seed=seed, chalk=CHALK)))


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_pass(frames, sd, est):
    """The reference's CPU algorithm (oracle port) for a few frames: torch-CPU fp32 FCN + oracle CC stage."""
    from oracle import fcn_oracle as FO
    t0 = time.perf_counter()
    for fr in frames:
        ink, _, _ = FO.handle_frame(sd, fr)
        est.add_frame(ink)
    return time.perf_counter() - t0


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cc_oracle as CO
    torch.set_num_threads(os.cpu_count())
    net = make_net()
    sd = net.state_dict()
    n_timed = max(1, min(args.steps, 24))
    n_warm = max(1, min(args.warmup, 2))
    frames = frame_pool(n_timed + n_warm, 1234)
    est = CO.StabilityOracle(W, H, 0.85, 0.85, 85, use_ref_lib=CO.ref_lib() is not None)
    cpu_reference_pass(frames[:n_warm], sd, est)
    dt = cpu_reference_pass(frames[n_warm:], sd, est)
    fps = n_timed / dt
    kind = "port"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / n_timed, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "1080p synthetic whiteboard video, binarize + CC label/stats + temporal match (BASELINE configs[1]); "
                                   "one frame per step", "weights": "random-init seed 0"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": "%d timed 1080p frames (torch-CPU fp32 FCN + scipy-equivalent label + %s CC_AgeBoundaries + "
                                       "Python matching), %d warm-up" % (n_timed, "reference-compiled" if CO.ref_lib() is not None else "ported", n_warm)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def gpu_reference(dev, net, frames_u8, flops_per_frame, iters=3):
    """The reference's PRODUCTION arm for the FCN: `lecture_net.cuda()` + forward under no_grad (R/pre_ST3D_v3.0_01_binarize.py:35-37,
    R/AccessMath/lecturenet_v1/FCN_lecturenet.py:442-459) = whatever torch/cuDNN dispatches for nn.Conv2d / BatchNorm2d / GELU on
    this GPU, run through the fp32 oracle restatement of the reference's layers (oracle/fcn_oracle.py, a baseline leg like
    cpu_baseline).  Device-resident forward only (the reference also pays 66 MB of fp32 H2D/D2H per frame, not charged here).
    Variants: batch 1 fp32 exactly as the reference drives it, batch 8 fp32, batch 8 bf16 channels_last (what a maintainer
    would try first); cudnn.benchmark on so cuDNN picks its best kernels."""
    import torch
    from oracle import fcn_oracle as FO
    out = {"what": "oracle/fcn_oracle.forward on cuda (torch %s / cuDNN %s), frames resident, cudnn.benchmark=True" %
                   (torch.__version__, torch.backends.cudnn.version()),
           "allow_tf32": {"cudnn": bool(torch.backends.cudnn.allow_tf32), "matmul": bool(torch.backends.cuda.matmul.allow_tf32)},
           "variants": {}}
    old_bench = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    sd32 = {k: v.detach().to(dev) for k, v in net.state_dict().items()}
    x_all = ((frames_u8.to(dev).flip(-1).permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5      # BGR -> RGB, prepare_image
    try:
        for name, b, dt, cl in (("fp32_batch1", 1, torch.float32, False), ("fp32_batch8", 8, torch.float32, False),
                                ("bf16_channels_last_batch8", 8, torch.bfloat16, True)):
            b = min(b, x_all.shape[0])
            sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd32.items()}
            x = x_all[:b].to(dt).contiguous()
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
                sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
            for _ in range(2):
                FO.forward(sd, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                FO.forward(sd, x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (iters * b)
            out["variants"][name] = {"frames_per_s": 1000.0 / ms, "ms_per_frame": ms, "tflops": flops_per_frame / (ms / 1000.0) / 1e12}
            del sd, x
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = old_bench
    best = max(out["variants"].values(), key=lambda v: v["frames_per_s"])
    out["best_frames_per_s"] = best["frames_per_s"]
    return out


def dropin_e2e(dev, net, pool_np, n_frames, batch):
    """The same metric through the REFERENCE'S OWN per-frame calls with host buffers, as its two stage scripts make them
    (R/pre_ST3D_v3.0_01_binarize.py:31-55, R/pre_ST3D_v3.0_02_cc_analaysis.py:19-41):
        worker.initialize; worker.handleFrame(frame, ...) per frame; worker.finalize()              -> compressed_frames (PNG per frame)
        Helper.decompress_binary_images(compressed_frames); estimator.add_frame(mask, True) per frame; estimator.finish_processing()
    `value` = frames / (stage 01 + stage 02 wall time), the two stages run one after the other like the reference's processes;
    `fused` = handleFrame per frame with the estimator attached to the worker (masks stay on the device, PNGs still written).
    Wall clock around device-synchronised phases: the host work (frame memcpy, PNG parse) is part of what is measured."""
    import torch
    from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    from lecturemath_b200.helper import Helper
    frames = [pool_np[i % len(pool_np)] for i in range(n_frames)]

    def stage01(worker):
        worker.initialize(W, H)
        for i, fr in enumerate(frames):
            worker.handleFrame(fr, None, 0, 33.3 * i, 33.3 * i, i)
        worker.finalize()
        torch.cuda.synchronize(dev)

    def stage02(est, masks):
        for m in masks:
            est.add_frame(m, True)
        est.flush()
        torch.cuda.synchronize(dev)

    out = {}
    for rep in range(2):                                  # first repetition = warm-up (plans, pinned buffers, PNG tables)
        # objects are built outside the timed regions (a stage script builds them once per video, not per 80-frame sample)
        worker = FCN_LectureNet_Binarizer(net, batch=batch)
        est = CCStabilityEstimator(W, H, 0.85, 0.85, 85, max_batch=batch)
        fused_est = CCStabilityEstimator(W, H, 0.85, 0.85, 85, max_batch=batch)
        fused = FCN_LectureNet_Binarizer(net, batch=batch, estimator=fused_est)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        stage01(worker)
        t1 = time.perf_counter()
        masks = Helper.decompress_binary_images(worker.compressed_frames)
        stage02(est, masks)
        t2 = time.perf_counter()
        t3 = time.perf_counter()
        stage01(fused)
        t4 = time.perf_counter()
        out = {"value": n_frames / (t2 - t0), "unit": "frames/s", "frames": n_frames, "frames_per_gpu_step": batch,
               "stage01_frames_per_s": n_frames / (t1 - t0), "stage02_frames_per_s": n_frames / (t2 - t1),
               "fused": {"value": n_frames / (t4 - t3), "unit": "frames/s",
                         "what": "handleFrame per frame, estimator attached to the worker (stage 01 + 02 in one pass)"},
               "png_kb_per_frame": float(np.mean([len(f) for f in worker.compressed_frames])) / 1e3,
               "h2d_bytes_per_frame": H * W * 3, "identical_state": bool(
                   est.tempo_count == fused_est.tempo_count and est.get_raw_cc_count() == fused_est.get_raw_cc_count()),
               "what": "FCN_LectureNet_Binarizer.handleFrame -> finalize -> Helper.decompress_binary_images -> CCStabilityEstimator.add_frame "
                       "-> flush, one call per 1080p frame, host numpy frames in, PNG bytes + estimator state out"}
        del worker, est, fused, fused_est, masks
    return out


def conv_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv launches of one step, from the committed ncu capture
    (profiles/conv_traffic.json, written by tools/summarize_profile.py); None when no capture of this build exists."""
    if args.traffic is not None:
        return args.traffic
    p = os.path.join(REPO, "profiles", "conv_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_step")
    return None


def cc_stage_roofline(dev, peaks, batch=148, match_frames=32, iters=5):
    """Secondary roofline: the CC stage alone on dense-handwriting masks (BASELINE configs[3]: >5k CCs per 1080p frame),
    `batch` frames per launch sequence (k_strip_label, k_resolve, k_crop_fill), masks resident in HBM, L2 flushed between
    iterations.  `achieved` = canonical operator-boundary bytes (12.125*P + 24*n per frame) / label-stage
    time; the fused path never materialises the int32 label image, so its real DRAM traffic is far below that figure --
    the strict single-pass floor, the run WITH the label image written (the reference's operator boundary) and the
    ncu-measured DRAM bytes are all reported next to it."""
    import torch
    from lecturemath_b200 import synth
    from lecturemath_b200.cc_engine import CCEngine, Estimator
    pool = np.stack(list(synth.glyph_masks(32, H, W, seed=0)))
    masks = pool[np.arange(batch) % 32]
    eng = CCEngine(W, H, batch, max_labels=65536, max_kept=65536, device=dev)
    bits = eng.pack(torch.from_numpy(masks).to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_label = t_match = t_full = 0.0
    labels = torch.empty((batch, H, W), dtype=torch.int32, device=dev)
    for it in range(iters + 2):
        est = Estimator(W, H, 0.85, 0.85, 85, device=dev)
        flush.fill_(it)
        ev[0].record()
        eng.label(bits, want_labels=False, sync=False)
        ev[1].record()
        est.add_frames(eng, 0, match_frames)
        ev[2].record()
        flush.fill_(it + 1)
        ev[3].record()
        eng.label(bits, want_labels=True, sync=False, out=labels)
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        torch.cuda.synchronize()
        if it >= 2:
            t_label += ev[0].elapsed_time(ev[1])
            t_match += ev[1].elapsed_time(ev[2])
            t_full += ev[3].elapsed_time(end)
    counts = eng.read_counts()
    st = est.state()
    n_labels = float(counts[:, 1].mean())
    bytes_frame, floor_frame = 12.125 * H * W + 24 * n_labels, 4.125 * H * W + 24 * n_labels
    fps_label = batch * iters / (t_label / 1000.0)
    fps_full = batch * iters / (t_full / 1000.0)
    achieved = fps_label * bytes_frame / 1e9
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    p = os.path.join(REPO, "profiles", "cc_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            traffic = json.load(f).get("dram_bytes_per_frame")
    # headline fraction = the strict single-pass floor (mask in once, labels out once, tables); the canonical operator-boundary
    # figure of SURVEY.md 8d (25 MB/frame, which the fused kernels never move) and the ncu-measured DRAM bytes are named extras
    floor_rate = fps_label * floor_frame / 1e9
    return {"bound": "hbm (in practice shared-memory / latency bound: see traffic_per_frame)",
            "kernel": "CC label+stats+crops (k_strip_label, k_resolve, k_crop_fill: 3 launches per %d-frame batch), dense glyph masks" % batch,
            "achieved": floor_rate, "peak": peak, "unit": "GB/s", "frac": floor_rate / peak,
            "bytes_per_frame": floor_frame, "bytes_definition": "strict single-pass floor 4.125*P + 24*n",
            "canonical_operator_boundary": {"bytes_per_frame": bytes_frame, "achieved": achieved, "frac": achieved / peak,
                                            "note": "12.125*P + 24*n (SURVEY.md 8d): counts the int32 label image three times; the fused path never writes it"},
            "real_dram": {"bytes_per_frame": traffic, "achieved": (fps_label * traffic / 1e9) if traffic else None,
                          "frac": (fps_label * traffic / 1e9 / peak) if traffic else None, "source": "ncu dram__bytes, profiles/cc_traffic.json"},
            "label_frames_per_s": fps_label, "us_per_frame": 1e6 / fps_label,
            "with_label_image": {"frames_per_s": fps_full, "achieved": fps_full * floor_frame / 1e9, "unit": "GB/s",
                                 "frac": fps_full * floor_frame / 1e9 / peak,
                                 "note": "int32 label image written: real traffic ~= the strict floor (write-dominated)"},
            "traffic_per_frame": traffic, "match_frames_per_s": match_frames * iters / (t_match / 1000.0),
            "ccs_per_frame": float(counts[:, 2].mean()), "labels_per_frame": n_labels, "tempo_count": st["tempo_count"]}


# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from lecturemath_b200.pipeline import StreamingExtractor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = None
    if world > 1 and not args.no_numa_bind:
        from lecturemath_b200.pipeline import bind_host_to_gpu
        numa_cores = bind_host_to_gpu(local)              # before any pinned allocation (first touch)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner (printed at the VERSION and WARN levels) and warnings go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    net = make_net()
    sx = StreamingExtractor(net, W, H, 0.85, 0.85, 85, batch=B, rank=rank, world=world, device=dev)
    # L2 between timed iterations: the input pool is LARGER than the 126 MB L2 and cycled (32 x 1080p frames = 199 MB; each step also
    # streams ~19 GB of activations through it), so no step finds its inputs cached; --l2-flush adds the 256 MB flush write as well
    frame_bytes = H * W * 3
    pool_n = max(4 * B, B * (-(-(160 << 20) // (B * frame_bytes))))
    # ONE synthetic video for the whole job (every rank builds the same seeded pool): global chunk c = step * world + rank reads pool
    # chunk c mod (pool / B), i.e. the ranks process consecutive chunks of the same cyclic video, as a sharded lecture would be
    pool_h = torch.from_numpy(frame_pool(pool_n, 1234)).pin_memory()
    pool_d = pool_h.to(dev)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.l2_flush else None
    main = torch.cuda.current_stream(dev)
    # --masks glyph: the FCN runs on the frames as always, but the CC stage gets dense-handwriting masks (> 4.5k CCs per frame,
    # BASELINE configs[3]) -- random-init weights never produce such masks, and they are what loads the labeling / matching
    # kernels and the ring hand-off (thousands of uniques + crops per chunk)
    glyph_bits = None
    if args.masks == "glyph":                             # ONE mask video for the whole ring: global chunk c reads pool chunk c mod (pool / B)
        from lecturemath_b200 import synth
        from lecturemath_b200.cc_engine import CCEngine
        m = np.stack(list(synth.glyph_masks(pool_n, H, W, seed=100)))
        glyph_bits = CCEngine(W, H, pool_n, device=dev).pack(torch.from_numpy(m).to(dev))
        del m

    def inject_of(r, i):
        if glyph_bits is None:
            return None
        s0 = ((i * world + r) % (pool_n // B)) * B
        return glyph_bits[s0:s0 + B]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def batch_of(pool, i, r=None):                        # global chunk i * world + r of the (pinned host | device) frame pool
        c = i * world + (rank if r is None else r)
        s = (c % (pool_n // B)) * B
        return pool[s:s + B]

    def run_video(n_steps, host_io, timing=None, keep=None, ex=None, chunk_of=None):
        """n_steps batches through the hot path (and the rank ring); returns (d2h bytes, CC rows seen).  Results of batch i are
        read back `lag` submits later (1 alone; 2 on a ring, whose ranks match chunk i-1 behind the FCN of chunk i).
        keep: dict chunk -> rows (ring-parity check); ex / chunk_of: another extractor / chunk numbering (the 1-rank replay)."""
        ex = ex or sx
        n_cc = 0
        d2h0 = ex.d2h_bytes
        lag = ex.lag

        def take(j):
            rows = ex.collect(j)
            if keep is not None:
                keep[j if chunk_of else j * world + rank] = rows
            return sum(len(r) for r in rows)

        for i in range(n_steps):
            if l2_flush is not None:
                l2_flush.fill_(i & 0xff)
            frames, inj = chunk_of(i) if chunk_of else (batch_of(pool_h if host_io else pool_d, i), inject_of(rank, i))
            ex.submit(frames, last=(i == n_steps - 1), timing=timing, inject_bits=inj)
            if host_io and i >= lag:
                n_cc += take(i - lag)
        if host_io:
            ex.flush()
            for j in range(max(0, n_steps - lag), n_steps):
                n_cc += take(j)
        return ex.d2h_bytes - d2h0, n_cc                   # bytes collect() actually copied device -> host

    # warm-up (also initialises the NCCL ring)
    run_video(Wm, False)
    sx.reset()                                            # every phase is a new video: fresh temporal state, outside the timed region
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---------------- device-resident throughput: K batches already in HBM ------------------------------------
    timing = []
    sx.launches = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_video(K, False, timing)
    e1.record()
    barrier()
    ms_own = e0.elapsed_time(e1)
    ms = torch.tensor([ms_own], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = sx.launches
    state = sx.finish()
    conv_ms = sum(a.elapsed_time(b) for _, a, b in timing)
    per_rank = None
    if world > 1:           # diagnostics: every rank's own timed region and conv time (the step time above is the maximum over the ranks)
        mine = torch.tensor([ms_own / max(K, 1), conv_ms / max(K, 1)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(t[0].item()), 3) for t in allr], "conv_ms_per_step": [round(float(t[1].item()), 3) for t in allr]}
    per_op = {}
    for i, a, b in timing:
        per_op.setdefault(i, []).append(a.elapsed_time(b))

    # ---------------- end to end through the public API with HOST buffers -------------------------------------
    sx.reset()
    run_video(2, True)
    sx.reset()
    barrier()
    want_parity = world > 1 and not args.no_ring_parity and world * K <= 512
    kept = {} if want_parity else None
    e0.record()
    d2h, n_cc = run_video(K, True, keep=kept)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    state_e2e = sx.finish()
    fin = torch.tensor([state_e2e["n_unique"], state_e2e["tempo_count"]], dtype=torch.int64, device=dev)
    if world > 1:                                         # the final temporal state lives on the last rank of the ring
        dist.broadcast(fin, src=world - 1)
    masks = sx.masks_host()
    ink_pct = 100.0 * float((masks != 0).mean())

    # ---------------- ring parity: the N-rank rows of the timed e2e run vs a 1-rank replay of the same chunks (untimed) -------
    ring_parity = None
    if want_parity:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(kept, gathered, dst=0)
        if rank == 0:
            rows_n = {}
            for g in gathered:
                rows_n.update(g)
            ref = StreamingExtractor(net, W, H, 0.85, 0.85, 85, batch=B, rank=0, world=1, device=dev)
            rows_1 = {}
            run_video(world * K, True, keep=rows_1, ex=ref,
                      chunk_of=lambda c: (batch_of(pool_h, c // world, c % world), inject_of(c % world, c // world)))
            st1 = ref.finish()
            same_rows = sorted(rows_n) == sorted(rows_1) and all(
                len(rows_n[c]) == len(rows_1[c]) and all(np.array_equal(a, b) for a, b in zip(rows_n[c], rows_1[c])) for c in rows_1)
            ring_parity = {"identical": bool(same_rows and st1["n_unique"] == int(fin[0].item()) and st1["tempo_count"] == int(fin[1].item())),
                           "chunks": len(rows_1), "rows": int(sum(len(r) for c in rows_1.values() for r in c)),
                           "n_unique": [int(fin[0].item()), st1["n_unique"]], "tempo_count": [int(fin[1].item()), st1["tempo_count"]],
                           "what": "per-frame result rows, unique count and tempo_count of the %d-rank e2e run vs a 1-rank replay of the same %d chunks"
                                   % (world, world * K)}
            del ref
        barrier()
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        frames_total = world * K * B
        fps = frames_total / (ms_total / 1000.0)
        flops_step = sx.plan.flops * B
        conv_ms_step = conv_ms / max(K, 1)
        achieved = flops_step / (conv_ms_step / 1000.0) / 1e12
        n_conv = len(per_op)
        # burst peak for a short timed region (the clocks have not settled under the power cap yet), sustained for a long one
        burst = float(peaks.get("bf16_tflops", 1590.0))
        sustained = float(peaks.get("bf16_tflops_sustained", burst))
        long_run = ms_total >= 2000.0
        peak = sustained if long_run else burst
        roof = {"bound": "tensor", "kernel": "k_conv_gemm (%d launches/step)" % n_conv,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": "%s %s (timed region %.2f s: %s)" % (peak_kind, "bf16_tflops_sustained" if long_run else "bf16_tflops (burst)",
                                                                  ms_total / 1000.0, "long step sequence" if long_run else "< 2 s, burst clocks"),
                "frac_vs_burst": achieved / burst, "frac_vs_sustained": achieved / sustained,
                "traffic": conv_traffic(args), "flop_per_step": flops_step, "conv_ms_per_step": conv_ms_step,
                "conv_share_of_step": conv_ms_step / (ms_total / K)}
        dropin = None
        if world == 1 and not args.no_dropin:
            dropin = dropin_e2e(dev, net, pool_h.numpy(), K * B, B)
        gref = None
        if world == 1 and not args.no_gpu_reference and (H, W) == (1080, 1920):
            gref = gpu_reference(dev, net, pool_h[:8], sx.plan.flops)
            gref["ours_over_best"] = fps / gref["best_frames_per_s"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cc_oracle as CO
            torch.set_num_threads(os.cpu_count())
            sd = net.state_dict()
            est = CO.StabilityOracle(W, H, 0.85, 0.85, 85, use_ref_lib=CO.ref_lib() is not None)
            fr = pool_h[:4].numpy()
            cpu_reference_pass(fr[:1], sd, est)
            dt = cpu_reference_pass(fr[1:4], sd, est)
            cpu = {"value": 3 / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": "3 timed 1080p frames of this workload after 1 warm-up frame (torch-CPU fp32 FCN, oracle CC stage)"}
        cc_roof = cc_stage_roofline(dev, peaks) if (world == 1 and not args.no_cc_stage) else None
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": ("1080p synthetic whiteboard video, binarize + CC label/stats + temporal match on B200 "
                                        "(BASELINE configs[1])") if (H, W) == (1080, 1920) else
                                       ("%dx%d synthetic %s video, binarize (FCN at %dx%d) + CC label/stats + temporal match at full size "
                                        "(NOT the headline workload)" % (W, H, "chalkboard" if CHALK else "whiteboard", sx.plan.W, sx.plan.H)),
                           "frames_per_step_per_gpu": B, "frame": [H, W],
                           "weights": "random-init seed 0 (FCN_LectureNet.conf widths)",
                           "l2": ("input pool of %d frames = %.0f MB > 126 MB L2, cycled (+ ~19 GB of activations per step)%s"
                                  % (pool_n, pool_n * frame_bytes / 1e6, "; plus a 256 MB flush write between steps" if args.l2_flush else "")),
                           "parallelism": "frame chunks of %d round-robin over %d GPU(s); temporal matching is one ordered scan, its "
                                          "active-set state handed rank to rank over %s"
                                          % (B, world, "NVLink peer memory (CUDA-IPC mailboxes, stream memory ops)" if sx.handoff == "p2p"
                                             else "NCCL p2p (ring, self-staggering)"),
                           "ink_pct": round(ink_pct, 2), "ccs_per_frame": round(n_cc / max(K * B, 1), 1),
                           "unique_ccs": int(fin[0].item()), "tempo_count": int(fin[1].item())},
                "clocks": sampler.summary(), "roofline": roof, "roofline_cc_stage": cc_roof, "cpu_baseline": cpu,
                "e2e": {"value": world * K * B / (e2e_ms / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 3,
                        "d2h_bytes_per_step": int(d2h / max(K, 1))},
                "gpu_launches": launches}
        if dropin is not None:
            line["e2e_dropin"] = dropin
        if gref is not None:
            line["gpu_reference"] = gref
        if ring_parity is not None:
            line["ring_parity"] = ring_parity
        if per_rank is not None:
            line["per_rank"] = per_rank
        if numa_cores is not None:
            line["config"]["host_binding"] = "each rank pinned to its GPU's NVML CPU affinity (%d cores on rank 0) before allocating pinned memory" % len(numa_cores)
        if args.masks != "fcn":
            line["config"]["cc_masks"] = "dense glyph masks injected into the CC stage (BASELINE configs[3]); the FCN runs on the frames as always"
        print(json.dumps(line))
        sys.stdout.flush()
        if args.layer_table:
            rows = []
            for i in sorted(per_op):
                d = sx.plan.ops[i][1]
                t = float(np.mean(per_op[i]))
                fl = sx.plan.op_flops.get(i, 0) * B
                info = sx.plan.conv_plan_info(i)
                rows.append({"op": i, "N": d.NT, "Ntot": d.Ntot, "KH": d.KH, "S": 1 if (d.Sy == 2 and d.in_ystep != 2) else d.Sx,
                             "Sy": d.in_ystep if d.in_ystep else 1, "RT": d.RT, "YT": d.YT,
                             "MT": info[0], "resident": info[1], "acc_stages": info[2], "stagesA": info[3], "stagesB": info[4],
                             "ms": round(t, 4), "tflops": round(fl / (t / 1000.0) / 1e12, 1),
                             "padded_over_algorithmic": round(sx.plan.executed_flops(i) * B / max(fl, 1), 3),
                             "executed_tflops": round(sx.plan.executed_flops(i) * B / (t / 1000.0) / 1e12, 1)})
            with open(args.layer_table, "w") as f:
                json.dump(rows, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cc-stage", action="store_true", help="skip the secondary CC-stage roofline measurement")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin each rank to the CPU cores next to its GPU")
    ap.add_argument("--l2-flush", action="store_true", help="also write a 256 MB buffer between steps (the input pool alone exceeds L2)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the measurement through the reference's per-frame calls")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip timing the reference's torch/cuDNN GPU arm")
    ap.add_argument("--no-ring-parity", action="store_true", help="N > 1: skip the 1-rank replay that checks the ring's results")
    ap.add_argument("--masks", default="fcn", choices=["fcn", "glyph"],
                    help="glyph: inject dense-handwriting masks into the CC stage (the FCN still runs); not the headline workload")
    ap.add_argument("--layer-table", default=None, help="write per-layer conv timings (json) here")
    ap.add_argument("--frame-size", default=None, help="WxH other than the headline 1920x1080, e.g. 3840x2160 (BASELINE configs[4]: "
                    "chalkboard frames, FCN at the LANCZOS-halved size, CC stage at full size); not the headline line")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram__bytes_read+write per conv launch (bytes, from profiles/ ncu --set full) to report in roofline.traffic")
    args = ap.parse_args()
    if args.frame_size:
        global H, W, CHALK
        W, H = (int(v) for v in args.frame_size.lower().split("x"))
        CHALK = W * H > 2500000
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
