"""GPU parity tests for the FCN binarizer (tcgen05 implicit-GEMM kernels, through the C ABI).

Floating point: bf16 operands, fp32 accumulation.  Stated tolerance (BASELINE.json north_star): probability maps
within 1e-2 absolute of the fp32 reference, mask disagreement <= 0.1 % of pixels."""
import numpy as np
import pytest
import torch

from oracle import fcn_oracle as FO
from tests.test_fcn_host_logic import golden_net

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2
MASK_TOL = 1e-3


def _sig(a):
    return 1.0 / (1.0 + np.exp(-a))


def _check_mask(ink, ink_ref, p_ref):
    """north_star bar: mask disagreement <= 0.1 % of pixels AND confined to near-threshold pixels, i.e. every pixel that
    differs has a reference probability within PROB_TOL of the 128/255 decision point (ink = p*255 truncated < 128)."""
    bad = ink != ink_ref
    assert bad.mean() <= MASK_TOL, "mask disagreement %.4f %% (%d of %d pixels)" % (100 * bad.mean(), bad.sum(), bad.size)
    if bad.any():
        far = np.abs(p_ref[bad] - 128.0 / 255.0).max()
        assert far < PROB_TOL, "a disagreeing pixel sits %.4f from the threshold in the reference" % far


@pytest.mark.parametrize("tag", ["tiny", "full"])
@pytest.mark.parametrize("mode", ["rowrun", "kx", "sy2", "mt1", "mt2", "mt4", "mt22", "poolall", "nopool", "poolx"])
def test_forward_vs_reference_golden(golden, tag, mode):
    """mode: rowrun = planner's choice, kx = one TMA load per horizontal tap, sy2 = 2-D (row-pair) packing forced,
    mtN = N M-tiles per work item (N MMA issuer warps) forced wherever TMEM / shared memory allow; mt22 = CTA pairs
    (cta_group::2 clusters, M = 256 MMAs); poolall / nopool = MaxPool2d fused into the epilogue of every eligible encoder conv /
    of none (the planner's default fuses it where K > 640); poolx = conv_down_block_1 / 2 forced into Sx = 4 / 2, Sy = 1 packings whose
    pool is fused through the unit-pair epilogue (kPOOLX)."""
    z = golden("fcn_forward.npz")
    net = golden_net(tag, z).cuda()
    net.rowrun = mode != "kx"
    net.plan_overrides = {"sy": 2} if mode == "sy2" else ({"mt": int(mode[2:])} if mode.startswith("mt") else None)
    if mode in ("poolall", "nopool"):
        net.plan_overrides = {"fused_pool": "all"} if mode == "poolall" else {"no_fused_pool": True}
    if mode == "poolx":
        from tests.test_fcn_host_logic import poolx_overrides
        net.plan_overrides = poolx_overrides(net)
    frame = z["frame_bgr"]
    plan = net.binarize_frames(frame[None], want_others=True)
    torch.cuda.synchronize()
    if mode == "poolx":
        assert sum(1 for k, d in plan.ops if k == "conv" and d.pool_out and d.Sx >= 2 and d.Sy == 1) == 2
    n_pool_ops = sum(1 for k, _ in plan.ops if k == "pool")
    n_fused = sum(1 for k, d in plan.ops if k == "conv" and d.pool_out)
    assert n_pool_ops + n_fused == 5
    if mode == "poolall":
        assert n_fused >= 3, "only %d encoder convs run with the fused max-pool" % n_fused
    if mode == "nopool":
        assert n_fused == 0
    if mode.startswith("mt"):
        want = int(mode[2:])
        n_forced = sum(1 for i, (k, d) in enumerate(plan.ops) if k == "conv" and
                       ((d.flags & 16) != 0 if want == 22 else ((d.flags & 16) == 0 and plan.conv_plan_info(i)[0] == want)))
        assert n_forced >= 8, "only %d conv launches run with MT = %d" % (n_forced, want)
    p = _sig(plan.logits[0].cpu().numpy())
    assert np.abs(p - _sig(z[tag + "_logit"])).max() < PROB_TOL
    assert np.abs(_sig(plan.text_logit[0].cpu().numpy()) - _sig(z[tag + "_text_logit"])).max() < PROB_TOL
    assert np.abs(plan.rec[0].permute(2, 0, 1).cpu().numpy() - z[tag + "_rec_raw"]).max() < 2e-2
    ink, text_m, rec = net.masks_from_plan(plan, 0)
    _check_mask(ink, z[tag + "_binary"], _sig(z[tag + "_logit"]))
    assert np.abs(rec.astype(int) - z[tag + "_rec"].astype(int)).max() <= 3


def test_binarize_and_worker_api_vs_oracle(golden):
    from PIL import Image
    import cv2
    from lecturemath_b200.fcn_binarizer_worker import FCN_LectureNet_Binarizer
    z = golden("fcn_forward.npz")
    net = golden_net("tiny", z).cuda()
    frame = z["frame_bgr"]
    pil = Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_RGB2BGR))
    binary, text_m, rec = net.binarize(pil, return_others=True, force_binary=True)
    _check_mask(255 - binary, z["tiny_binary"], _sig(z["tiny_logit"]))
    soft = net.binarize(pil)
    sd = {k: v for k, v in net.state_dict().items()}
    ref_soft, _, _ = FO.binarize(sd, frame[:, :, ::-1], force_binary=False)
    assert np.abs(soft.astype(int) - ref_soft.astype(int)).max() <= 3
    worker = FCN_LectureNet_Binarizer(net)
    worker.initialize(frame.shape[1], frame.shape[0])
    worker.handleFrame(frame, None, 0, 40.0, 40.0, 7)
    assert worker.frame_indices == [7] and worker.frame_times == [40.0] and worker.getWorkName()
    decoded = cv2.imdecode(worker.compressed_frames[0], cv2.IMREAD_GRAYSCALE)
    np.testing.assert_array_equal(decoded, worker.last_binary)
    _check_mask(decoded, z["tiny_binary"], _sig(z["tiny_logit"]))
    lg, tx, rc = net.forward(net.prepare_image(pil))
    assert lg.shape == (1, 1) + frame.shape[:2] and rc.shape == (1, 3) + frame.shape[:2]
    assert np.abs(_sig(lg[0, 0].cpu().numpy()) - _sig(z["tiny_logit"])).max() < PROB_TOL


@pytest.mark.parametrize("hw", [(720, 1280), (1080, 1920), (1080, 1920, "poolx")])
def test_full_size_random_init_vs_fp32_oracle_on_gpu(hw):
    """BASELINE configs 1/2 shapes with the reference's seed-0 random init: compare with the fp32 oracle run on the
    same GPU (torch/cuDNN fp32, TF32 off).  Also a size-independent property: batch invariance.  "poolx": the encoder convs 1 / 2 in
    the Sx-packed form whose max-pool is fused through unit pairs (what shapes without a tuned table get)."""
    from lecturemath_b200 import synth
    from lecturemath_b200.configuration import Configuration
    from lecturemath_b200.fcn_lecturenet import FCN_LectureNet
    from tests.conftest import GOLDEN
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h, w = hw[:2]
    torch.manual_seed(0)
    net = FCN_LectureNet.CreateFromConfig(Configuration.from_file(GOLDEN + "/fcn_full.conf"), 3, False).eval().cuda()
    if len(hw) > 2:
        from tests.test_fcn_host_logic import poolx_overrides
        net.plan_overrides = poolx_overrides(net)
    frames = np.stack(list(synth.whiteboard_frames(2, h, w, seed=1234)))
    plan = net.binarize_frames(frames)
    torch.cuda.synchronize()
    logits = plan.logits.clone()
    bits = plan.bits.clone()
    sd = {k: v.cuda() for k, v in net.state_dict().items()}
    for f in range(2):
        x0 = FO.prepare_image(frames[f][:, :, ::-1]).cuda()
        ref, _, _ = FO.forward(sd, x0)
        p, pr = torch.sigmoid(logits[f]), torch.sigmoid(ref[0, 0])
        assert (p - pr).abs().max().item() < PROB_TOL
        ink = ((p * 255).to(torch.uint8) < 128)
        ink_ref = ((pr * 255).to(torch.uint8) < 128)
        dis = (ink != ink_ref).float().mean().item()
        assert dis <= MASK_TOL, "mask disagreement %.4f%%" % (100 * dis)
    single = net.binarize_frames(frames[1:2])
    torch.cuda.synchronize()
    assert torch.equal(single.bits[0], bits[1])                          # batch invariance (bit-exact)


@pytest.mark.gpu
def test_programmatic_dependent_launch_gives_identical_bits():
    """AM_B200_PDL=1 (opt-in): the conv launches carry programmaticStreamSerialization and wait for their predecessor with
    griddepcontrol.wait -- same kernels, same arithmetic, so the masks and logits must be bit-identical to the default launches.
    The switch is read once per process: both arms run in their own interpreter."""
    import os
    import subprocess
    import sys
    from tests.conftest import GOLDEN
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, hashlib, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from lecturemath_b200 import synth\n"
        "from lecturemath_b200.configuration import Configuration\n"
        "from lecturemath_b200.fcn_lecturenet import FCN_LectureNet\n"
        "torch.manual_seed(0)\n"
        "net = FCN_LectureNet.CreateFromConfig(Configuration.from_file(%r), 3, False).eval().cuda()\n"
        "frames = np.stack(list(synth.whiteboard_frames(2, 720, 1280, seed=99)))\n"
        "h = hashlib.sha256()\n"
        "for _ in range(3):\n"
        "    plan = net.binarize_frames(frames)\n"
        "    torch.cuda.synchronize()\n"
        "    h.update(plan.bits.cpu().numpy().tobytes()); h.update(plan.logits.cpu().numpy().tobytes())\n"
        "print('HASH', h.hexdigest())\n") % (repo, os.path.join(GOLDEN, "fcn_full.conf"))
    hashes = []
    for flag in ("0", "1"):
        env = dict(os.environ, AM_B200_PDL=flag)
        out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        hashes.append([ln for ln in out.stdout.splitlines() if ln.startswith("HASH")][-1])
    assert hashes[0] == hashes[1]
