"""CPU: host-side logic of the FCN path -- seeded init equals the reference's, BN folding, weight packing and
the row-run GEMM formulation (emulated in torch from the very descriptors the CUDA kernel receives) reproduce
the reference's logits within bf16 tolerance."""
import hashlib

import numpy as np
import pytest
import torch

from lecturemath_b200.configuration import Configuration
from lecturemath_b200.fcn_lecturenet import FCN_LectureNet, FCNPlan
from tests.conftest import GOLDEN
from tests.emulate_fcn import emulate_plan


def _state_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode()); h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def golden_net(tag, z):
    """Rebuild the exact weights oracle/gen_golden.py gave the reference (seed 0 + perturbed BatchNorm)."""
    path = GOLDEN + ("/fcn_tiny.conf" if tag == "tiny" else "/fcn_full.conf")
    cfg = Configuration.from_file(path)
    torch.manual_seed(0)
    net = FCN_LectureNet.CreateFromConfig(cfg, 3, False)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for _, mod in net.params.named_modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                mod.weight.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                mod.bias.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    return net.eval()


def test_seeded_init_reproduces_reference_weights(golden):
    z = golden("fcn_forward.npz")
    net = golden_net("tiny", z)
    assert _state_hash(net.state_dict()) == str(z["tiny_sd_hash"])
    for k, v in net.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), z["tiny_sd/" + k])


def poolx_overrides(net):
    """Forces conv_down_block_1 / 2 into Sx = 4 / 2, Sy = 1 packings with one M-tile per item: the max-pool is fused through the
    unit-pair epilogue (csrc/fcn_conv.cu kPOOLX)."""
    sd = net.state_dict()
    c1, c2 = sd["conv_down_block_1.0.weight"].shape[0], sd["conv_down_block_2.0.weight"].shape[0]
    return {"cfg": {"conv_down_block_1": (4, 1, 4 * c1, 1), "conv_down_block_2": (2, 1, 2 * c2, 1)}}


@pytest.mark.parametrize("tag,mode", [("tiny", "rowrun"), ("tiny", "kx"), ("tiny", "sy2"), ("tiny", "poolx"), ("full", "rowrun"), ("full", "poolx")])
def test_emulated_kernel_plan_matches_reference_logits(golden, tag, mode):
    """mode: rowrun = planner's choice, kx = one TMA load per horizontal tap, sy2 = 2-D (row-pair) packing forced,
    poolx = Sx-packed encoder convs with the max-pool fused.  tag: the golden file's tiny net or the reference's full widths."""
    z = golden("fcn_forward.npz")
    net = golden_net(tag, z)
    frame = z["frame_bgr"]
    plan = FCNPlan(net.params, 1, frame.shape[0], frame.shape[1], torch.device("cpu"), rowrun=(mode != "kx"),
                   overrides={"sy": 2} if mode == "sy2" else (poolx_overrides(net) if mode == "poolx" else None))
    if mode == "sy2":
        assert sum(1 for k, d in plan.ops if k == "conv" and d.in_ystep == 2) >= 10
    if mode == "poolx":
        assert sum(1 for k, d in plan.ops if k == "conv" and d.pool_out and d.Sx >= 2 and d.Sy == 1) == 2
    logits, text, rec, ink = emulate_plan(plan, frame[None])
    # bf16 activations/weights with fp32 accumulation: stated tolerance 1e-2 on probabilities (BASELINE north_star)
    p, p_ref = torch.sigmoid(logits[0]).numpy(), 1 / (1 + np.exp(-z[tag + "_logit"]))
    assert np.abs(p - p_ref).max() < 1e-2
    assert np.abs(torch.sigmoid(text[0]).numpy() - 1 / (1 + np.exp(-z[tag + "_text_logit"]))).max() < 1e-2
    assert np.abs(rec[0].permute(2, 0, 1).numpy() - z[tag + "_rec_raw"]).max() < 2e-2
    ref_ink = z[tag + "_binary"] > 0
    assert (ink[0] != ref_ink).mean() <= 2e-3
    assert plan.flops > 0


def test_full_config_flops_match_survey():
    cfg = Configuration.from_file(GOLDEN + "/fcn_full.conf")
    net = FCN_LectureNet.CreateFromConfig(cfg, 3, False)
    assert sum(p.numel() for p in net.parameters()) == 15815538
    for (h, w, gf) in ((720, 1280, 414.5), (1080, 1920, 931.8)):
        plan = FCNPlan.__new__(FCNPlan)                                    # flop count only: no buffers
        assert abs(FCNPlan.count_flops(net.params, h, w) / 1e9 - gf) < 0.5


@pytest.mark.parametrize("hw", [(1080, 1920), (720, 1280)])
def test_tuned_table_entries_are_adopted_for_their_shapes(hw):
    """Every (S, Sy, NT, MT) of lecturemath_b200/tuned_plans.json for the reference architecture at batch 8 is among the feasible
    candidates of its layer (a stale table would silently fall back to the cycle model), and the plan built from it pools all five
    encoder levels in the conv epilogues (no separate k_maxpool2 pass in the headline shapes)."""
    from lecturemath_b200 import fcn_lecturenet as F
    h, w = hw
    net = FCN_LectureNet.CreateFromConfig(Configuration.from_file(GOLDEN + "/fcn_full.conf"), 3, False)
    plan = FCNPlan(net.params, 8, h, w, torch.device("cpu"))
    table = F.tuned_table().get(plan.arch, {}).get("B8_%dx%d" % (h, w))
    assert table, "no tuned entry for B8 %dx%d under architecture %s" % (h, w, plan.arch)
    for name, cfg in table.items():
        assert name in plan.specs, name
        assert tuple(plan.specs[name]["cfg"]) == tuple(cfg), "%s: tuned %s not adopted (plan uses %s)" % (name, cfg, plan.specs[name]["cfg"])
    kinds = [k for k, _ in plan.ops]
    assert kinds.count("pool") == 0 and kinds.count("conv") == 20                  # 15 conv launches (the two heads share one) + 5 transposed convs
    assert sum(1 for k, d in plan.ops if k == "conv" and d.pool_out) == 5
