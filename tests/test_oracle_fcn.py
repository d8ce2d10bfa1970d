"""CPU: the torch-fp32 FCN oracle against outputs captured from the unmodified reference module."""
import numpy as np
import torch

from oracle import fcn_oracle as F


def _tiny_sd(z):
    return {k[len("tiny_sd/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("tiny_sd/")}


def test_tiny_forward_and_binarize_match_reference(golden):
    z = golden("fcn_forward.npz")
    sd = _tiny_sd(z)
    frame = z["frame_bgr"]
    logit, text, rec = F.forward(sd, F.prepare_image(frame[:, :, ::-1]))
    # same ops in the same order on the same backend: expect (near) bit equality; tolerance covers thread-count effects
    np.testing.assert_allclose(logit[0, 0].numpy(), z["tiny_logit"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(text[0, 0].numpy(), z["tiny_text_logit"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(rec[0].numpy(), z["tiny_rec_raw"], atol=2e-5, rtol=0)
    ink, text_m, rec_bgr = F.handle_frame(sd, frame)
    assert (ink != z["tiny_binary"]).mean() <= 1e-3
    assert (text_m != z["tiny_text"]).mean() <= 1e-3
    assert np.abs(rec_bgr.astype(int) - z["tiny_rec"].astype(int)).max() <= 1


def test_threshold_rule_is_logit_ge_ln_128_over_127():
    # (uint8)(sigmoid(z)*255) >= 128  <=>  z >= ln(128/127)   (SURVEY.md "hard parts")
    zs = torch.linspace(-0.05, 0.05, 20001)
    u8 = (torch.sigmoid(zs).numpy() * 255).astype(np.uint8)
    thr = float(np.log(128.0 / 127.0))
    agree = (u8 >= 128) == (zs.numpy() >= thr)
    assert agree.mean() > 0.9995          # only float rounding right at the threshold may differ
