"""oracle/fcn_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU/torch-fp32 oracle for the FCN binarizer).

Functional fp32 restatement of FCN_LectureNet inference, driven directly by a reference-format state_dict
(R/ = /root/reference/ACCESS2021_release/):
  * encode_decode   <- R/AccessMath/lecturenet_v1/FCN_lecturenet.py:260-323 (layers :26-139)
  * forward         <- FCN_lecturenet.py:364-403 (non-reconstruction branch; heads :153-160, :164-201)
  * prepare_image   <- FCN_lecturenet.py:607-618
  * binarize        <- FCN_lecturenet.py:430-505 (<=2.5 MP guard, sigmoid, *255 -> uint8, >=128 threshold)
  * handle_frame    <- R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:47-60
All arithmetic is PyTorch's (the reference's own third-party dependency, unpinned; torch 2.11.0 here).

Pinned by tests/golden/fcn_forward.npz (outputs of the unmodified reference, oracle/gen_golden.py).
Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import cv2
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _cbn(sd, name, x, act, padding):
    """Conv2d -> BatchNorm2d(eval) -> activation, i.e. one nn.Sequential block of the reference."""
    x = F.conv2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], stride=1, padding=padding)
    x = F.batch_norm(x, sd[name + ".1.running_mean"], sd[name + ".1.running_var"], sd[name + ".1.weight"],
                     sd[name + ".1.bias"], training=False, eps=BN_EPS)
    if act == "gelu":
        x = F.gelu(x)                      # nn.GELU() = exact erf form
    elif act == "tanh":
        x = torch.tanh(x)
    return x


def _up(sd, k, x, out_hw):
    """ConvTranspose2d(k=2,s=2, output_size=skip shape) -> BatchNorm2d -> GELU (FCN_lecturenet.py:280-281)."""
    w = sd["transposed_conv_%d.weight" % k]
    oph, opw = out_hw[0] - 2 * x.shape[2], out_hw[1] - 2 * x.shape[3]
    x = F.conv_transpose2d(x, w, sd["transposed_conv_%d.bias" % k], stride=2, padding=0, output_padding=(oph, opw))
    n = "upsample_block_%d" % k
    x = F.batch_norm(x, sd[n + ".0.running_mean"], sd[n + ".0.running_var"], sd[n + ".0.weight"], sd[n + ".0.bias"],
                     training=False, eps=BN_EPS)
    return F.gelu(x)


def encode_decode(sd, x0):
    p = (sd["conv_down_block_1.0.weight"].shape[-1] - 1) // 2
    pre, x = [], x0
    for k in range(1, 6):                                  # :263-276
        x = _cbn(sd, "conv_down_block_%d" % k, x, "gelu", p)
        pre.append(x)
        x = F.max_pool2d(x, 2)                             # floors odd sizes
    x = _cbn(sd, "mid_block", x, "gelu", p)                # :278
    sizes = [x0.shape[2:]] + [F.max_pool2d(t, 2).shape[2:] for t in pre[:4]]   # x0, x_conv1..4
    for k in range(5, 0, -1):                              # :280-303
        x = _up(sd, k, x, sizes[k - 1])
        x = torch.cat((x, pre[k - 1]), 1)
        x = _cbn(sd, "conv_up_block_%d" % k, x, "gelu", p)
    return x


@torch.no_grad()
def forward(sd, x0):
    """-> (output_logit, text_mask_logit, rec_img), each (N, C, H, W) fp32  (FCN_lecturenet.py:364-403)."""
    x_up1 = encode_decode(sd, x0)
    pk = (sd["conv_pixels_1.0.weight"].shape[-1] - 1) // 2
    text = _cbn(sd, "conv_text_mask_out", x_up1, None, pk)                                   # :370
    rec = _cbn(sd, "conv_reconstruct", x_up1, "tanh", (sd["conv_reconstruct.0.weight"].shape[-1] - 1) // 2)   # :376
    diff = (x0 - rec) * torch.sigmoid(text)                                                  # :377
    p1 = _cbn(sd, "conv_pixels_1", torch.cat((diff, x_up1), 1), "gelu", pk)                  # :383-386
    p2 = _cbn(sd, "conv_pixels_2", torch.cat((diff, p1), 1), "gelu", pk)                     # :390-393
    out = _cbn(sd, "conv_out", torch.cat((diff, p2), 1), None, pk)                           # :397-400
    return out, text, rec


def prepare_image(rgb_u8):
    """uint8 HxWx3 RGB -> (1,3,H,W) fp32 in [-1,1]  (to_tensor + normalize(.5,.5), :607-618)."""
    t = torch.from_numpy(np.ascontiguousarray(rgb_u8)).permute(2, 0, 1).float().div(255.0)
    return ((t - 0.5) / 0.5).unsqueeze(0)


def _u8(prob):
    return (prob.numpy() * 255).astype(np.uint8)           # :461-462 (truncation)


@torch.no_grad()
def binarize(sd, rgb_u8, force_binary=True, threshold=128):
    """-> (binary, text_mask, rec_bgr) uint8, as FCN_LectureNet.binarize(pil, True, force_binary) (:430-505).

    Images above 2.5 MP are halved with PIL LANCZOS first and masks resized back with INTER_NEAREST."""
    import PIL.Image
    h0, w0 = rgb_u8.shape[:2]
    img = rgb_u8
    h, w = h0, w0
    while w * h > 2500000:                                 # :434-437
        img = np.asarray(PIL.Image.fromarray(img).resize((int(w / 2), int(h / 2)), PIL.Image.LANCZOS))
        h, w = img.shape[:2]
    logit, text, rec = forward(sd, prepare_image(img))
    binary = _u8(torch.sigmoid(logit)[0, 0])
    text_m = _u8(torch.sigmoid(text)[0, 0])
    if force_binary:                                       # :464-467, :473-476
        binary = np.where(binary >= threshold, 255, 0).astype(np.uint8)
        text_m = np.where(text_m >= threshold, 255, 0).astype(np.uint8)
    r = rec[0].numpy().transpose(1, 2, 0) * 0.5 + 0.5      # from_img_space_to_cv2 :532-554
    r = np.clip(r[:, :, ::-1] * 255, 0, 255).astype(np.uint8)
    if w != w0:                                            # :481-494
        interp = cv2.INTER_NEAREST if force_binary else cv2.INTER_CUBIC
        binary = cv2.resize(binary, (w0, h0), interpolation=interp)
        text_m = cv2.resize(text_m, (w0, h0), interpolation=interp)
        r = cv2.resize(r, (w0, h0), interpolation=cv2.INTER_NEAREST)
    return binary, text_m, r


def handle_frame(sd, frame_bgr):
    """FCN_LectureNet_Binarizer.handleFrame minus PNG: BGR frame -> ink mask (ink = 255)  (:50-54)."""
    binary, text_m, rec = binarize(sd, frame_bgr[:, :, ::-1], True)
    return 255 - binary, text_m, rec
