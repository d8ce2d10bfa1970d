"""Frames above 2.5 MP (4K lecture video) on the device.

FCN_LectureNet.binarize halves such images with PIL LANCZOS until they fit (R/AccessMath/lecturenet_v1/
FCN_lecturenet.py:434-437), runs the network at the reduced size and resizes the thresholded masks back with
cv2.INTER_NEAREST (:481-494); the CC stage then works at the ORIGINAL size.  LargeFrameAdapter does both resizes on the
GPU (csrc/resize.cu: am_lanczos_resize_u8, am_bits_resize_nearest -- bit-identical to Pillow / OpenCV), so a 4K frame
crosses PCIe once and never visits the host again."""
import ctypes

import torch

from . import _lib


def working_sizes(width, height):
    """[(w0, h0), (w1, h1), ...]: the sizes the 2.5 MP guard visits; the FCN runs at the last one."""
    sizes = [(int(width), int(height))]
    while sizes[-1][0] * sizes[-1][1] > 2500000:
        w, h = sizes[-1]
        sizes.append((int(w / 2), int(h / 2)))
    return sizes


class LargeFrameAdapter:
    def __init__(self, batch, height, width, device):
        self.lib = _lib.lib()
        self.batch, self.height, self.width, self.device = batch, height, width, torch.device(device)
        self.sizes = working_sizes(width, height)
        ow, oh = ctypes.c_int(0), ctypes.c_int(0)
        n = self.lib.am_fcn_working_size(width, height, ctypes.byref(ow), ctypes.byref(oh))
        assert n == len(self.sizes) - 1 and (ow.value, oh.value) == self.sizes[-1]
        self.fcn_width, self.fcn_height = self.sizes[-1]
        self.active = len(self.sizes) > 1
        if self.active:
            self.frames = torch.empty((batch, height, width, 3), dtype=torch.uint8, device=self.device)
            self.mid = [torch.empty((batch, h, w, 3), dtype=torch.uint8, device=self.device) for (w, h) in self.sizes[1:-1]]
            self.bits = torch.zeros((batch, height, self.lib.am_words_per_row(width)), dtype=torch.int32, device=self.device)
        self.launches_per_run = len(self.sizes) if self.active else 0      # one LANCZOS launch per halving + one mask upscale

    def downscale(self, dst_frames, stream, src=None):
        """self.frames (or src) [B][H][W][3] uint8 -> dst_frames [B][h][w][3] (the FCN plan's input buffer)."""
        st = ctypes.c_void_p(stream)
        cur = self.frames if src is None else src
        chain = self.mid + [dst_frames]
        for (w0, h0), (w1, h1), dst in zip(self.sizes[:-1], self.sizes[1:], chain):
            _lib.check(self.lib.am_lanczos_resize_u8(cur.data_ptr(), self.batch, h0, w0, 3, h1, w1, dst.data_ptr(), st),
                       "am_lanczos_resize_u8")
            cur = dst

    def upscale_bits(self, bits_small, stream, out=None):
        """bit-packed mask at the FCN size -> self.bits at the original size (one INTER_NEAREST resize, :481-486)."""
        out = self.bits if out is None else out
        _lib.check(self.lib.am_bits_resize_nearest(bits_small.data_ptr(), self.batch, self.fcn_height, self.fcn_width, self.height,
                                                   self.width, out.data_ptr(), ctypes.c_void_p(stream)), "am_bits_resize_nearest")
        return out


class LargeView:
    """What FCN_LectureNet.binarize_frames returns for frames above 2.5 MP: the FCN plan (logits, text_logit, rec at the
    reduced size) with `bits`, `H`, `W` replaced by the full-size ink mask."""

    def __init__(self, plan, adapter):
        self.plan, self.adapter = plan, adapter
        self.bits, self.H, self.W = adapter.bits, adapter.height, adapter.width

    def __getattr__(self, name):
        return getattr(self.plan, name)
