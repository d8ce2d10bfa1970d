"""The 01 -> 02 wire format written on the device.

The reference stores every binarized frame as PNG bytes (`flag, raw_data = cv2.imencode(".png", binary)`,
R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:56-64) and reads them back with
`cv2.imdecode(raw, IMREAD_GRAYSCALE)` (R/AccessMath/preprocessing/content/helper.py:27-34).  encode_png_frames produces entries
with the same type (1-D uint8 numpy arrays) and the same decoded pixels, from the bit-packed masks the FCN left in HBM:
csrc/png.cu writes a 1-bit grayscale PNG with stored deflate blocks (checksums included) and only the finished files cross PCIe."""
import ctypes

import numpy as np
import torch

from . import _lib


class PngEncoder:
    def __init__(self, width, height, max_batch=1, device=None):
        self.lib = _lib.lib()
        self.width, self.height, self.max_batch = int(width), int(height), int(max_batch)
        self.size = int(self.lib.am_png1_size(self.width, self.height))
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.d_out = torch.empty((self.max_batch, self.size), dtype=torch.uint8, device=self.device)
        self.h_out = torch.empty((self.max_batch, self.size), dtype=torch.uint8).pin_memory()

    def encode(self, bits, n=None):
        """bits: bit-packed CUDA masks [n][H][WPR] -> list of n uint8 arrays (one PNG file each)."""
        n = bits.shape[0] if n is None else n
        if n > self.max_batch:
            raise ValueError("batch %d exceeds the encoder's capacity %d" % (n, self.max_batch))
        st = torch.cuda.current_stream(self.device)
        _lib.check(self.lib.am_png1_encode(bits.data_ptr(), n, self.height, self.width, self.d_out.data_ptr(), ctypes.c_void_p(st.cuda_stream)),
                   "am_png1_encode")
        self.h_out[:n].copy_(self.d_out[:n], non_blocking=True)
        st.synchronize()
        return [self.h_out[f].numpy().copy() for f in range(n)]


def encode_png_frames(bits, width, height):
    """One-shot form of PngEncoder.encode."""
    return PngEncoder(width, height, bits.shape[0], bits.device).encode(bits)
