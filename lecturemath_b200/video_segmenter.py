"""VideoSegmenter.compute_binary_sums (R/AccessMath/preprocessing/content/video_segmenter.py:21-28), the per-frame reducer stage 04
runs over the reconstructed binary frames (R/pre_ST3D_v3.0_04_vid_segmentation.py:31-39: `Helper.decompress_binary_images` then
`binary.sum() / 255` per frame) -- SURVEY.md 8f rank 4.  The rest of VideoSegmenter (regression tree, interval logic) is scikit-learn /
list code outside the path and stays with the reference.

Frames arrive as this package's Helper returns them: PackedMask (bit-packed, from the 1-bit PNGs of stage 01 / 03) and / or uint8
arrays (any other PNG, e.g. a clean frame that wrapped to 254).  Both kinds are reduced on the device in batches -- 260 KB per
packed 1080p frame cross PCIe instead of being unpacked to 2 MB on the host -- and the result is the reference's list of floats,
bit for bit (an exact integer sum, one IEEE fp64 division by 255)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .packed_mask import PackedMask


class VideoSegmenter:
    @staticmethod
    def compute_binary_sums(all_binary, batch=64):
        lib = _lib.lib()
        out = [None] * len(all_binary)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        groups = {}                                                      # (kind, shape) -> frame positions
        for i, b in enumerate(all_binary):
            if isinstance(b, PackedMask):
                key = ("scan" if b._words is None else "words",) + b.shape
            else:
                b = np.asarray(b)
                if b.dtype != np.uint8:                                  # not a decoded PNG: numpy's own reduction, as the reference
                    out[i] = b.sum() / 255
                    continue
                key = ("u8",) + b.shape
            groups.setdefault(key, []).append(i)
        for key, idxs in groups.items():
            kind = key[0]
            for s in range(0, len(idxs), batch):
                part = idxs[s:s + batch]
                n = len(part)
                d_sums = torch.zeros(n, dtype=torch.int64, device="cuda")
                if kind == "u8":
                    host = np.stack([np.ascontiguousarray(all_binary[i]) for i in part])
                    d = torch.from_numpy(host).cuda()
                    _lib.check(lib.am_frame_sums_u8(d.data_ptr(), n, int(host[0].size), d_sums.data_ptr(), st), "am_frame_sums_u8")
                else:
                    h, w = key[1], key[2]
                    if kind == "scan":
                        d_scan = torch.from_numpy(np.stack([all_binary[i].scan for i in part])).cuda()
                        d_bits = torch.empty((n, h, lib.am_words_per_row(w)), dtype=torch.int32, device="cuda")
                        _lib.check(lib.am_png1_scanlines_to_bits(d_scan.data_ptr(), n, h, w, d_bits.data_ptr(), st), "am_png1_scanlines_to_bits")
                    else:
                        d_bits = torch.from_numpy(np.stack([all_binary[i].words for i in part]).view(np.int32)).cuda()
                    _lib.check(lib.am_frame_sums_bits(d_bits.data_ptr(), n, h, w, d_sums.data_ptr(), st), "am_frame_sums_bits")
                for i, v in zip(part, d_sums.cpu().numpy().astype(np.uint64)):
                    out[i] = v / 255                                      # numpy uint64 scalar / int -> float64, as `binary.sum() / 255`
        return out
