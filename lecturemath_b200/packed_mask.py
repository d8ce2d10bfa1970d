"""Bit-packed binary frames on the host side of the 01 -> 02 wire format ("a compact packed-bit container
with a lazy compat view").

The reference decodes every entry of compressed_frames to a uint8 H x W array up front (`cv2.imdecode(raw, IMREAD_GRAYSCALE)`,
R/AccessMath/preprocessing/content/helper.py:27-34: 14.8 ms and 2 MB per 1080p frame, the whole video held in RAM) and hands the
arrays to CCStabilityEstimator.add_frame one by one (R/pre_ST3D_v3.0_02_cc_analaysis.py:24-38).  PackedMask is what
Helper.decompress_binary_images returns for the 1-bit PNGs this package writes: the frame stays bit-packed (260 KB at 1080p, the
layout of the device masks: uint32 [H][WPR], bit b of word w = pixel 32 w + b, ink = 1) and only turns into the uint8 0 / 255 array
when somebody asks for pixels (np.asarray, indexing).  CCStabilityEstimator.add_frame takes the packed words as they are."""
import struct
import zlib

import numpy as np

_PNG_SIG = b"\x89PNG\r\n\x1a\n"
_REV = np.array([int("{:08b}".format(i)[::-1], 2) for i in range(256)], dtype=np.uint8)     # bit reversal of a byte


def words_per_row(width):
    """am_words_per_row (csrc/am_common.cuh): 32-bit words per mask row, padded to a multiple of 4 (128-bit loads)."""
    return ((int(width) + 31) // 32 + 3) // 4 * 4


class PackedMask:
    """Lazy uint8 (H, W) 0 / 255 view of a bit-packed binary frame.  Holds the frame either as device-layout words (uint32 (H, WPR),
    padding bits zero) or as the PNG's own scanline bytes (uint8 (H, 1 + ceil(W / 8)): filter byte 0, then pixels MSB first) --
    the estimator uploads whichever is there and am_png1_scanlines_to_bits converts scanlines on the device."""
    __slots__ = ("_words", "scan", "shape", "_dense")
    dtype = np.dtype(np.uint8)
    ndim = 2

    def __init__(self, height, width, words=None, scan=None):
        self._words, self.scan = words, scan
        self.shape = (int(height), int(width))
        self._dense = None

    @property
    def words(self):
        if self._words is None:
            h, w = self.shape
            rb = (w + 7) // 8
            rows = np.zeros((h, words_per_row(w) * 4), dtype=np.uint8)
            rows[:, :rb] = _REV[self.scan[:, 1:]]               # PNG packs the leftmost pixel into the MSB
            if w & 7:
                rows[:, rb - 1] &= np.uint8((1 << (w & 7)) - 1)   # spec: the unused low bits of the last byte are unspecified
            self._words = rows.view("<u4")
        return self._words

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    def __len__(self):
        return self.shape[0]

    def dense(self):
        if self._dense is None:
            h, w = self.shape
            bits = np.unpackbits(np.ascontiguousarray(self.words).view(np.uint8).reshape(h, -1), axis=1, bitorder="little")
            self._dense = np.ascontiguousarray(bits[:, :w]) * np.uint8(255)
        return self._dense

    def __array__(self, dtype=None, copy=None):
        a = self.dense()
        return a if dtype is None or np.dtype(dtype) == a.dtype else a.astype(dtype)

    def __getitem__(self, idx):
        return self.dense()[idx]

    def astype(self, dtype, *a, **k):
        return self.dense().astype(dtype, *a, **k)

    def count_nonzero(self):
        """Ink pixels, without unpacking (VideoSegmenter.compute_binary_sums, R/AccessMath/preprocessing/content/video_segmenter.py:21-28)."""
        return int(np.unpackbits(np.ascontiguousarray(self.words).view(np.uint8)).sum())

    @staticmethod
    def from_dense(mask):
        m = np.asarray(mask) != 0
        h, w = m.shape
        wpr = words_per_row(w)
        rows = np.zeros((h, wpr * 4), dtype=np.uint8)
        packed = np.packbits(m, axis=1, bitorder="little")
        rows[:, :packed.shape[1]] = packed
        return PackedMask(h, w, words=rows.view("<u4"))


def parse_png1(raw):
    """PNG bytes -> PackedMask when the file is a 1-bit grayscale, non-interlaced PNG whose scanlines all use filter 0 (what
    csrc/png.cu writes, with stored or compressed deflate blocks); None for any other PNG (the caller falls back to cv2.imdecode)."""
    b = raw.tobytes() if isinstance(raw, np.ndarray) else bytes(raw)
    if len(b) < 57 or b[:8] != _PNG_SIG or b[12:16] != b"IHDR":
        return None
    w, h, depth, colour, comp, flt, interlace = struct.unpack(">IIBBBBB", b[16:29])
    if (depth, colour, comp, flt, interlace) != (1, 0, 0, 0, 0):
        return None
    pos, idat = 33, []
    while pos + 8 <= len(b):
        n, kind = struct.unpack(">I", b[pos:pos + 4])[0], b[pos + 4:pos + 8]
        if kind == b"IDAT":
            idat.append(b[pos + 8:pos + 8 + n])
        elif kind == b"IEND":
            break
        elif kind in (b"tRNS", b"PLTE", b"gAMA", b"sBIT"):        # anything that changes how samples map to grey levels
            return None
        pos += 12 + n
    rb = (w + 7) // 8
    try:
        data = zlib.decompress(b"".join(idat))                    # also verifies the Adler-32 of the stream
    except zlib.error:
        return None
    if len(data) != h * (1 + rb):
        return None
    lines = np.frombuffer(data, dtype=np.uint8).reshape(h, 1 + rb)
    if lines[:, 0].any():                                         # a filtered scanline: leave it to a full decoder
        return None
    return PackedMask(h, w, scan=lines)
