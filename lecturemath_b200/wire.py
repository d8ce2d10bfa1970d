"""The 01 -> 02 wire format written on the device.

The reference stores every binarized frame as PNG bytes (`flag, raw_data = cv2.imencode(".png", binary)`,
R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:56-64) and reads them back with
`cv2.imdecode(raw, IMREAD_GRAYSCALE)` (R/AccessMath/preprocessing/content/helper.py:27-34).  PngEncoder produces entries
with the same type (1-D uint8 numpy arrays) and the same decoded pixels, from the bit-packed masks the FCN left in HBM:
csrc/png.cu writes a 1-bit grayscale PNG -- compressed (fixed-Huffman deflate with run-length matches, the default: the size of
cv2's files on whiteboard masks) or with stored blocks -- checksums included, and only the finished files cross PCIe."""
import ctypes

import torch

from . import _lib

FIRST_D2H_BYTES = 64 << 10           # per frame in the first read-back of the compressed form (typical files are 10-30 KB)


class PngEncoder:
    def __init__(self, width, height, max_batch=1, device=None, compress=True, depth=1):
        """depth = 1: input = bit-packed masks [n][H][WPR]; depth = 8: input = uint8 grayscale frames [n][H][W] (compressed only)."""
        self.lib = _lib.lib()
        self.width, self.height, self.max_batch, self.compress = int(width), int(height), int(max_batch), bool(compress)
        self.depth = int(depth)
        if self.depth == 8:
            if not compress:
                raise ValueError("the 8-bit writer only has the compressed form")
            self.size = int(self.lib.am_png8_capacity(self.width, self.height))
        else:
            self.size = int(self.lib.am_png1_capacity(self.width, self.height) if compress else self.lib.am_png1_size(self.width, self.height))
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.d_out = torch.empty((self.max_batch, self.size), dtype=torch.uint8, device=self.device)
        self.h_out = torch.empty((self.max_batch, self.size), dtype=torch.uint8).pin_memory()
        self.d_sizes = torch.zeros((self.max_batch,), dtype=torch.int64, device=self.device)
        self.h_sizes = torch.zeros((self.max_batch,), dtype=torch.int64).pin_memory()
        self.first = min(self.size, FIRST_D2H_BYTES)

    def launch(self, bits, n, stream, copy_stream=None):
        """Enqueue the encode of n frames on `stream` and the read-back of the files on `copy_stream` (default: the same stream);
        finish() returns the files once that stream has drained."""
        if n > self.max_batch:
            raise ValueError("batch %d exceeds the encoder's capacity %d" % (n, self.max_batch))
        copy_stream = copy_stream or stream
        st = ctypes.c_void_p(stream.cuda_stream)
        if self.depth == 8:
            _lib.check(self.lib.am_png8_encode_deflate(bits.data_ptr(), n, self.height, self.width, self.d_out.data_ptr(),
                                                       self.d_sizes.data_ptr(), st), "am_png8_encode_deflate")
        elif self.compress:
            _lib.check(self.lib.am_png1_encode_deflate(bits.data_ptr(), n, self.height, self.width, self.d_out.data_ptr(),
                                                       self.d_sizes.data_ptr(), st), "am_png1_encode_deflate")
        else:
            _lib.check(self.lib.am_png1_encode(bits.data_ptr(), n, self.height, self.width, self.d_out.data_ptr(), st), "am_png1_encode")
        if copy_stream is not stream:
            if getattr(self, "_ev", None) is None:
                self._ev = torch.cuda.Event()
            self._ev.record(stream)
            copy_stream.wait_event(self._ev)
        with torch.cuda.stream(copy_stream):
            if self.compress:
                self.h_sizes[:n].copy_(self.d_sizes[:n], non_blocking=True)
                for f in range(n):                                       # contiguous row by row: torch turns a strided device -> host
                    self.h_out[f, :self.first].copy_(self.d_out[f, :self.first], non_blocking=True)   # copy into a SYNCHRONOUS one
            else:
                self.h_out[:n].copy_(self.d_out[:n], non_blocking=True)
        # finish() waits for THIS launch's read-back only: the stream may already hold the next batch's (whose encode has not run yet)
        if getattr(self, "_ev_done", None) is None:
            self._ev_done = torch.cuda.Event()
        self._ev_done.record(copy_stream)
        self._finish_stream = copy_stream

    def finish(self, n, stream=None):
        stream = stream or self._finish_stream
        self._ev_done.synchronize()
        if not self.compress:
            return [self.h_out[f].numpy().copy() for f in range(n)]
        sizes = [int(v) for v in self.h_sizes[:n]]
        big = [f for f in range(n) if sizes[f] > self.first]
        if big:                                                          # rare: a file beyond the first read-back
            with torch.cuda.stream(stream):
                for f in big:
                    self.h_out[f, self.first:sizes[f]].copy_(self.d_out[f, self.first:sizes[f]], non_blocking=True)
            stream.synchronize()
        return [self.h_out[f, :sizes[f]].numpy().copy() for f in range(n)]

    def encode(self, bits, n=None):
        """bits: bit-packed CUDA masks [n][H][WPR] -> list of n uint8 arrays (one PNG file each)."""
        n = bits.shape[0] if n is None else n
        st = torch.cuda.current_stream(self.device)
        self.launch(bits, n, st)
        return self.finish(n, st)


def encode_png_frames(bits, width, height, compress=True):
    """One-shot form of PngEncoder.encode."""
    return PngEncoder(width, height, bits.shape[0], bits.device, compress).encode(bits)
