"""Helper.decompress_binary_images (R/AccessMath/preprocessing/content/helper.py:27-34): the decode side of the 01 -> 02 wire
format (one PNG per frame, written by FCN_LectureNet_Binarizer.handleFrame)."""
import concurrent.futures
import os

import cv2
import numpy as np

from .packed_mask import PackedMask, parse_png1


class Helper:
    @staticmethod
    def decompress_binary_images(compressed_images, lazy=True):
        """-> list of frames, one per entry.  The reference returns `cv2.imdecode(raw, IMREAD_GRAYSCALE)` arrays; here the 1-bit PNGs
        this package's binarizer writes come back as PackedMask -- same `.shape`, same pixels through np.asarray / indexing, but
        still bit-packed, so CCStabilityEstimator.add_frame uploads 1/8 of the bytes and nothing is unpacked on the host.  Any
        other PNG (e.g. written by the reference's cv2.imencode), or lazy=False, takes the reference's path."""
        def one(raw):
            m = parse_png1(raw) if lazy else None
            return m if m is not None else cv2.imdecode(np.asarray(raw), cv2.IMREAD_GRAYSCALE)

        if len(compressed_images) < 32:
            return [one(raw) for raw in compressed_images]
        # inflating is the whole cost (0.2-0.4 ms per 1080p frame) and zlib / cv2 release the GIL: a few threads, order kept
        with concurrent.futures.ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as pool:
            return list(pool.map(one, compressed_images, chunksize=16))


__all__ = ["Helper", "PackedMask"]
