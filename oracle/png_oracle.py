"""oracle/png_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle for the device PNG writer).

The 01 -> 02 wire format is a list of PNG byte arrays that Helper.decompress_binary_images reads back with
cv2.imdecode(raw, IMREAD_GRAYSCALE) (R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:56-64,
R/AccessMath/preprocessing/content/helper.py:27-34).  Any valid PNG that decodes to the same 0/255 pixels is a drop-in.  png1()
restates the container am_png1_encode writes -- 1-bit grayscale, filter 0, zlib "stored" blocks (PNG / RFC 1950 / RFC 1951 are
the published specifications; zlib.crc32 / zlib.adler32 are the checksum references) -- so the device output can be compared
BYTE for byte, and decode() is the reference-side reader.  Only tests/ may import this module."""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89PNG\r\n\x1a\n"


def _chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data))


def png1(mask):
    """uint8 (H, W) mask (nonzero = ink = white) -> bytes of a 1-bit grayscale PNG with stored deflate blocks."""
    h, w = mask.shape
    rows = np.packbits(np.asarray(mask) != 0, axis=1, bitorder="big")            # PNG packs the leftmost pixel into the MSB
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rows], axis=1).tobytes()   # filter type 0 in front of every scanline
    z, pos = b"\x78\x01", 0
    while pos < len(raw):
        blk = raw[pos:pos + 65535]
        pos += len(blk)
        z += struct.pack("<BHH", 1 if pos >= len(raw) else 0, len(blk), len(blk) ^ 0xFFFF) + blk
    z += struct.pack(">I", zlib.adler32(raw))
    return SIGNATURE + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 1, 0, 0, 0, 0)) + _chunk(b"IDAT", z) + _chunk(b"IEND", b"")


def size(width, height):
    raw = height * (1 + (width + 7) // 8)
    return 8 + 25 + 12 + (2 + raw + 5 * ((raw + 65534) // 65535) + 4) + 12


def decode(raw_png):
    """What the reference does with an entry of compressed_frames (helper.py:31)."""
    import cv2
    return cv2.imdecode(np.frombuffer(bytes(raw_png), np.uint8), cv2.IMREAD_GRAYSCALE)
