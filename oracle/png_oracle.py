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


# ---- the compressed container am_png1_encode_deflate writes (csrc/png.cu: k_png1_deflate) ------------------------------------------
# RFC 1951 fixed Huffman codes, matches of distance 1 only (run lengths).  Raw scanline bytes are cut into warp-segments of 8192 bytes
# = one deflate block each, and every block into 32 lane ranges of 256 bytes that are tokenised independently (runs stop at lane
# boundaries); a fixed block is followed by an empty stored block so that the next one starts on a byte boundary.
LANE_BYTES, SEG_BYTES = 256, 32 * 256


class _Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, value, nbits):                              # LSB first
        self.acc |= (value & ((1 << nbits) - 1)) << self.n
        self.n += nbits

    def align(self):
        self.n = (self.n + 7) // 8 * 8

    def tobytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def _fixed_code(sym):
    """(code with its bits reversed for the LSB-first stream, length), RFC 1951 3.2.6."""
    if sym < 144:
        code, n = 0x30 + sym, 8
    elif sym < 256:
        code, n = 0x190 + (sym - 144), 9
    elif sym < 280:
        code, n = sym - 256, 7
    else:
        code, n = 0xC0 + (sym - 280), 8
    return int(format(code, "0%db" % n)[::-1], 2), n


_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


def _length_symbol(length):
    """match length 3 .. 258 -> (symbol, extra bits, extra value): the table of RFC 1951 3.2.5."""
    for i in range(len(_LEN_BASE) - 1, -1, -1):
        if length >= _LEN_BASE[i]:
            return 257 + i, _LEN_EXTRA[i], length - _LEN_BASE[i]
    raise ValueError(length)


def _tokens(out, data):
    i, n = 0, len(data)
    while i < n:
        j = i + 1
        while j < n and data[j] == data[i]:
            j += 1
        run = j - i
        lc, ln = _fixed_code(data[i])
        if run >= 4:                                          # literal + (length run - 1, distance 1)
            sym, eb, ev = _length_symbol(run - 1)
            sc, sn = _fixed_code(sym)
            out.put(lc, ln); out.put(sc, sn); out.put(ev, eb); out.put(0, 5)
        else:
            for _ in range(run):
                out.put(lc, ln)
        i = j


def png1_deflate(mask):
    """uint8 (H, W) mask -> bytes of the 1-bit grayscale PNG with the compressed zlib stream, byte for byte what the device writes."""
    h, w = mask.shape
    rows = np.packbits(np.asarray(mask) != 0, axis=1, bitorder="big")
    return _deflate_png(np.concatenate([np.zeros((h, 1), np.uint8), rows], axis=1).tobytes(), w, h, 1)


def png8_deflate(frame):
    """uint8 (H, W) grayscale frame -> the 8-bit PNG am_png8_encode_deflate writes (same block structure, one byte per pixel)."""
    h, w = frame.shape
    return _deflate_png(np.concatenate([np.zeros((h, 1), np.uint8), np.asarray(frame, dtype=np.uint8)], axis=1).tobytes(), w, h, 8)


def _deflate_png(raw, w, h, depth):
    n_seg = (len(raw) + SEG_BYTES - 1) // SEG_BYTES
    z = b"\x78\x01"
    for s in range(n_seg):
        last = s == n_seg - 1
        b = _Bits()
        b.put(3 if last else 2, 3)                            # BFINAL, BTYPE = 01
        for lane in range(32):
            r0 = min(len(raw), s * SEG_BYTES + lane * LANE_BYTES)
            _tokens(b, raw[r0:min(len(raw), r0 + LANE_BYTES)])
        b.put(0, 7)                                           # end of block
        if not last:
            b.put(0, 3)                                       # empty stored block: header, pad, LEN = 0, NLEN = 0xFFFF
            b.align()
            b.put(0x0000, 16); b.put(0xFFFF, 16)
        z += b.tobytes()
    z += struct.pack(">I", zlib.adler32(raw))
    return SIGNATURE + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, 0, 0, 0, 0)) + _chunk(b"IDAT", z) + _chunk(b"IEND", b"")


def capacity(width, height, depth=1):
    raw = height * (1 + (width if depth == 8 else (width + 7) // 8))
    n_seg = (raw + SEG_BYTES - 1) // SEG_BYTES
    return (8 + 25 + 12 + (2 + (raw * 9 + 7) // 8 + 8 * n_seg + 4) + 12 + 15) // 16 * 16
