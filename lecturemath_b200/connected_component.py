"""ConnectedComponent value type, mirroring R/AM_CommonTools/data/connected_component.py:21-41 for the fields
and methods the hot path and stage 03 read (cc_id, min/max x/y, size, img, start_time/end_time, getBoxArea,
getWidth/getHeight, getOverlapArea, getOverlapFMeasure).  `img` (uint8 h x w, 0/255, labeler.py:183) is
materialised lazily from the bit-packed, word-aligned crop the GPU produced."""
import numpy as np


def unpack_crop(words, min_x, max_x, min_y, max_y):
    """bit-packed word-aligned crop (uint32[h*cw]) -> uint8 (h, w) 0/255."""
    h = max_y - min_y + 1
    cw = (max_x >> 5) - (min_x >> 5) + 1
    bits = np.unpackbits(np.ascontiguousarray(words, dtype="<u4").view(np.uint8).reshape(h, cw * 4), axis=1, bitorder="little")
    x0 = min_x - ((min_x >> 5) << 5)
    return np.ascontiguousarray(bits[:, x0:x0 + (max_x - min_x + 1)]) * np.uint8(255)


def pack_crop(img, min_x, max_x, min_y, max_y):
    """uint8 (h, w) crop (nonzero = ink) -> bit-packed word-aligned uint32[h*cw], bits at their absolute x position
    (the layout K-crop writes and K-match ANDs, csrc/cc_kernels.cu)."""
    h = max_y - min_y + 1
    cw = (max_x >> 5) - (min_x >> 5) + 1
    bits = np.zeros((h, cw * 32), dtype=np.uint8)
    x0 = min_x - ((min_x >> 5) << 5)
    bits[:, x0:x0 + (max_x - min_x + 1)] = (np.asarray(img) != 0)
    return np.packbits(bits, axis=1, bitorder="little").view("<u4").reshape(-1).copy()


class ConnectedComponent:
    def __init__(self, cc_id, min_x, max_x, min_y, max_y, size, img=None, packed=None):
        self.cc_id = cc_id
        self.min_x, self.max_x, self.min_y, self.max_y = min_x, max_x, min_y, max_y
        self.size = size
        self._img = img
        self._packed = packed            # uint32 words (host) of the bit-packed crop
        self.normalized = None
        self.start_time = None
        self.end_time = None
        self.next_cc = None
        self.prev_cc = None

    def __setstate__(self, state):
        """Also accepts the attribute dictionary of a reference ConnectedComponent (plain `img` attribute), so that files
        written by the reference's stage 02 load into this class (lecturemath_b200/compat.py)."""
        state = dict(state)
        if "img" in state:
            state["_img"] = state.pop("img")
        state.setdefault("_img", None)
        state.setdefault("_packed", None)
        self.__dict__.update(state)

    @property
    def img(self):
        if self._img is None and self._packed is not None:
            self._img = unpack_crop(self._packed, self.min_x, self.max_x, self.min_y, self.max_y)
        return self._img

    @img.setter
    def img(self, value):
        self._img = value

    def getBoundingBox(self):
        return (self.min_x, self.max_x), (self.min_y, self.max_y)

    def getWidth(self):
        return self.max_x - self.min_x + 1

    def getHeight(self):
        return self.max_y - self.min_y + 1

    def getBoxArea(self):
        return (self.max_x - self.min_x + 1) * (self.max_y - self.min_y + 1)

    def getOverlapArea(self, other):
        if (self.min_x <= other.max_x and other.min_x <= self.max_x) and (self.min_y <= other.max_y and other.min_y <= self.max_y):
            w = min(self.max_x, other.max_x) - max(self.min_x, other.min_x) + 1
            h = min(self.max_y, other.max_y) - max(self.min_y, other.min_y) + 1
            return w * h
        return 0

    def getOverlapFMeasure(self, other, verbose=False, single_score=True):
        """Host-side convenience for downstream stages (the hot path does this on the GPU, K-match)."""
        if not (self.max_y >= other.min_y and other.max_y >= self.min_y and self.max_x >= other.min_x and other.max_x >= self.min_x):
            return 0.0 if single_score else (0.0, 0.0)
        x0, x1 = max(self.min_x, other.min_x), min(self.max_x, other.max_x)
        y0, y1 = max(self.min_y, other.min_y), min(self.max_y, other.max_y)
        a = self.img[y0 - self.min_y:y1 - self.min_y + 1, x0 - self.min_x:x1 - self.min_x + 1]
        b = other.img[y0 - other.min_y:y1 - other.min_y + 1, x0 - other.min_x:x1 - other.min_x + 1]
        match = int(np.count_nonzero(np.bitwise_and(a, b)))
        if single_score:
            return (2.0 * match) / float(self.size + other.size)
        return match / float(self.size), match / float(other.size)

    def __str__(self):
        return "ConnectedComponent -> Id = %s\n -> X : [%s, %s] \n -> Y : [%s, %s]" % (self.cc_id, self.min_x, self.max_x, self.min_y, self.max_y)
