"""Helper.decompress_binary_images (R/AccessMath/preprocessing/content/helper.py:27-34): the decode side of the
01 -> 02 wire format (PNG per frame).  Wire-format code, not accelerated in this round."""
import cv2


class Helper:
    @staticmethod
    def decompress_binary_images(compressed_images):
        return [cv2.imdecode(raw, cv2.IMREAD_GRAYSCALE) for raw in compressed_images]
