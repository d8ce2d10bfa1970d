"""CPU: the C-ABI library builds, loads and exports every symbol include/accessmath_b200.h declares.
No compute calls here (there is no GPU on the CPU tier)."""
import os
import re

import pytest

from lecturemath_b200 import _lib, build as B

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(REPO, "include", "accessmath_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    legacy = "adapthisteq|regionCumulativeDistribution|combine_results|speaker_detection_handle_frame"
    return sorted(set(re.findall(r"\b((?:am_|CC_)\w+|%s)\s*\(" % legacy, src)))


def test_library_exports_every_declared_symbol():
    B.build()
    lib = _lib.load()
    names = declared_symbols()
    assert "CC_AgeBoundaries" in names and len(names) >= 20
    for legacy in ("adapthisteq", "regionCumulativeDistribution", "combine_results", "speaker_detection_handle_frame"):
        assert legacy in names          # the drop-in .so exports all five reference symbols
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_lib.SIGNATURES) == names


def test_no_device_means_loud_failure():
    lib = _lib.load()
    if lib.am_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.AccessMathB200Error):
        _lib.lib()


def test_words_per_row():
    lib = _lib.load()
    assert lib.am_words_per_row(1920) == 60 and lib.am_words_per_row(1280) == 40
    assert lib.am_words_per_row(121) == 4 and lib.am_words_per_row(1) == 4 and lib.am_words_per_row(129) == 8
