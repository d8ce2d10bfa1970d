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


def test_header_is_plain_c_and_struct_sizes_match_the_ctypes_mirrors(tmp_path):
    """include/accessmath_b200.h must compile as C99 (the boundary is a C ABI), and the structs passed by pointer must have the
    layout the Python side assumes."""
    import ctypes
    import shutil
    import subprocess
    from lecturemath_b200.cc_grouping import UniqueView
    from lecturemath_b200.fcn_lecturenet import ConvDesc
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "accessmath_b200.h"\n'
                   'int main(void){ printf("%zu %zu\\n", sizeof(am_conv_desc), sizeof(am_unique_view)); return 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(ConvDesc), ctypes.sizeof(UniqueView)]
