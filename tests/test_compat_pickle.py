"""CPU: pickle interchange with the reference's stage CLIs (lecturemath_b200/compat.py).

The estimator must reach an UNMODIFIED reference stage 03 as the reference's own class
(AccessMath.preprocessing.content.cc_stability_estimator.CCStabilityEstimator), and files written by the reference must load
into this package's classes.  The third test runs the real reference in a subprocess when /root/reference exists (build
container only; skipped on the GPU box)."""
import json
import os
import pickle
import pickletools
import subprocess
import sys

import numpy as np
import pytest

from lecturemath_b200 import compat
from lecturemath_b200.cc_stability_estimator import CCStabilityEstimator
from lecturemath_b200.connected_component import ConnectedComponent, pack_crop
from oracle import cc_oracle as CO
from tests.conftest import REPO, unpack_masks

REF = "/root/reference/ACCESS2021_release"


def _host_estimator(golden, name="blobs_gap6"):
    """This package's estimator in the state stage 02 leaves it in, built without a device from the oracle's run."""
    zs = golden("cc_stability.npz")
    masks = unpack_masks(zs, name)
    r, p, gap = zs[name + "_params"]
    stab = CO.StabilityOracle(masks.shape[2], masks.shape[1], float(r), float(p), int(gap))
    for m in masks:
        stab.add_frame(m)
    est = CCStabilityEstimator.__new__(CCStabilityEstimator)
    conv = {}

    def cc_of(o):
        if id(o) not in conv:
            c = ConnectedComponent(o.cc_id, np.int32(o.min_x), np.int32(o.max_x), np.int32(o.min_y), np.int32(o.max_y), np.int32(o.size),
                                   packed=pack_crop(o.img, o.min_x, o.max_x, o.min_y, o.max_y))       # lazily materialised, like the GPU path
            c.start_time = c.end_time = np.float32(0.0)
            conv[id(o)] = c
        return conv[id(o)]
    est.width, est.height = stab.width, stab.height
    est.min_recall, est.min_precision, est.max_gap = float(r), float(p), int(gap)
    est.unique_cc_objects = [cc_of(o) for o in stab.unique_cc_objects]
    est.unique_cc_frames = [list(f) for f in stab.unique_cc_frames]
    est.cc_idx_per_frame = [[(u, cc_of(o)) for u, o in row] for row in stab.cc_idx_per_frame]
    est.fake_age, est.img_idx, est.tempo_count, est.verbose = None, len(masks), stab.tempo_count, False
    return est, stab


def test_dump_names_only_reference_classes(golden):
    est, _ = _host_estimator(golden)
    raw = compat.dumps_reference_pickle(est)
    names, strings = set(), []
    for op, arg, _ in pickletools.genops(raw):
        if op.name in ("SHORT_BINUNICODE", "BINUNICODE", "UNICODE"):
            strings.append(arg)
        elif op.name == "STACK_GLOBAL":
            names.add((strings[-2], strings[-1]))
        elif op.name == "GLOBAL":
            names.add(tuple(arg.split(" ")))
    assert {compat.EST_PATH, compat.CC_PATH, compat.IDX_PATH} <= names
    assert not [n for n in names if n[0].startswith("lecturemath_b200")], names
    assert b"lecturemath_b200" not in raw and b"_packed" not in raw


def test_reference_shaped_pickle_loads_into_this_package(golden):
    est, stab = _host_estimator(golden)
    raw = compat.dumps_reference_pickle(est)                # == what the reference's stage 02 writes (same paths, same attributes)
    aliased = compat.install_aliases()
    assert compat.EST_PATH[0] in aliased or compat._genuine(compat.EST_PATH) is not None
    if compat._genuine(compat.EST_PATH) is not None:
        pytest.skip("the genuine reference is importable in this process")
    back = pickle.loads(raw)
    assert type(back) is CCStabilityEstimator and type(back.unique_cc_objects[0]) is ConnectedComponent
    assert back.unique_cc_frames == est.unique_cc_frames and back.tempo_count == est.tempo_count and back.img_idx == est.img_idx
    assert back.get_raw_cc_count() == est.get_raw_cc_count()
    for a, b in zip(back.unique_cc_objects, stab.unique_cc_objects):
        np.testing.assert_array_equal(a.img, b.img)
        assert (a.cc_id, a.min_x, a.max_x, a.min_y, a.max_y, a.size) == (b.cc_id, b.min_x, b.max_x, b.min_y, b.max_y, b.size)
    # shared identity survives: the per-frame entry of a unique's first appearance IS the unique object
    u0, cc0 = back.cc_idx_per_frame[0][0]
    assert cc0 is back.unique_cc_objects[u0]
    last, active = compat.active_uniques(est)
    assert back.cc_last_frame == last and back.cc_active == active
    assert sorted(d for row in back.cc_int_index_x.intervals.values() for lst in row.values() for d in lst) == active


_REF_SCRIPT = r"""
import json, os, pickle, shutil, sys, tempfile, warnings
warnings.filterwarnings("ignore")
repo, ref, path = sys.argv[1:4]
tmp = tempfile.mkdtemp(prefix="amref_")
shutil.copy(os.path.join(repo, "oracle", "_ref", "accessmath_lib_ref.so"), os.path.join(tmp, "accessmath_lib.so"))
os.chdir(tmp)                                   # the reference CDLL-loads ./accessmath_lib.so at import time (labeler.py:24)
sys.path.insert(0, ref)
with open(path, "rb") as f:
    frame_times, frame_indices, est = pickle.load(f)          # stage 02's hand-off tuple (pre_ST3D_v3.0_02_cc_analaysis.py:43)
assert frame_indices == list(range(len(frame_times)))
import AccessMath.preprocessing.content.cc_stability_estimator as M
assert type(est) is M.CCStabilityEstimator and not hasattr(M, "__lecturemath_b200_alias__")
split_gap, min_times, t_window = (int(v) for v in sys.argv[4:7])
out = {"raw": est.get_raw_cc_count(), "split": est.split_stable_cc_by_gaps(split_gap, min_times)}
stable = est.get_stable_cc_idxs(min_times)
out["stable"] = [int(v) for v in stable]
t_ov, total, all_ov = est.compute_overlapping_stable_cc(stable, t_window)
out["n_time_overlaps"] = int(sum(len(v) for v in t_ov.values())) if isinstance(t_ov, dict) else int(sum(len(v) for v in t_ov))
est.add_frame(__import__("numpy").zeros((est.height, est.width), dtype="uint8"), True)      # the object is live: stage 02 can continue on it
out["img_idx"] = est.img_idx
print("RESULT " + json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_unmodified_reference_loads_and_runs_stage03_on_it(golden, tmp_path):
    name = "blobs_gap6"
    est, _ = _host_estimator(golden, name)
    zg = golden("cc_grouping.npz")
    split_gap, min_times, t_window, _, _ = zg[name + "/params"]
    path = tmp_path / "tempo_stability_test.dat"
    with open(path, "wb") as f:
        compat.dump_reference_pickle(([40.0 * t for t in range(est.img_idx)], list(range(est.img_idx)), est), f)
    script = tmp_path / "load_in_reference.py"
    script.write_text(_REF_SCRIPT)
    CO.build()
    r = subprocess.run([sys.executable, str(script), REPO, REF, str(path), str(int(split_gap)), str(int(min_times)), str(int(t_window))],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    assert out["raw"] == est.get_raw_cc_count()
    assert out["split"] == int(zg[name + "/split_count"])
    np.testing.assert_array_equal(np.array(out["stable"]), zg[name + "/stable"])
    assert out["img_idx"] == est.img_idx + 1
