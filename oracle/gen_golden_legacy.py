"""oracle/gen_golden_legacy.py -- TEST INFRASTRUCTURE.  Regenerates tests/golden/legacy_ops.npz by calling the
reference's own C (oracle/_ref/accessmath_lib_ref.so = R/accessmath_lib.c compiled unmodified, see oracle/Makefile)
on seeded inputs.  Run in the build container:   python oracle/gen_golden_legacy.py"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import legacy_oracle as L  # noqa: E402


def legacy_inputs():
    """name -> dict of seeded inputs (shared with the tests, which regenerate them instead of storing them)."""
    rng = np.random.default_rng(20211)
    out = {}

    def board(h, w, lo=40, hi=235):
        yy, xx = np.mgrid[0:h, 0:w]
        g = hi - 60.0 * xx / max(w - 1, 1) - 35.0 * yy / max(h - 1, 1) + rng.normal(0, 6, (h, w))
        ink = rng.random((h, w)) < 0.06
        g = np.where(ink, lo + rng.normal(0, 10, (h, w)), g)
        return np.clip(g, 0, 255).astype(np.uint8)

    out["ahe_97x131_g8x8"] = dict(gray=board(97, 131), slope=0.04, gx=8, gy=8)
    out["ahe_64x64_g3x5"] = dict(gray=board(64, 64), slope=0.04, gx=3, gy=5)
    out["ahe_50x77_noclip"] = dict(gray=board(50, 77), slope=0.0, gx=4, gy=4)
    out["ahe_40x40_bigslope"] = dict(gray=board(40, 40), slope=0.9, gx=2, gy=2)
    out["ahe_33x45_const"] = dict(gray=np.full((33, 45), 200, np.uint8), slope=0.04, gx=4, gy=3)
    out["ahe_30x30_g1x1"] = dict(gray=board(30, 30), slope=0.04, gx=1, gy=1)
    out["comb_61x95"] = dict(board=rng.integers(0, 256, (61, 95), dtype=np.uint8), eq=rng.integers(0, 256, (61, 95), dtype=np.uint8), thr=97)
    out["comb_64x64"] = dict(board=rng.integers(120, 136, (64, 64), dtype=np.uint8), eq=rng.integers(0, 256, (64, 64), dtype=np.uint8), thr=128)
    a = rng.integers(0, 256, (54, 96, 3), dtype=np.uint8)
    b = a.copy(); b[10:31, 20:70] = rng.integers(0, 256, (21, 50, 3), dtype=np.uint8)
    out["spk_54x96x3_j1"] = dict(frame=b, last=a, thr=30, jump=1)
    out["spk_54x96x3_j3"] = dict(frame=b, last=a, thr=30, jump=3)
    out["spk_54x96x3_same"] = dict(frame=a, last=a, thr=30, jump=2)
    g = rng.integers(0, 256, (45, 301), dtype=np.uint8)
    g2 = np.clip(g.astype(np.int32) + rng.integers(-60, 60, g.shape), 0, 255).astype(np.uint8)
    out["spk_45x301x1_j2"] = dict(frame=g2, last=g, thr=25, jump=2)
    return out


def main():
    lib = L.ref()
    assert lib is not None, "build oracle/_ref first (make -C oracle)"
    z = {}
    for name, d in legacy_inputs().items():
        if name.startswith("ahe"):
            z[name] = L.ref_adapthisteq(lib, d["gray"], d["slope"], d["gx"], d["gy"])
            h, w = d["gray"].shape
            z[name + "_cdf"] = L.ref_region_cdf(lib, d["gray"], w // 5, w - 3, h // 4, h - 2, d["slope"])
        elif name.startswith("comb"):
            z[name] = L.ref_combine_results(lib, d["board"], d["eq"], d["thr"])
        else:
            t, b, a, dv = L.ref_speaker_detection(lib, d["frame"], d["last"], d["thr"], d["jump"])
            z[name] = np.concatenate([b, a, dv, [float(t)]])
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "legacy_ops.npz"), **z)
    print("wrote", len(z), "arrays")


if __name__ == "__main__":
    main()
