"""CPU: the host ConnectedComponent value type (lecturemath_b200/connected_component.py, the type stage 03 / downstream stages hold)
against the oracle restatement of R/AM_CommonTools/data/connected_component.py:171-250 -- getOverlapFMeasure in both score forms,
getOverlapArea, box helpers -- on components extracted from seeded masks, including the bit-packed (lazy `img`) form the GPU path
creates and a pack -> unpack round trip at every x alignment."""
import numpy as np
import pytest

from lecturemath_b200.connected_component import ConnectedComponent, pack_crop, unpack_crop
from oracle import cc_oracle as CO


def _components(seed, h=96, w=160, density=0.55):
    rng = np.random.default_rng(seed)
    mask = (rng.random((h, w)) < density).astype(np.uint8) * 255
    comps, _, _ = CO.extract_components(mask, filter_small=False)
    return comps


def _host(o, packed):
    if packed:
        return ConnectedComponent(o.cc_id, o.min_x, o.max_x, o.min_y, o.max_y, o.size,
                                  packed=pack_crop(o.img, o.min_x, o.max_x, o.min_y, o.max_y))
    return ConnectedComponent(o.cc_id, o.min_x, o.max_x, o.min_y, o.max_y, o.size, img=o.img)


@pytest.mark.parametrize("packed", [False, True])
def test_overlap_measures_match_oracle(packed):
    a_list, b_list = _components(1), _components(2)
    big_a = sorted(a_list, key=lambda c: -c.size)[:40]
    big_b = sorted(b_list, key=lambda c: -c.size)[:40]
    n_overlapping = 0
    for oa in big_a:
        ha = _host(oa, packed)
        for ob in big_b:
            hb = _host(ob, packed)
            rec, prec = CO.overlap_measure(oa, ob)
            assert ha.getOverlapFMeasure(hb, False, False) == (rec, prec)                     # fp64, bit-exact (:239-240)
            match = rec * oa.size
            assert ha.getOverlapFMeasure(hb) == (2.0 * round(match)) / float(oa.size + ob.size)
            w = min(oa.max_x, ob.max_x) - max(oa.min_x, ob.min_x) + 1
            h = min(oa.max_y, ob.max_y) - max(oa.min_y, ob.min_y) + 1
            assert ha.getOverlapArea(hb) == (w * h if w > 0 and h > 0 else 0)
            n_overlapping += rec > 0
    assert n_overlapping >= 50                                                                 # the case set is not vacuous
    c = _host(big_a[0], packed)
    assert c.getBoxArea() == c.getWidth() * c.getHeight() == (big_a[0].max_x - big_a[0].min_x + 1) * (big_a[0].max_y - big_a[0].min_y + 1)
    assert c.getBoundingBox() == ((c.min_x, c.max_x), (c.min_y, c.max_y))


def test_disjoint_boxes_score_zero():
    img = np.full((3, 3), 255, np.uint8)
    a = ConnectedComponent(0, 0, 2, 0, 2, 9, img=img)
    b = ConnectedComponent(1, 3, 5, 0, 2, 9, img=img)                   # touching columns 2 | 3: boxes do not intersect
    assert a.getOverlapFMeasure(b) == 0.0 and a.getOverlapFMeasure(b, False, False) == (0.0, 0.0) and a.getOverlapArea(b) == 0
    c = ConnectedComponent(2, 2, 4, 2, 4, 9, img=img)                   # one shared pixel (2, 2)
    assert a.getOverlapFMeasure(c, False, False) == (1 / 9.0, 1 / 9.0) and a.getOverlapArea(c) == 1


@pytest.mark.parametrize("x0", [0, 1, 31, 32, 33, 63, 95])
def test_pack_unpack_round_trip(x0):
    rng = np.random.default_rng(x0)
    for w in (1, 2, 31, 32, 33, 70):
        img = (rng.random((5, w)) < 0.5).astype(np.uint8) * 255
        words = pack_crop(img, x0, x0 + w - 1, 7, 11)
        assert len(words) == 5 * (((x0 + w - 1) >> 5) - (x0 >> 5) + 1)
        np.testing.assert_array_equal(unpack_crop(words, x0, x0 + w - 1, 7, 11), img)
