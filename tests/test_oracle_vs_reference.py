"""CPU: randomised comparison of the plain-C oracle with the reference's own compiled objects
(oracle/_ref/libhmref.so).  Skipped where neither the prebuilt library nor /root/reference exists."""
import numpy as np
import pytest

import oracle_bindings as ob
from common import fme

SHAPES = [(64, 64), (32, 32), (16, 16), (8, 8), (16, 4), (12, 16), (8, 4), (4, 8), (32, 24), (64, 48), (16, 12), (4, 16),
          (24, 32), (48, 64), (64, 16), (16, 64), (32, 8), (8, 32), (64, 32), (32, 64), (16, 8), (8, 16), (32, 16), (16, 32)]


@pytest.fixture(scope="module")
def refl(ref):
    if ref is None:
        pytest.skip("reference library not available")
    return ref


def test_filters_random(orc, refl):
    rng = np.random.default_rng(1)
    for bd in (8, 10, 12):
        src = rng.integers(0, 1 << bd, (40, 48)).astype(np.int16)
        inter = rng.integers(-14312, 14248, (40, 48)).astype(np.int16)
        off = 10 * 48 + 10
        for luma in (1, 0):
            for frac in range(4 if luma else 8):
                for w, h in ((8, 8), (5, 9), (17, 12)):
                    for last in (0, 1):
                        assert np.array_equal(orc.filter_hor(luma, src, off, 48, w, h, frac, last, bd),
                                              refl.filter_hor(luma, src, off, 48, w, h, frac, last, bd))
                        for first in (0, 1):
                            s = src if first else inter
                            assert np.array_equal(orc.filter_ver(luma, s, off, 48, w, h, frac, first, last, bd),
                                                  refl.filter_ver(luma, s, off, 48, w, h, frac, first, last, bd))


def test_add_avg_random(orc, refl):
    rng = np.random.default_rng(12)
    refl.init(22, 1, 1)
    for (w, h) in SHAPES + [(2, 4), (4, 2), (6, 8), (32, 32)]:
        a = rng.integers(-14312, 14249, (64, 72)).astype(np.int16)
        b = rng.integers(-14312, 14249, (64, 80)).astype(np.int16)
        assert np.array_equal(orc.add_avg(a, 3, 72, b, 5, 80, w, h), refl.add_avg(a, 3, 72, b, 5, 80, w, h)), (w, h)


def test_dist_random(orc, refl):
    rng = np.random.default_rng(2)
    refl.init(22, 1, 1)
    for (w, h) in SHAPES:
        org = rng.integers(0, 256, (64, 64)).astype(np.int16)
        cur = rng.integers(0, 256, (80, 80)).astype(np.int16)
        for kind in (0, 1, 2):
            for ss in ((0, 1) if kind != 1 else (0,)):
                assert orc.dist(kind, org, 0, 64, cur, 81, 80, w, h, 8, ss) == refl.dist(kind, org, 0, 64, cur, 81, 80, w, h, 8, ss)


def test_frac_dif_random(orc, refl):
    rng = np.random.default_rng(3)
    S = 160
    for t in range(240):
        ref = rng.integers(0, 256, (S, S)).astype(np.int32)
        ref = ((ref + np.roll(ref, 1, 0) + np.roll(ref, 1, 1) + np.roll(ref, -1, 0)) // 4).astype(np.int16)
        w, h = SHAPES[t % len(SHAPES)]
        org = np.zeros((64, 64), np.int16)
        dx, dy = rng.integers(0, 2, 2)
        org[:h, :w] = np.clip((ref[40 + dy:40 + dy + h, 40 + dx:40 + dx + w].astype(np.int32) + ref[40:40 + h, 40:40 + w]) // 2
                              + rng.integers(-3, 4, (h, w)), 0, 255)
        lam = float(rng.uniform(4, 120))
        had = t % 5 != 0
        lossless = int(t % 11 == 0)
        refl.init(22, int(had), 1)
        refl.set_lambda(lam)
        mvx, mvy = [int(v) for v in rng.integers(-2, 3, 2)]
        px, py = [int(v) for v in rng.integers(-20, 20, 2)]
        a = orc.frac_dif(org, 0, 64, w, h, ref, 40 * S + 40, S, mvx, mvy, px, py, lam, int(had), lossless)
        b = refl.frac_dif(org, 0, 64, w, h, ref, 40 * S + 40, S, mvx, mvy, px, py, lossless)
        assert a == b, (w, h, had, lossless)
        # the scratch planes the reference leaves behind (m_filteredBlock) equal the oracle's
        for (v, hh) in ((0, 0), (2, 0), (0, 2), (2, 2), (1, 1), (3, 3)):
            assert np.array_equal(orc.filtered_block(v, hh, w, h), refl.filtered_block(v, hh, w, h))


def test_nn_pred_random(orc, refl):
    rng = np.random.default_rng(4)
    for qp in (22, 27, 32, 37, 30):
        refl.init(qp, 1, 1)
        blob = fme.nn_weights.load_blob(qp)
        for t in range(300):
            base = rng.uniform(100, 3e5)
            e = (base * rng.uniform(0.2, 3, 9)).astype(np.uint32)
            if t % 7 == 0:
                e[:] = rng.integers(0, 2 ** 31, 9)
            hh = int(rng.choice([4, 8, 12, 16, 24, 32, 64, 48]))
            ww = int(rng.choice([4, 8, 12, 16, 24, 32, 64, 48]))
            a = orc.nn_pred(blob, e, hh, ww)
            assert (a[0], a[2], a[3]) == refl.nn_pred(e, hh, ww)


def test_subpel_plane_equals_search_scratch(orc, refl):
    """the whole-plane definition P[fy][fx] equals what the per-PU path leaves in m_filteredBlock."""
    rng = np.random.default_rng(5)
    S = 128
    ref = rng.integers(0, 256, (S, S)).astype(np.int16)
    org = rng.integers(0, 256, (64, 64)).astype(np.int16)
    refl.init(22, 1, 1)
    refl.set_lambda(10.0)
    w = h = 16
    refl.frac_dif(org, 0, 64, w, h, ref, 40 * S + 40, S, 0, 0, 0, 0)
    # half planes are generated from (x-1, y-1): m_filteredBlock[2][2] holds P[2][2] at origin (39, 39)
    got = refl.filtered_block(2, 2, w + 1, h + 1)
    want = orc.subpel_plane(ref, 0, S, 39, 39, w + 1, h + 1, 2, 2)
    assert np.array_equal(got, want)
    got = refl.filtered_block(0, 2, w + 1, h)
    assert np.array_equal(got, orc.subpel_plane(ref, 0, S, 39, 40, w + 1, h, 0, 2))
    got = refl.filtered_block(2, 0, w, h + 1)
    assert np.array_equal(got, orc.subpel_plane(ref, 0, S, 40, 39, w, h + 1, 2, 0))
