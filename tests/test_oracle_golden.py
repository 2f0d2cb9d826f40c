"""CPU: the plain-C oracle against the committed fixtures generated from the reference's own code
(tests/golden/make_golden.py) and against the known answers of SURVEY.md appendix B."""
import numpy as np

import oracle_bindings as ob
from common import (FILT_OFF, FILT_STRIDE, filter_case_iter, filter_src, fme, golden, golden_recs, golden_res,
                    nn_fields, std_fields)


def test_filters_match_reference_fixture(orc):
    g = golden()
    n = 0
    for (is_ver, luma, frac, w, h, first, last, bd, want) in filter_case_iter(g):
        src = filter_src(g, is_ver, first, bd)
        if is_ver:
            got = orc.filter_ver(luma, src, FILT_OFF, FILT_STRIDE, w, h, frac, first, last, bd)
        else:
            got = orc.filter_hor(luma, src, FILT_OFF, FILT_STRIDE, w, h, frac, last, bd)
        assert np.array_equal(got, want), (is_ver, luma, frac, w, h, first, last, bd)
        n += 1
    assert n == 432


def test_distortions_match_reference_fixture(orc):
    g = golden()
    for (w, h, kind, ss, pair), want in zip(g["dist_meta"], g["dist_val"]):
        o, c = (g["dist_org"], g["dist_cur"]) if pair == 0 else (g["dist_org2"], g["dist_cur2"])
        got = orc.dist(int(kind), o, 0, 64, c, 0, 80, int(w), int(h), 8, int(ss))
        assert got == int(want), (w, h, kind, ss, pair)


def test_add_avg_matches_reference_fixture(orc):
    """TComYuv::addAvg (bi-prediction average of 14-bit intermediates), outputs of the compiled reference."""
    g = golden()
    a, b, pos = g["avg_a"], g["avg_b"], 0
    for (w, h) in g["avg_shapes"]:
        want = g["avg_out"][pos:pos + w * h].reshape(h, w)
        pos += w * h
        assert np.array_equal(orc.add_avg(a, 0, 64, b, 0, 64, int(w), int(h)), want), (w, h)


def test_mv_cost_matches_reference_fixture(orc):
    g = golden()
    for (lam, x, y, sc, px, py), want in zip(g["mv_meta"], g["mv_val"]):
        assert orc.mv_cost(float(lam), int(x), int(y), int(sc), int(px), int(py)) == int(want)


def test_frac_dif_small_frame(orc):
    g = golden()
    recs = golden_recs(g)
    frame = ob.CpuFrame(g["small_org"], list(g["small_refs"]), margin=80)
    lam = float(g["small_lambda"][0])
    for key, had in (("small_res_had", 1), ("small_res_sad", 0)):
        got = frame.oracle_run(recs, 1, lam, had, None)
        mv_g, cost_g = std_fields(got)
        mv_w, cost_w = std_fields(golden_res(g, key))
        assert np.array_equal(mv_g, mv_w), key
        assert np.array_equal(cost_g, cost_w), key


def test_int_surface_small_frame(orc):
    g = golden()
    recs = golden_recs(g).copy()
    want = recs["err"].copy()
    recs["err"] = 0
    frame = ob.CpuFrame(g["small_org"], list(g["small_refs"]), margin=80)
    frame.oracle_fill_surface(recs, fen=1)
    assert np.array_equal(recs["err"], want)


def test_nn_pred_small_frame_and_grids(orc):
    g = golden()
    recs = golden_recs(g)
    frame = ob.CpuFrame(g["small_org"], list(g["small_refs"]), margin=80)
    for qp in (22, 27, 32, 37):
        blob = fme.nn_weights.load_blob(qp)
        got = frame.oracle_run(recs, 2, 1.0, 1, blob)
        assert np.array_equal(nn_fields(got), nn_fields(golden_res(g, "small_res_nn%d" % qp))), qp
        for row, want in zip(g["nn_grids"], g["nn_out%d" % qp]):
            cls, _, hxy, qxy = orc.nn_pred(blob, row[:9], int(row[9]), int(row[10]))
            assert [cls, hxy[0], hxy[1], qxy[0], qxy[1]] == list(want), (qp, row)


def test_survey_appendix_b_known_answers(orc):
    """formula-defined plane, pred (5,-3), lambda(QP22+3): half=(1,1), qter=(1,0)."""
    S = 256
    x = np.arange(1024)[None, :]
    y = np.arange(1024)[:, None]
    raw = ((37 * x + 101 * y + ((x * y) >> 3) + 13 * ((x >> 2) ^ (y >> 3))) & 255).astype(np.int64)
    X = np.arange(S)
    ref = np.zeros((S, S), np.int64)
    for j in (-1, 0, 1):
        for i in (-1, 0, 1):
            ref += raw[np.ix_((X + j) & 1023, (X + i) & 1023)]
    ref = (ref // 9).astype(np.int16)
    lam = 0.4624 * 2 ** ((22 + 3 - 12) / 3)
    want = {(64, 64): 25129, (32, 32): 6731, (16, 16): 1713, (8, 8): 467, (16, 4): 503, (12, 16): 1303, (8, 4): 277,
            (4, 8): 257, (32, 24): 5139, (64, 48): 19056}
    for (w, h), cost in want.items():
        a = ref[96:96 + h, 96:96 + w].astype(np.int32); b = ref[96:96 + h, 97:97 + w].astype(np.int32)
        c = ref[97:97 + h, 96:96 + w].astype(np.int32); d = ref[97:97 + h, 97:97 + w].astype(np.int32)
        org = np.zeros((64, 64), np.int16)
        org[:h, :w] = ((a + 3 * b) + (c + 3 * d) + 4) >> 3
        half, qter, got = orc.frac_dif(org, 0, 64, w, h, ref, 96 * S + 96, S, 0, 0, 5, -3, lam, 1)
        assert (half, qter, got) == ((1, 1), (1, 0), cost), (w, h)


def test_nn_blob_matches_reference_tables():
    """Appendix B class table through the shipped blobs."""
    orc = ob.oracle()
    grids = [([1200, 900, 1300, 800, 500, 850, 1250, 950, 1400], 8, 8, (24, 24, 24, 24)),
             ([250000, 180000, 240000, 150000, 60000, 90000, 230000, 120000, 200000], 32, 32, (32, 32, 32, 32)),
             ([5000, 5200, 5100, 4800, 4700, 4900, 5300, 5250, 5400], 4, 8, (16, 22, 24, 24)),
             ([800000, 500000, 700000, 300000, 100000, 350000, 750000, 450000, 820000], 64, 64, (23, 23, 24, 24)),
             ([30000, 20000, 25000, 15000, 7000, 12000, 28000, 18000, 26000], 12, 16, (25, 24, 24, 24)),
             ([30000, 20000, 25000, 15000, 7000, 12000, 28000, 18000, 26000], 16, 12, (25, 25, 25, 25))]
    for e, h, w, want in grids:
        for qp, cls in zip((22, 27, 32, 37), want):
            assert orc.nn_pred(fme.nn_weights.load_blob(qp), e, h, w)[0] == cls, (e, h, w, qp)
