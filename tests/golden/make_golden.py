"""Generate tests/golden/fme_golden.npz from the REFERENCE's own compiled code (oracle/_ref/libhmref.so).

Run in the dev container (needs /root/reference to build libhmref.so):
    python tests/golden/make_golden.py
The fixture stores inputs AND the reference's outputs, so the tests that consume it need neither the
reference tree nor a particular numpy RNG stream.  Everything in it was produced by the reference's
compiled objects: filterHor/filterVer, DistFunc (SSE/SAD/HADs), getCostOfVectorWithPredictor,
xPatternSearchFracDIF and NN_pred (the latter over oracle/eigen_standin, see its header).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fme_loader  # noqa: E402
import oracle_bindings as ob  # noqa: E402

fme = fme_loader.load()
R = ob.reference()
assert R is not None, "reference library unavailable"
rng = np.random.default_rng(20261018)
out = {}

# ---- interpolation filter cases ---------------------------------------------------------------
fcases = []
srcs8 = rng.integers(0, 256, (36, 40)).astype(np.int16)
srcs10 = rng.integers(0, 1024, (36, 40)).astype(np.int16)
inter = rng.integers(-14312, 14249, (36, 40)).astype(np.int16)
srcs8[0:4, :] = 255; srcs8[4:8, :] = 0  # saturating rows
out["filt_src8"], out["filt_src10"], out["filt_inter"] = srcs8, srcs10, inter
OFF = 8 * 40 + 8
for bd, src in ((8, srcs8), (10, srcs10)):
    for luma in (1, 0):
        for frac in range(4 if luma else 8):
            for (w, h) in ((8, 8), (13, 5), (24, 16)):
                for last in (0, 1):
                    fcases.append((0, luma, frac, w, h, 1, last, bd, R.filter_hor(luma, src, OFF, 40, w, h, frac, last, bd)))
                    for first in (0, 1):
                        s = src if first else inter
                        fcases.append((1, luma, frac, w, h, first, last, bd,
                                       R.filter_ver(luma, s, OFF, 40, w, h, frac, first, last, bd)))
out["filt_meta"] = np.array([c[:8] for c in fcases], np.int32)
out["filt_out"] = np.concatenate([c[8].ravel() for c in fcases]).astype(np.int16)

# ---- distortion cases -----------------------------------------------------------------------------
R.init(22, 1, 1)
shapes = [(4, 8), (8, 4), (8, 8), (16, 8), (8, 16), (16, 16), (12, 16), (16, 12), (16, 4), (4, 16), (32, 32), (24, 32),
          (32, 24), (32, 8), (8, 32), (64, 64), (48, 64), (64, 48), (64, 16), (16, 64), (64, 32), (32, 64)]
dorg = rng.integers(0, 256, (64, 64)).astype(np.int16)
dcur = rng.integers(0, 256, (64, 80)).astype(np.int16)
dorg2 = np.where(rng.integers(0, 2, (64, 64)) > 0, 255, 0).astype(np.int16)  # worst-case magnitudes
dcur2 = np.ascontiguousarray(np.pad(255 - dorg2, ((0, 0), (0, 16))))
out["dist_org"], out["dist_cur"], out["dist_org2"], out["dist_cur2"] = dorg, dcur, dorg2, dcur2
dmeta, dval = [], []
for (w, h) in shapes:
    for kind in (0, 1, 2):
        for ss in ((0, 1) if kind != 1 else (0,)):
            for pair in (0, 1):
                o, c = (dorg, dcur) if pair == 0 else (dorg2, dcur2)
                dmeta.append((w, h, kind, ss, pair))
                dval.append(R.dist(kind, o, 0, 64, c, 0, 80, w, h, 8, ss))
out["dist_meta"] = np.array(dmeta, np.int32)
out["dist_val"] = np.array(dval, np.uint32)

# ---- MV-bit cost -----------------------------------------------------------------------------------
mmeta, mval = [], []
for t in range(400):
    lam = float(rng.uniform(1, 400))
    R.set_lambda(lam)
    x, y, px, py = [int(v) for v in rng.integers(-700, 700, 4)]
    sc = int(rng.integers(0, 3))
    mmeta.append((lam, x, y, sc, px, py))
    mval.append(R.mv_cost(x, y, sc, px, py))
out["mv_meta"] = np.array(mmeta, np.float64)
out["mv_val"] = np.array(mval, np.uint32)

# ---- a small frame through xPatternSearchFracDIF + NN_pred -------------------------------------------
W, H = 128, 96
org, refs, motions = fme.pu_list.synth_frames(W, H, n_refs=2, seed=77)
recs = fme.pu_list.make_records(W, H, motions, seed=5, amp=True)
# a few hostile records: integer MVs near the clip range (TComDataCU.cpp:2773-2786) and large predictors
recs["mvIntX"][::37] = -71 - recs["x"][::37]
recs["mvIntY"][::41] = (H + 7) - recs["y"][::41]
recs["mvPredX"][::29] = 511
recs["mvPredY"][::31] = -640
frame = ob.CpuFrame(org, refs, margin=80)
lam = fme.pu_list.slice_lambda(22)
for i in range(len(recs)):
    p = recs[i]
    roff = frame.ref_offs[p["refSlot"]] + (int(p["y"]) + int(p["mvIntY"])) * frame.rstride + int(p["x"]) + int(p["mvIntX"])
    recs["err"][i] = R.int_surface(frame.org, int(p["y"]) * W + int(p["x"]), W, int(p["w"]), int(p["h"]),
                                   frame.refs[p["refSlot"]], roff, frame.rstride)
out["small_org"], out["small_refs"] = org, np.stack(refs)
out["small_recs"] = recs.view(np.uint8).reshape(len(recs), -1)
out["small_lambda"] = np.array([lam])
R.init(22, 1, 1); R.set_lambda(lam)
out["small_res_had"] = frame.reference_run(recs, 1).view(np.uint8).reshape(len(recs), -1)
R.init(22, 0, 1); R.set_lambda(lam)
out["small_res_sad"] = frame.reference_run(recs, 1).view(np.uint8).reshape(len(recs), -1)
for qp in (22, 27, 32, 37):
    R.init(qp, 1, 1); R.set_lambda(lam)
    out["small_res_nn%d" % qp] = frame.reference_run(recs, 2).view(np.uint8).reshape(len(recs), -1)

# ---- NN_pred on hand-made grids (incl. SURVEY.md appendix B) -------------------------------------------
grids = [([1200, 900, 1300, 800, 500, 850, 1250, 950, 1400], 8, 8),
         ([60000, 41000, 52000, 30000, 9000, 33000, 58000, 39000, 61000], 16, 16),
         ([250000, 180000, 240000, 150000, 60000, 90000, 230000, 120000, 200000], 32, 32),
         ([5000, 5200, 5100, 4800, 4700, 4900, 5300, 5250, 5400], 4, 8),
         ([800000, 500000, 700000, 300000, 100000, 350000, 750000, 450000, 820000], 64, 64),
         ([30000, 20000, 25000, 15000, 7000, 12000, 28000, 18000, 26000], 12, 16),
         ([30000, 20000, 25000, 15000, 7000, 12000, 28000, 18000, 26000], 16, 12)]
for t in range(300):
    base = float(rng.uniform(50, 4e5))
    e = (base * rng.uniform(0.2, 3.0, 9)).astype(np.uint32)
    if t % 9 == 0:
        e = rng.integers(0, 2 ** 31, 9).astype(np.uint32)
    grids.append((list(map(int, e)), int(rng.choice([4, 8, 12, 16, 24, 32, 48, 64])),
                  int(rng.choice([4, 8, 12, 16, 24, 32, 48, 64]))))
out["nn_grids"] = np.array([g[0] + [g[1], g[2]] for g in grids], np.int64)  # 9 errors, H, W
for qp in (22, 27, 32, 37):
    R.init(qp, 1, 1)
    res = []
    for e, hh, ww in grids:
        cls, hxy, qxy = R.nn_pred(e, hh, ww)
        res.append((cls, hxy[0], hxy[1], qxy[0], qxy[1]))
    out["nn_out%d" % qp] = np.array(res, np.int32)

# ---- bi-prediction average: TComYuv::addAvg on 14-bit intermediates (own RNG stream: added later) ----
rng2 = np.random.default_rng(20261019)
R.init(22, 1, 1)
avg_a = rng2.integers(-14312, 14249, (64, 64)).astype(np.int16)
avg_b = rng2.integers(-14312, 14249, (64, 64)).astype(np.int16)
avg_a[0, :8] = [-14312, 14248, -8192, 8128, 0, -1, 1, 14248]
avg_b[0, :8] = [-14312, 14248, -8192, 8128, 0, -1, -64, -14312]
out["avg_a"], out["avg_b"] = avg_a, avg_b
avg_shapes = [(64, 64), (32, 16), (16, 32), (8, 8), (4, 8), (8, 4), (12, 16), (2, 4), (6, 8)]  # incl. chroma-sized
out["avg_shapes"] = np.array(avg_shapes, np.int32)
out["avg_out"] = np.concatenate([R.add_avg(avg_a, 0, 64, avg_b, 0, 64, w, h).ravel() for (w, h) in avg_shapes]).astype(np.int16)

path = os.path.join(HERE, "fme_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(fcases), "filter cases,", len(dval), "dist cases,",
      len(recs), "PUs,", len(grids), "NN grids")
