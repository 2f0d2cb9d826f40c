"""Loader for tests/golden/real_encode_416x240*.npz: the reference encoder's own FME calls captured on real
416x240 encodes (oracle/capture/make_capture.py): lowdelay_P QP22 (3 frames) and QP37 (2 frames, faster motion),
and randomaccess QP32 (4 frames: B slices, two lists; bi-predictive refinement calls are flagged `uni == False` and
become FME_PU_BI records: pattern 2*org - other list's prediction, other slot / MV in err[0] / err[1])."""
import os

import numpy as np

from common import HERE, PU_DTYPE

COLS = ["poc", "x", "y", "w", "h", "refPoc", "list", "mvIntX", "mvIntY", "predX", "predY", "esize",
        "e0", "e1", "e2", "e3", "e4", "e5", "e6", "e7", "C", "bBi", "lossless",
        "halfX", "halfY", "qterX", "qterY", "cost", "nnHx", "nnHy", "nnQx", "nnQy", "nnOut"]


CAPTURES = ["real_encode_416x240.npz", "real_encode_416x240_qp37.npz", "real_encode_416x240_randomaccess_qp32.npz"]


def load(name=CAPTURES[0]):
    g = dict(np.load(os.path.join(HERE, "golden", name)))
    qp = int(g["qp"][0]) if "qp" in g else 22
    recs = g["recs"]
    col = {n: recs[:, i] for i, n in enumerate(COLS)}
    other = g["other"] if "other" in g else None   # per call: other list's POC, MV x/y, list, pattern checksum (bi calls)
    pictures = []
    for poc in sorted(set(col["poc"].tolist())):
        sel = np.nonzero(col["poc"] == poc)[0]
        ref_pocs = set(col["refPoc"][sel].tolist())
        if other is not None:
            ref_pocs |= set(other[sel][other[sel][:, 0] >= 0][:, 0].tolist())
        ref_pocs = sorted(ref_pocs)
        slot_of = {rp: s for s, rp in enumerate(ref_pocs)}
        pus = np.zeros(len(sel), PU_DTYPE)
        pus["x"], pus["y"], pus["w"], pus["h"] = col["x"][sel], col["y"][sel], col["w"][sel], col["h"][sel]
        pus["refSlot"] = [slot_of[rp] for rp in col["refPoc"][sel]]
        pus["flags"] = col["lossless"][sel] & 1
        pus["mvIntX"], pus["mvIntY"] = col["mvIntX"][sel], col["mvIntY"][sel]
        pus["mvPredX"], pus["mvPredY"] = col["predX"][sel], col["predY"][sel]
        e = np.stack([col["e%d" % i][sel] for i in range(8)], 1).astype(np.uint32)
        # IN_errors << array_e[0..3], C, array_e[4..7]  (TEncSearch.cpp:88)
        pus["err"][:, 0:4] = e[:, 0:4]
        pus["err"][:, 4] = col["C"][sel].astype(np.uint32)
        pus["err"][:, 5:9] = e[:, 4:8]
        bi = col["bBi"][sel] != 0
        if other is not None and bi.any():
            # bi-predictive refinement calls: the pattern is 2*org - (other list's prediction); see fme_pu / orc_pu
            o = other[sel]
            pus["flags"][bi] |= 0x04
            pus["err"][bi] = 0
            pus["err"][bi, 0] = [slot_of[int(rp)] for rp in o[bi, 0]]
            pus["err"][bi, 1] = (o[bi, 1].astype(np.int64) & 0xffff) | ((o[bi, 2].astype(np.int64) & 0xffff) << 16)
        lam = g["lambda"][sel]
        assert (lam == lam[0]).all()
        pictures.append(dict(poc=poc, qp=qp, org=g["org_%d" % poc], refs=[g["ref_%d_%d" % (poc, rp)] for rp in ref_pocs],
                             pus=pus, lam=float(lam[0]),
                             uni=(col["bBi"][sel] == 0), nn_ok=(col["esize"][sel] == 8) & (col["bBi"][sel] == 0),
                             pat_chk=(other[sel][:, 4].astype(np.int64) | (other[sel][:, 5].astype(np.int64) << 31))
                             if other is not None else None,
                             want_std=np.stack([col[k][sel] for k in ("halfX", "halfY", "qterX", "qterY", "cost")], 1),
                             want_nn=np.stack([col[k][sel] for k in ("nnHx", "nnHy", "nnQx", "nnQy", "nnOut")], 1)))
    return pictures
