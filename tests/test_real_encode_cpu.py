"""CPU: the oracle reproduces every fractional-ME decision the REFERENCE ENCODER ITSELF made on a real encode
(416x240, encoder_lowdelay_P_main.cfg, QP22 -- BASELINE.json configs[0]); fixture captured by
oracle/capture/make_capture.py at the call site TEncSearch.cpp:4534-4541."""
import numpy as np
import pytest

import oracle_bindings as ob
import real_encode
from common import fme


@pytest.mark.parametrize("capture,calls", [(real_encode.CAPTURES[0], 31017), (real_encode.CAPTURES[1], 10311),
                                           (real_encode.CAPTURES[2], 71782)])
def test_oracle_reproduces_the_reference_encoders_fme_decisions(capture, calls):
    pics = real_encode.load(capture)
    total = 0
    for p in pics:
        blob = fme.nn_weights.load_blob(p["qp"])
        frame = ob.CpuFrame(p["org"], p["refs"])
        got = frame.oracle_run(p["pus"], 3, p["lam"], 1, blob)
        std = np.stack([got["halfX"], got["halfY"], got["qterX"], got["qterY"], got["cost"]], 1).astype(np.int64)
        nn = np.stack([got["nnHalfX"], got["nnHalfY"], got["nnQterX"], got["nnQterY"], got["nnClass"]], 1).astype(np.int64)
        u, k = p["uni"], p["nn_ok"]
        assert u.all() or "randomaccess" in capture  # lowdelay_P: no bi-prediction refinement calls
        # uni-prediction calls and bi-predictive refinement calls (pattern 2*org - other list's prediction) alike
        bad = np.nonzero((std != p["want_std"]).any(1))[0]
        assert len(bad) == 0, (p["poc"], len(bad), p["pus"][bad[:3]], std[bad[:3]], p["want_std"][bad[:3]])
        assert np.array_equal(nn[k], p["want_nn"][k]), p["poc"]
        total += len(u)
    assert total > 5000 and (calls is None or total == calls)


def test_bi_pattern_checksum_matches_the_reference_encoder(orc):
    """The search pattern of bi-predictive refinement calls, rebuilt from the captured other-list MV, has the checksum
    the reference encoder computed over its own TComYuv::removeHighFreq output (TEncSearch.cpp:4462-4472)."""
    pics = real_encode.load(real_encode.CAPTURES[2])
    M, checked = 80, 0
    for p in pics:
        org = p["org"].astype(np.int16)
        refs = [ob.pad_plane(r, M) for r in p["refs"]]
        S = refs[0].shape[1]
        bi = np.nonzero(~p["uni"])[0][::37]
        for i in bi:
            r = p["pus"][i]
            x, y, w, h = int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])
            slot = int(r["err"][0]) & 0xff
            s16 = lambda v: ((v & 0xffff) ^ 0x8000) - 0x8000
            mvx, mvy = s16(int(r["err"][1])), s16(int(r["err"][1]) >> 16)
            pat = orc.bi_pattern(org, y * org.shape[1] + x, org.shape[1], refs[slot], (M + y) * S + M + x, S, w, h, mvx, mvy)
            chk = int((pat.astype(np.int64).ravel() * np.arange(1, w * h + 1)).sum()) & ((1 << 62) - 1)
            assert chk == int(p["pat_chk"][i]), (p["poc"], i, r)
            checked += 1
    assert checked > 400
