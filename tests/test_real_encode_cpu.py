"""CPU: the oracle reproduces every fractional-ME decision the REFERENCE ENCODER ITSELF made on a real encode
(416x240, encoder_lowdelay_P_main.cfg, QP22 -- BASELINE.json configs[0]); fixture captured by
oracle/capture/make_capture.py at the call site TEncSearch.cpp:4534-4541."""
import numpy as np
import pytest

import oracle_bindings as ob
import real_encode
from common import fme


@pytest.mark.parametrize("capture,calls", [(real_encode.CAPTURES[0], 31017), (real_encode.CAPTURES[1], 10311),
                                           (real_encode.CAPTURES[2], 51491)])
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
        bad = np.nonzero((std[u] != p["want_std"][u]).any(1))[0]
        assert len(bad) == 0, (p["poc"], len(bad), p["pus"][u][bad[:3]], std[u][bad[:3]], p["want_std"][u][bad[:3]])
        assert np.array_equal(nn[k & u], p["want_nn"][k & u]), p["poc"]
        total += int(u.sum())
    assert total > 5000 and (calls is None or total == calls)
