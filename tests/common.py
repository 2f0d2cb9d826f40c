"""Shared helpers for the test suite."""
import os

import numpy as np

import fme_loader
import oracle_bindings as ob

HERE = os.path.dirname(os.path.abspath(__file__))
fme = fme_loader.load()
PU_DTYPE, RESULT_DTYPE = fme.PU_DTYPE, fme.RESULT_DTYPE

_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = dict(np.load(os.path.join(HERE, "golden", "fme_golden.npz")))
    return _golden


def golden_recs(g):
    return np.ascontiguousarray(g["small_recs"]).view(PU_DTYPE).reshape(-1)


def golden_res(g, key):
    return np.ascontiguousarray(g[key]).view(RESULT_DTYPE).reshape(-1)


def std_fields(r):
    return np.stack([r["halfX"], r["halfY"], r["qterX"], r["qterY"]], 1).astype(np.int64), r["cost"].astype(np.int64)


def nn_fields(r):
    return np.stack([r["nnHalfX"], r["nnHalfY"], r["nnQterX"], r["nnQterY"], r["nnClass"].astype(np.int8)], 1).astype(np.int64)


def filter_case_iter(g):
    meta, flat = g["filt_meta"], g["filt_out"]
    pos = 0
    for (is_ver, luma, frac, w, h, first, last, bd) in meta:
        n = int(w) * int(h)
        yield (int(is_ver), int(luma), int(frac), int(w), int(h), int(first), int(last), int(bd),
               flat[pos:pos + n].reshape(int(h), int(w)))
        pos += n


def filter_src(g, is_ver, first, bd):
    if is_ver and not first:
        return g["filt_inter"]
    return g["filt_src8"] if bd == 8 else g["filt_src10"]


FILT_OFF, FILT_STRIDE = 8 * 40 + 8, 40
