"""The drop-in boundary, EXECUTED: the reference encoder with its two fractional-ME call sites
(TEncSearch.cpp:4534 xPatternSearchFracDIF, :4541 NN_pred) bound to libfme_b200.so through
hm16.9-nn_fme_b200/adaptor/fme_hm_adaptor.h must produce the stock encoder's bitstream byte for byte, and the adaptor's
batched / immediate / block-level entry points must agree with the reference's own objects.

The binaries are built by oracle/dropin/make_dropin.py (in `__graft_entry__.build()`, where /root/reference exists)
into oracle/_ref/dropin/ and travel to the GPU box with the snapshot.
"""
import hashlib
import os
import struct
import subprocess

import numpy as np
import pytest

from common import fme, HERE
import real_encode

ROOT = os.path.dirname(HERE)
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
WEIGHTS = os.path.join(HERE, "golden", "blowing")   # <qp>/1..14.*.csv written by tools/fmnn_to_csv.py


def _need(*names):
    missing = [n for n in names if not os.path.exists(os.path.join(DROPIN, n))]
    assert not missing, ("oracle/_ref/dropin/%s missing: run `python oracle/dropin/make_dropin.py` in the dev container "
                         "(it is part of __graft_entry__.build())" % missing)


def _synth_yuv(path, w, h, frames, seed=11):
    """Moving low-pass texture + noise, 4:2:0 planar 8-bit (the generator of oracle/capture/make_capture.py)."""
    rng = np.random.default_rng(seed)
    base = fme.pu_list._lowpass_noise(h + 64, w + 64, rng)
    base = np.clip(base + rng.normal(0, 5.0, base.shape), 0, 255)
    with open(path, "wb") as f:
        for t in range(frames):
            dx, dy = 1.3 * t, 0.7 * t
            ix, iy, fx, fy = int(dx), int(dy), dx - int(dx), dy - int(dy)
            a = base[16 + iy:16 + iy + h + 1, 16 + ix:16 + ix + w + 1]
            y = (a[:-1, :-1] * (1 - fx) + a[:-1, 1:] * fx) * (1 - fy) + (a[1:, :-1] * (1 - fx) + a[1:, 1:] * fx) * fy
            y = np.clip(np.rint(y + rng.normal(0, 1.5, y.shape)), 0, 255).astype(np.uint8)
            f.write(y.tobytes())
            c = np.full((h // 2, w // 2), 128, np.uint8)
            f.write(c.tobytes())
            f.write(c.tobytes())


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.gpu
def test_encoder_with_engine_produces_the_stock_bitstream(tmp_path):
    """416x240 lowdelay_P QP22 (BASELINE.json configs[0]), one I and two P pictures: every xPatternSearchFracDIF of a
    uni-predictive PU and every NN_pred of the encode are served by the engine (immediate mode, weights through
    fme_load_nn_csv_dir); bitstream and reconstruction must equal the stock encoder's."""
    _need("TAppEncoderStock", "TAppEncoderFme", "cfg/encoder_lowdelay_P_main.cfg", "cfg/BlowingBubbles.cfg")
    yuv = str(tmp_path / "syn_416x240.yuv")
    _synth_yuv(yuv, 416, 240, 3)
    out = {}
    for tag in ("Stock", "Fme"):
        bits, rec = str(tmp_path / (tag + ".bin")), str(tmp_path / (tag + "_rec.yuv"))
        cmd = [os.path.join(DROPIN, "TAppEncoder" + tag), "-c", os.path.join(DROPIN, "cfg", "encoder_lowdelay_P_main.cfg"),
               "-c", os.path.join(DROPIN, "cfg", "BlowingBubbles.cfg"), "-i", yuv, "-f", "3", "-q", "22", "-b", bits, "-o", rec]
        r = subprocess.run(cmd, env=dict(os.environ, FME_WEIGHTS_DIR=WEIGHTS), capture_output=True, text=True, timeout=900,
                           cwd=str(tmp_path))
        assert r.returncode == 0, (tag, r.stdout[-1500:], r.stderr[-1500:])
        out[tag] = (bits, rec, r.stderr)
    served = [l for l in out["Fme"][2].splitlines() if "fme_b200 binding" in l]
    assert served, out["Fme"][2][-500:]
    n_frac, n_nn = int(served[0].split(":")[1].split()[0]), int(served[0].split("+")[1].split()[0])
    assert n_frac > 10000 and n_nn >= n_frac, served   # the engine really was on the path
    assert os.path.getsize(out["Stock"][0]) > 1000
    assert _md5(out["Stock"][0]) == _md5(out["Fme"][0]), "bitstreams differ"
    assert _md5(out["Stock"][1]) == _md5(out["Fme"][1]), "reconstructions differ"


@pytest.mark.gpu
def test_adaptor_batched_and_immediate_against_reference_objects(tmp_path):
    """FmeHmAdaptor (C++) over the PU list of a captured reference-encoder picture: enqueue + flush must reproduce the
    outputs the reference encoder itself produced for these calls; slotOf / immediate calls / distFunc / filterHor /
    filterVer are checked inside the program against the reference's own objects."""
    _need("adaptor_check")
    pic = real_encode.load()[-1]            # last P picture of the lowdelay_P QP22 capture: two references
    sel = np.nonzero(pic["uni"])[0]
    pus = np.ascontiguousarray(pic["pus"][sel])
    H, W = pic["org"].shape
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<6id", W, H, len(pic["refs"]), len(pus), 1, pic["qp"], pic["lam"]))
        f.write(np.ascontiguousarray(pic["org"], np.uint8).tobytes())
        for r in pic["refs"]:
            f.write(np.ascontiguousarray(r, np.uint8).tobytes())
        f.write(pus.tobytes())
    r = subprocess.run([os.path.join(DROPIN, "adaptor_check"), inp, outp, WEIGHTS], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    assert "adaptor_check ok" in r.stdout
    got = np.fromfile(outp, fme.RESULT_DTYPE)
    assert len(got) == len(pus)
    want_std, want_nn = pic["want_std"][sel], pic["want_nn"][sel]
    for k, fld in enumerate(("halfX", "halfY", "qterX", "qterY", "cost")):
        np.testing.assert_array_equal(got[fld].astype(np.int64), want_std[:, k], err_msg=fld)
    ok = pic["nn_ok"][sel]                  # calls whose array_e held 8 fresh values (the others read stale globals)
    for k, fld in enumerate(("nnHalfX", "nnHalfY", "nnQterX", "nnQterY", "nnClass")):
        np.testing.assert_array_equal(got[fld].astype(np.int64)[ok], want_nn[ok, k], err_msg=fld)


def test_csv_weight_directory_equals_the_shipped_blob():
    """tests/golden/blowing/22 (what fme_load_nn_csv_dir reads in the tests above) packs to the shipped QP22 blob."""
    assert fme.nn_weights.blob_from_csv_dir(os.path.join(WEIGHTS, "22")) == fme.nn_weights.load_blob(22)
