"""CPU: host-side logic (PU lists, weight containers, lambda, sharding incl. a 2-rank gloo run)."""
import os
import subprocess
import sys

import numpy as np

from common import fme

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pu_list_counts_match_survey():
    for (w, h), n in (((416, 240), 10295), ((1920, 1080), 214500), ((3840, 2160), 860100)):
        x, y, pw, ph = fme.pu_list.enumerate_pus(w, h)
        assert len(x) == n
        assert int((x.astype(int) + pw <= w).all()) and int((y.astype(int) + ph <= h).all())
    # measured PU histogram of the reference encode: 8x4 == 4x8 == 2 * 8x8 etc.
    x, y, pw, ph = fme.pu_list.enumerate_pus(416, 240)
    cnt = lambda a, b: int(((pw == a) & (ph == b)).sum())
    assert cnt(8, 4) == cnt(4, 8) == 2 * cnt(8, 8) and cnt(64, 64) == 18 and cnt(8, 8) == 1560


def test_amp_shapes_are_hevc_pu_sizes():
    x, y, pw, ph = fme.pu_list.enumerate_pus(128, 128, amp=True)
    shapes = set(zip(pw.tolist(), ph.tolist()))
    for s in ((64, 16), (64, 48), (16, 64), (48, 64), (32, 8), (32, 24), (8, 32), (24, 32), (16, 4), (16, 12), (4, 16), (12, 16)):
        assert s in shapes


def test_slice_lambda_matches_reference_formula():
    assert abs(fme.pu_list.slice_lambda(22) - 9.3213998955) < 1e-9
    import oracle_bindings as ob
    L = ob.oracle().L
    for qp in (22, 27, 32, 37):
        for off, fac in zip(fme.pu_list.LOWDELAY_P_QP_OFFSETS, fme.pu_list.LOWDELAY_P_QP_FACTORS):
            for had in (True, False):
                assert fme.pu_list.slice_lambda(qp, off, fac, 0, had) == L.orc_slice_lambda(qp + off, fac, 0, int(had))


def test_weight_blobs():
    nw = fme.nn_weights
    for qp in nw.QPS:
        blob = nw.load_blob(qp)
        h = nw.parse_header(blob)
        assert h["hidden"] == [22, 20] and h["nOut"] == 49 and h["nEmb"] == 2
        assert nw.flops_per_pu(blob) == 3588
        assert len(blob) == 64 + 4 * (27 + 64 + 22 * 17 + 66 + 20 * 22 + 60 + 49 * 20 + 49)
    assert nw.select_qp(30) == 22 and nw.select_qp(27) == 27  # TEncSearch.cpp:925 fallback
    b3 = nw.synthetic_blob((40, 40, 40), n_emb=0, seed=1)
    assert nw.parse_header(b3)["hidden"] == [40, 40, 40] and nw.flops_per_pu(b3) == 11040
    if os.path.isdir("/root/reference/DL/blowing/22"):
        assert nw.blob_from_csv_dir("/root/reference/DL/blowing/22") == nw.load_blob(22)


def test_band_sharding_partitions_the_list():
    org, refs, motions = fme.pu_list.synth_frames(256, 192, n_refs=1, seed=3)
    recs = fme.pu_list.make_records(256, 192, motions)
    for nb in (1, 2, 3, 4, 8):
        parts = [fme.pu_list.band_of_pus(recs, b, nb, 192) for b in range(nb)]
        assert sum(len(p) for p in parts) == len(recs)
        for b, p in enumerate(parts):
            lo, hi = fme.pu_list.band_rows(b, nb, 192)
            if len(p):
                assert (p["y"] // 64 >= lo).all() and (p["y"] // 64 < hi).all()
    # 2160p over 8 bands: 34 CTU rows -> every band non-empty
    assert all(hi > lo for lo, hi in (fme.pu_list.band_rows(b, 8, 2160) for b in range(8)))


import pytest  # noqa: E402


@pytest.mark.parametrize("split", ["rows", "balanced"])
def test_two_rank_gloo_band_merge(split):
    """world_size-2 run of the banded host logic over gloo: each rank owns a band of CTUs (whole rows, or the
    pixel-balanced runs bench.py uses), computes its PUs (oracle as the stand-in compute on CPU), results gathered on
    rank 0 equal the single-process run."""
    script = os.path.join(ROOT, "tests", "gloo_band_worker.py")
    port = "29541" if split == "rows" else "29543"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", port, script, split],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "BAND_MERGE_OK" in out.stdout


def test_yuv420_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    frames = [(rng.integers(0, 256, (48, 64)).astype(np.uint8), rng.integers(0, 256, (24, 32)).astype(np.uint8),
               rng.integers(0, 256, (24, 32)).astype(np.uint8)) for _ in range(3)]
    p = tmp_path / "t.yuv"
    with open(p, "wb") as f:
        for y, cb, cr in frames:
            fme.formats.write_yuv420_frame(f, y, cb, cr)
    assert os.path.getsize(p) == 3 * fme.formats.yuv420_frame_bytes(64, 48)
    for i, (y, cb, cr) in enumerate(frames):
        gy, gcb, gcr = fme.formats.read_yuv420_frame(str(p), 64, 48, i)
        assert np.array_equal(gy, y) and np.array_equal(gcb, cb) and np.array_equal(gcr, cr)
    import pytest
    with pytest.raises(EOFError):
        fme.formats.read_yuv420_frame(str(p), 64, 48, 3)


def test_h5_models_equal_csv_weights():
    """The reference ships the same weights three times; the .h5 state_dicts must pack to the same
    blob as the CSV directory up to float32 rounding of the stored values."""
    import glob
    import pytest
    models = sorted(glob.glob("/root/reference/DL/models/QP22_*.h5"))
    if not models:
        pytest.skip("reference models not available on this box")
    a = fme.formats.blob_from_state_dict(models[0], "/root/reference/DL/blowing/22/14.mapper_22.csv")
    b = fme.nn_weights.load_blob(22)
    assert a[:64] == b[:64]
    fa, fb = np.frombuffer(a[64:], "<f4"), np.frombuffer(b[64:], "<f4")
    assert np.allclose(fa, fb, rtol=2e-6, atol=1e-7)


def test_balanced_bands_partition_the_list_and_even_out_the_work():
    """Pixel-balanced CTU bands (bench.py banded leg): a partition of the list, CTUs never split, work within a CTU's
    worth of the mean -- against whole-row bands, which at 2160p / 8 differ by a full CTU row."""
    motions = [(1.25, 0.75), (2.5, 1.5)]
    for (W, H, nb) in ((416, 240, 3), (3840, 2160, 8), (1920, 1080, 4)):
        recs = fme.pu_list.make_records(W, H, motions[:1], seed=2)
        masks = [fme.pu_list.band_mask_balanced(recs, b, nb, W) for b in range(nb)]
        assert np.array_equal(np.sum(masks, axis=0), np.ones(len(recs), np.int64))       # disjoint cover
        px = recs["w"].astype(np.int64) * recs["h"] + fme.pu_list.PER_PU_WORK
        work = np.array([px[m].sum() for m in masks], np.float64)
        ctu_work = 12.0 * 64 * 64 * 1.5                                                  # 12 coverings of a CTU's pixels + per-PU work
        assert work.max() - work.min() <= 2 * ctu_work, (W, H, work)
        ctus_x = (W + 63) // 64
        cid = (recs["y"] // 64).astype(np.int64) * ctus_x + recs["x"] // 64
        owners = [set(np.unique(cid[m]).tolist()) for m in masks]
        assert sum(len(o) for o in owners) == len(set(cid.tolist()))                      # a CTU has one owner
        if (W, H) == (3840, 2160):
            rows = np.array([px[(recs["y"] // 64 >= lo) & (recs["y"] // 64 < hi)].sum()
                             for lo, hi in (fme.pu_list.band_rows(b, nb, H) for b in range(nb))], np.float64)
            assert work.max() / work.mean() < 1.01 < rows.max() / rows.mean()


def test_band_row_ranges_and_compact_records():
    """Host helpers of the banded mode and the 44-byte record format: a band's source rows are exactly the rows its PUs
    cover, its referenced plane rows contain every row a search can touch (integer MV, +-1 row for the half- / quarter-pel
    regions); compact_of keeps every grid below 2^24 in the record and routes the others to the full-grid list."""
    W, H = 416, 240
    _, _, motions = fme.pu_list.synth_frames(W, H, n_refs=2, seed=3)
    recs = fme.pu_list.make_records(W, H, motions, seed=4, amp=True)
    for band in range(3):
        mine = recs[fme.pu_list.band_mask_balanced(recs, band, 3, W)]
        lo, hi = fme.pu_list.source_rows(mine)
        assert lo == int(mine["y"].min()) and hi == int((mine["y"].astype(int) + mine["h"]).max()) and 0 <= lo < hi <= H
        rlo, rhi = fme.pu_list.referenced_rows(mine)
        top = mine["y"].astype(int) + mine["mvIntY"]
        assert rlo <= int(top.min()) - 1 and rhi >= int((top + mine["h"]).max()) + 1
    assert fme.pu_list.source_rows(recs[:0]) == (0, 0)
    r = recs[:5].copy()
    r["err"] = 0
    r["err"][0, 8] = (1 << 24) - 1      # fits
    r["err"][1, 0] = 1 << 24            # does not
    r["err"][3, 4] = 0xffffffff
    comp, big = fme.pu_list.compact_of(r)
    assert list(big["pu"]) == [1, 3] and np.array_equal(big["err"], r["err"][[1, 3]])
    e24 = comp["err24"].reshape(5, 9, 3).astype(np.uint32)
    val = e24[:, :, 0] | (e24[:, :, 1] << 8) | (e24[:, :, 2] << 16)
    assert np.array_equal(val, r["err"] & 0xffffff)
    for f in fme.HEAD_DTYPE.names:
        assert np.array_equal(comp[f], r[f])
