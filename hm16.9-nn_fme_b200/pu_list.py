"""PU batch records and the synthetic workloads of SURVEY.md section 8d.

The record layouts mirror `fme_pu` / `fme_result` / `fme_mc_pu` in include/fme_b200.h.  The PU list is
what the reference's RD recursion visits for every fully-inside CU at depths 0-3 (TEncCu.cpp:451-614):
2Nx2N, 2NxN x2, Nx2N x2 (5 PUs per CU; counts 10 295 / 214 500 / 860 100 per frame and reference for
416x240 / 1080p / 2160p), optionally plus the AMP shapes.
"""
import numpy as np

PU_DTYPE = np.dtype(
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot", "u1"), ("flags", "u1"),
     ("mvIntX", "<i2"), ("mvIntY", "<i2"), ("mvPredX", "<i2"), ("mvPredY", "<i2"), ("err", "<u4", (9,))],
    align=True)
HEAD_DTYPE = np.dtype(   # fme_pu_head: the record without the error grid (computed on the device)
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot", "u1"), ("flags", "u1"),
     ("mvIntX", "<i2"), ("mvIntY", "<i2"), ("mvPredX", "<i2"), ("mvPredY", "<i2")])
assert HEAD_DTYPE.itemsize == 16


def heads_of(recs):
    """fme_pu records -> fme_pu_head records (drops err[])."""
    h = np.zeros(len(recs), HEAD_DTYPE)
    for f in HEAD_DTYPE.names:
        h[f] = recs[f]
    return h


GRID_DTYPE = np.dtype([("pu", "<i4"), ("err", "<u4", (9,))])   # fme_err_grid: a host error grid for heads[pu]
assert GRID_DTYPE.itemsize == 40


def grids_of(recs, min_area):
    """fme_err_grid entries for the records of at least `min_area` luma samples (the caller's array_e / C): the split
    a caller makes between bus bytes and K0 time (fme_submit_heads_grids*)."""
    idx = np.nonzero(recs["w"].astype(np.int32) * recs["h"].astype(np.int32) >= min_area)[0]
    g = np.zeros(len(idx), GRID_DTYPE)
    g["pu"] = idx
    g["err"] = recs["err"][idx]
    return g


COMPACT_DTYPE = np.dtype(   # fme_pu_compact: head + nine 24-bit little-endian grid values + one reserved byte
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot", "u1"), ("flags", "u1"),
     ("mvIntX", "<i2"), ("mvIntY", "<i2"), ("mvPredX", "<i2"), ("mvPredY", "<i2"), ("err24", "u1", (27,)), ("reserved", "u1")])
assert COMPACT_DTYPE.itemsize == 44


def compact_of(recs):
    """fme_pu records -> (fme_pu_compact records, fme_err_grid entries for the PUs with a grid value of 2^24 or more):
    what fme_pu_compact_pack does per record (include/fme_b200.h)."""
    c = np.zeros(len(recs), COMPACT_DTYPE)
    for f in HEAD_DTYPE.names:
        c[f] = recs[f]
    e = recs["err"].astype(np.uint32)
    b = np.empty((len(recs), 9, 3), np.uint8)
    b[:, :, 0] = e & 0xff
    b[:, :, 1] = (e >> 8) & 0xff
    b[:, :, 2] = (e >> 16) & 0xff
    c["err24"] = b.reshape(len(recs), 27)
    idx = np.nonzero((e >> 24).any(axis=1))[0]
    g = np.zeros(len(idx), GRID_DTYPE)
    g["pu"] = idx
    g["err"] = recs["err"][idx]
    return c, g


RESULT_DTYPE = np.dtype(
    [("halfX", "i1"), ("halfY", "i1"), ("qterX", "i1"), ("qterY", "i1"), ("cost", "<u4"),
     ("nnHalfX", "i1"), ("nnHalfY", "i1"), ("nnQterX", "i1"), ("nnQterY", "i1"), ("nnClass", "u1"),
     ("pad", "u1", (3,))],
    align=True)
MC_PU_DTYPE = np.dtype(
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot", "u1"), ("flags", "u1"),
     ("mvX", "<i2"), ("mvY", "<i2")], align=True)
MC_BI_PU_DTYPE = np.dtype(
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot0", "u1"), ("refSlot1", "u1"),
     ("mv0X", "<i2"), ("mv0Y", "<i2"), ("mv1X", "<i2"), ("mv1Y", "<i2")])
assert MC_BI_PU_DTYPE.itemsize == 16
CAND_DTYPE = np.dtype(   # fme_cand_pu: one AMVP-template / merge candidate
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot", "u1"), ("flags", "u1"),
     ("mvX", "<i2"), ("mvY", "<i2"), ("bits", "<u2"), ("groupStart", "<u2")])
assert CAND_DTYPE.itemsize == 16
assert PU_DTYPE.itemsize == 52 and RESULT_DTYPE.itemsize == 16 and MC_PU_DTYPE.itemsize == 12

# lowdelay_P GOP entry QP offsets / factors (cfg/encoder_lowdelay_P_main.cfg:24-27)
LOWDELAY_P_QP_OFFSETS = (3, 2, 3, 1)
LOWDELAY_P_QP_FACTORS = (0.4624, 0.4624, 0.4624, 0.578)


def slice_lambda(qp, qp_offset=3, qp_factor=0.4624, depth=0, had_me=True):
    """TEncSlice.cpp:290-325: lambda = QPfactor * 2^((QP + offset - 12)/3) [* clip(2,4,(qp-12)/6) if depth>0]
    [* 0.95 if HadamardME is off]."""
    qp_temp = float(qp + qp_offset) - 12.0
    lam = qp_factor * 2.0 ** (qp_temp / 3.0)
    if depth > 0:
        lam *= min(max(qp_temp / 6.0, 2.0), 4.0)
    if not had_me:
        lam *= 0.95
    return lam


def enumerate_pus(width, height, ctu=64, min_cu=8, amp=False):
    """(x, y, w, h) int arrays for one frame and one reference."""
    xs, ys, ws, hs = [], [], [], []

    def add(x, y, w, h):
        xs.append(x.ravel()); ys.append(y.ravel())
        ws.append(np.full(x.size, w)); hs.append(np.full(x.size, h))

    cu = ctu
    while cu >= min_cu:
        nx, ny = width // cu, height // cu
        gy, gx = np.meshgrid(np.arange(ny) * cu, np.arange(nx) * cu, indexing="ij")
        add(gx, gy, cu, cu)                                   # 2Nx2N
        add(gx, gy, cu, cu // 2); add(gx, gy + cu // 2, cu, cu // 2)   # 2NxN
        add(gx, gy, cu // 2, cu); add(gx + cu // 2, gy, cu // 2, cu)   # Nx2N
        if amp and cu >= 16:
            q = cu // 4
            add(gx, gy, cu, q); add(gx, gy + q, cu, cu - q)            # 2NxnU
            add(gx, gy, cu, cu - q); add(gx, gy + cu - q, cu, q)       # 2NxnD
            add(gx, gy, q, cu); add(gx + q, gy, cu - q, cu)            # nLx2N
            add(gx, gy, cu - q, cu); add(gx + cu - q, gy, q, cu)       # nRx2N
        cu //= 2
    return (np.concatenate(xs).astype(np.int16), np.concatenate(ys).astype(np.int16),
            np.concatenate(ws).astype(np.uint8), np.concatenate(hs).astype(np.uint8))


def _lowpass_noise(h, w, rng):
    """Low-pass-filtered uniform noise in 0..255 (three separable box blurs via running sums)."""
    a = rng.uniform(0.0, 1.0, (h, w)).astype(np.float32)
    for r in (9, 5, 3):
        for axis in (1, 0):
            p = np.concatenate([np.take(a, range(-r, 0), axis=axis), a, np.take(a, range(0, r + 1), axis=axis)], axis=axis)
            cs = np.cumsum(p, axis=axis, dtype=np.float64)
            n = a.shape[axis]
            hi = np.take(cs, range(2 * r + 1, 2 * r + 1 + n), axis=axis)
            lo = np.take(cs, range(0, n), axis=axis)
            a = ((hi - lo) / (2 * r + 1)).astype(np.float32)
    a = (a - a.min()) / max(float(a.max() - a.min()), 1e-6)
    return a * 255.0


def synth_frames(width, height, n_refs=4, seed=1000):
    """Source frame + n_refs reference frames (uint8) and the true motion (dx, dy) in pixels of each
    reference relative to the source: ref_k(x, y) ~ org(x - dx_k, y - dy_k), (dx, dy) = (1.25, 0.75)*(k+1)."""
    rng = np.random.default_rng(seed)
    pad = 16
    base = _lowpass_noise(height + 2 * pad, width + 2 * pad, rng)
    # add texture so that SATD/SSE surfaces are not degenerate
    base = np.clip(base + rng.normal(0.0, 6.0, base.shape), 0, 255).astype(np.float32)
    org = np.clip(np.rint(base[pad:pad + height, pad:pad + width]), 0, 255).astype(np.uint8)
    refs, motions = [], []
    for k in range(n_refs):
        dx, dy = 1.25 * (k + 1), 0.75 * (k + 1)
        ix, iy = int(np.floor(dx)), int(np.floor(dy))
        fx, fy = dx - ix, dy - iy
        # ref(x, y) = base(x - dx, y - dy), bilinear
        y0, x0 = pad - iy - 1, pad - ix - 1
        a = base[y0:y0 + height + 1, x0:x0 + width + 1]
        top = a[:-1, :-1] * fx + a[:-1, 1:] * (1 - fx)
        bot = a[1:, :-1] * fx + a[1:, 1:] * (1 - fx)
        r = top * fy + bot * (1 - fy)
        r = r + rng.normal(0.0, 2.0, r.shape)
        refs.append(np.clip(np.rint(r), 0, 255).astype(np.uint8))
        motions.append((dx, dy))
    return org, refs, motions


def make_records(width, height, motions, seed=0, amp=False, err_on_gpu=False):
    """The per-frame PU batch: every PU of enumerate_pus() against every reference slot.
    intMV = round(true motion) + U{-1,0,1}; mvPred = 4*intMV + U{-6..6}."""
    rng = np.random.default_rng(seed)
    x, y, w, h = enumerate_pus(width, height, amp=amp)
    n1 = len(x)
    recs = np.zeros(n1 * len(motions), PU_DTYPE)
    for s, (dx, dy) in enumerate(motions):
        r = recs[s * n1:(s + 1) * n1]
        r["x"], r["y"], r["w"], r["h"] = x, y, w, h
        r["refSlot"] = s
        r["flags"] = 0x02 if err_on_gpu else 0
        mx = int(np.rint(dx)) + rng.integers(-1, 2, n1)
        my = int(np.rint(dy)) + rng.integers(-1, 2, n1)
        r["mvIntX"], r["mvIntY"] = mx, my
        r["mvPredX"] = mx * 4 + rng.integers(-6, 7, n1)
        r["mvPredY"] = my * 4 + rng.integers(-6, 7, n1)
    return recs


def band_of_pus(recs, band, n_bands, height, ctu=64):
    """CTU-row band sharding: band b owns CTU rows [b*rows//n, (b+1)*rows//n) -- balanced to within
    one CTU row, never empty while there are at least n_bands CTU rows."""
    lo, hi = band_rows(band, n_bands, height, ctu)
    ctu_row = recs["y"] // ctu
    return recs[(ctu_row >= lo) & (ctu_row < hi)]


def band_rows(band, n_bands, height, ctu=64):
    rows = (height + ctu - 1) // ctu
    return band * rows // n_bands, (band + 1) * rows // n_bands


PER_PU_WORK = 22.0   # work that scales with the PU count (K3, binning, bookkeeping) in PU-pixel equivalents: 1080p bench,
                     # K3 0.146 ms per 858 000 PUs against K2 0.77 ms per 97.5 M PU pixels


def band_mask_balanced(recs, band, n_bands, width, ctu=64):
    """Boolean mask of band `band` when the PU list is cut into n_bands contiguous runs of CTUs (raster order) carrying
    equal work (PU pixels + PER_PU_WORK per PU): boundaries fall wherever the cumulative sum says, i.e. also mid-row
    (whole-row bands leave 34 CTU rows / 8 GPUs = 5-row and 4-row bands, a 25 % imbalance at 2160p).  Every PU of a
    CTU stays with its CTU; the union of the bands is the list and the bands are disjoint."""
    ctus_x = (width + ctu - 1) // ctu
    cid = (recs["y"].astype(np.int64) // ctu) * ctus_x + recs["x"].astype(np.int64) // ctu
    work = np.bincount(cid, weights=recs["w"].astype(np.float64) * recs["h"] + PER_PU_WORK)
    cum = np.cumsum(work)
    total = cum[-1] if len(cum) else 0.0
    # CTU c belongs to band floor(n * (work before c) / total)
    before = cum - work
    owner = np.minimum((before * n_bands / max(total, 1.0)).astype(np.int64), n_bands - 1)
    return owner[cid] == band


def band_of_pus_balanced(recs, band, n_bands, width, ctu=64):
    return recs[band_mask_balanced(recs, band, n_bands, width, ctu)]


def source_rows(recs):
    """Picture-row range [begin, end) of the SOURCE picture the records read (rows y .. y + h - 1 of every PU): what a rank
    of the banded mode has to upload (fme_upload_org_device_u8_rows)."""
    if len(recs) == 0:
        return 0, 0
    y = recs["y"].astype(np.int64)
    return int(y.min()), int((y + recs["h"]).max())


def referenced_rows(recs):
    """Picture-row range [begin, end) of the sub-pel planes the records can read (banded mode: the rows a rank has to
    interpolate).  A PU at row y with integer MV my and height h reads plane rows y + my - 1 .. y + my + h - 1 in K2 (half-
    pel regions start one row up, quarter-pel candidates shift by -1 / 0) and y + my - 1 .. y + my + h in K0; two rows of
    slack on either side."""
    if len(recs) == 0:
        return 0, 0
    top = recs["y"].astype(np.int64) + recs["mvIntY"]
    return int(top.min()) - 3, int((top + recs["h"]).max()) + 3
