"""NN_pred weight containers.

The reference keeps the per-QP MLP weights in three equivalent places: hard-coded
Eigen comma initialisers in TEncSearch::init (TEncSearch.cpp:470-1073), the CSV dumps
`DL/blowing/<qp>/{1..14}.*.csv` (format written by DL/edit.sh:13-17: tab-indented, comma-separated,
';'-terminated) and torch state_dicts `DL/models/*.h5`.  This module reads the CSV directory layout
and packs it into the flat "FMNN" blob that `fme_set_nn_weights` (include/fme_b200.h) consumes.

Blob = 16 x int32 header + float32 payload:
  magic 'FMNN', version, nErr, nEmb, embRows, embDim, nHidden, hidden[4], nOut, outSigmoid, reserved[3]
  mean[nErr] stdev[nErr] gammaIn[nErr] emb[nEmb][embRows][embDim]
  per hidden layer: W[out][in] b[out] gamma[out] beta[out];  output: W[nOut][in] b[nOut]
"""
import os
import struct

import numpy as np

MAGIC = 0x4E4E4D46  # "FMNN"
VERSION = 1
QPS = (22, 27, 32, 37)
WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")


def select_qp(qp):
    """TEncSearch.cpp:472,625,775,925: QP 27/32/37 pick their own set, anything else the QP22 set."""
    return qp if qp in (27, 32, 37) else 22


def _read_csv(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip().rstrip(";").rstrip(",").strip()
            if line:
                rows.append([float(t) for t in line.split(",") if t.strip()])
    return rows


def pack_blob(mean, stdev, gamma_in, embs, hidden, out_w, out_b, out_sigmoid=False):
    """hidden: list of (W[out][in], b, gamma, beta); embs: list of [rows][dim] tables (0 or 2)."""
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32)  # double literal -> float, as the C++ does
    n_err = len(mean)
    emb_rows, emb_dim = (len(embs[0]), len(embs[0][0])) if embs else (0, 0)
    sizes = [len(h[1]) for h in hidden] + [0] * (4 - len(hidden))
    hdr = struct.pack("<16i", MAGIC, VERSION, n_err, len(embs), emb_rows, emb_dim, len(hidden), *sizes,
                      len(out_b), int(out_sigmoid), 0, 0, 0)
    parts = [f32(mean), f32(stdev), f32(gamma_in)]
    n_in = n_err + len(embs) * emb_dim
    for e in embs:
        parts.append(f32(e).reshape(-1))
    for (w, b, g, be) in hidden:
        w = f32(w)
        assert w.shape == (len(b), n_in), (w.shape, len(b), n_in)
        parts += [w.reshape(-1), f32(b), f32(g), f32(be)]
        n_in = len(b)
    ow = f32(out_w)
    assert ow.shape == (len(out_b), n_in)
    parts += [ow.reshape(-1), f32(out_b)]
    return hdr + np.concatenate(parts).astype("<f4").tobytes()


def blob_from_csv_dir(path):
    """Pack one `DL/blowing/<qp>` directory (files 1..14) into a blob."""
    files = {}
    for fn in os.listdir(path):
        if fn.endswith(".csv"):
            files[int(fn.split(".")[0])] = os.path.join(path, fn)
    r = {k: _read_csv(v) for k, v in files.items()}
    embs = [r[1], r[2]]
    hidden = [(r[3], r[6][0], r[10][0], r[12][0]), (r[4], r[7][0], r[11][0], r[13][0])]
    mean, stdev = r[14][0], r[14][1]
    return pack_blob(mean, stdev, r[9][0], embs, hidden, r[5], r[8][0])


def synthetic_blob(hidden_sizes=(40, 40, 40), n_emb=0, seed=0, n_out=49):
    """Seeded random weights for architectures whose trained weights are not in the reference checkout
    (the 3-layer `blowing40` branch, README.md:106-108).  Input statistics mimic the shipped mappers."""
    rng = np.random.default_rng(seed)
    mean = rng.uniform(1.5e4, 6e4, 9)
    stdev = rng.uniform(1.2e5, 2.1e5, 9)
    gamma_in = rng.uniform(0.1, 0.9, 9)
    embs = [rng.normal(0, 0.3, (8, 4)) for _ in range(n_emb)]
    n_in = 9 + 4 * n_emb
    hidden = []
    for h in hidden_sizes:
        hidden.append((rng.normal(0, 1.0 / np.sqrt(n_in), (h, n_in)), rng.normal(0, 0.3, h),
                       rng.uniform(0.2, 1.2, h), rng.normal(0, 0.3, h)))
        n_in = h
    return pack_blob(mean, stdev, gamma_in, embs, hidden, rng.normal(0, 1.0 / np.sqrt(n_in), (n_out, n_in)),
                     rng.normal(0, 0.3, n_out))


def backup3_arrays(npz):
    """The reference's own 3-layer network (Backups/4...cpp, extracted by tests/golden/make_backup3_weights.py):
    (mean, stdev, gamma_in, hidden=[(W, b, gamma, beta)] * 3, out_w, out_b) as float64 arrays."""
    g = npz
    hidden = [(g["in_h1"], g["b1"], g["BN_gamma_1"], g["BN_beta_1"]), (g["h1_h2"], g["b2"], g["BN_gamma_2"], g["BN_beta_2"]),
              (g["h2_h3"], g["b3"], g["BN_gamma_3"], g["BN_beta_3"])]
    return g["mean"], g["stdev"], g["BN_gamma_in"], hidden, g["h3_out"], g["bout"]


def blob_from_backup3(npz):
    """FMNN blob (float32, 9-40-40-40-49, sigmoid output, no embeddings) of the reference's 3-layer backup network."""
    mean, stdev, gin, hidden, ow, ob = backup3_arrays(npz)
    return pack_blob(mean, stdev, gin, [], hidden, ow, ob, out_sigmoid=True)


def payload_f64(npz):
    """The same payload sequence as doubles, for the oracle's double-precision restatement (orc_nn_pred_f64)."""
    mean, stdev, gin, hidden, ow, ob = backup3_arrays(npz)
    parts = [mean, stdev, gin]
    for (w, b, g, be) in hidden:
        parts += [np.asarray(w).reshape(-1), b, g, be]
    parts += [np.asarray(ow).reshape(-1), ob]
    return np.concatenate([np.asarray(p, np.float64) for p in parts])


def load_blob(qp):
    """Shipped blob for a QP (generated from the reference's DL/blowing/<qp> by tools/pack_weights.py)."""
    with open(os.path.join(WEIGHTS_DIR, "qp%d.fmnn" % select_qp(qp)), "rb") as f:
        return f.read()


def parse_header(blob):
    h = struct.unpack("<16i", blob[:64])
    assert h[0] == MAGIC, "not an FMNN blob"
    return dict(nErr=h[2], nEmb=h[3], embRows=h[4], embDim=h[5], nHidden=h[6], hidden=list(h[7:7 + h[6]]),
                nOut=h[11], outSigmoid=h[12])


def flops_per_pu(blob):
    """MACs*2 per PU (2-layer = 3588 flop)."""
    h = parse_header(blob)
    n_in = h["nErr"] + h["nEmb"] * h["embDim"]
    macs = 0
    for s in h["hidden"]:
        macs += s * n_in
        n_in = s
    macs += h["nOut"] * n_in
    return 2 * macs
