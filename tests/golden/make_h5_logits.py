"""Independent pin for NN_pred: a torch float32 forward from the reference's own trained state_dicts
`DL/models/QP<qp>_blowing_200_train_acc*.h5`, with the network semantics of TEncSearch.cpp:85-134 (batch-norm layers
reduced to scale and shift, input normalisation from DL/blowing/<qp>/14.mapper_<qp>.csv, the 16<->12 height-index
swap of :93-102).  Two forwards are stored: float32 (torch's matmul has its own summation order, so it differs from
the oracle's ascending-k products in the last bits) and float64, the arbiter: tests/test_nn_pins.py asserts that the
oracle's logits are within 1e-5 relative of the float64 values (the float32 torch forward itself reaches 1.4e-5 on the
same inputs) and reports the class agreement.

The reference zeroes OUT before NN_pred returns (TEncSearch.cpp:199-201), so its logits can only be pinned this way.

  python tests/golden/make_h5_logits.py      (dev container: needs /root/reference and torch)
-> tests/golden/h5_torch_logits.npz  (per QP: err[N][9] uint32, hw[N][2], logits[N][49] float32, logits64[N][49])
"""
import glob
import os

import numpy as np
import torch

REF = "/root/reference/DL"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "h5_torch_logits.npz")
N = 400
H_INDEX = {4: 1, 8: 2, 16: 3, 12: 4, 24: 5, 32: 6, 64: 7}   # TEncSearch.cpp:93-102 (sic)
W_INDEX = {4: 1, 8: 2, 12: 3, 16: 4, 24: 5, 32: 6, 64: 7}   # TEncSearch.cpp:104-113
SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64),
          (16, 4), (4, 16), (16, 12), (12, 16), (32, 8), (8, 32), (32, 24), (24, 32), (64, 16), (16, 64), (64, 48), (48, 64)]


def read_mapper(qp):
    rows = []
    for line in open(os.path.join(REF, "blowing", str(qp), "14.mapper_%d.csv" % qp)):
        line = line.strip().rstrip(";").rstrip(",")
        if line:
            rows.append([float(t) for t in line.split(",") if t.strip()])
    return np.array(rows[0]), np.array(rows[1])


def main():
    out = {}
    rng = np.random.default_rng(2024)
    for qp in (22, 27, 32, 37):
        sd = torch.load(glob.glob(os.path.join(REF, "models", "QP%d_*.h5" % qp))[0], map_location="cpu", weights_only=False)
        mean, stdev = read_mapper(qp)
        mean_t = torch.tensor(mean, dtype=torch.float64).to(torch.float32)    # double literal -> float, as the C++ does
        stdev_t = torch.tensor(stdev, dtype=torch.float64).to(torch.float32)
        # inputs: 3x3 integer error surfaces and PU shapes the reference encoder itself produced (the real-encode
        # captures of oracle/capture/make_capture.py: array_e[0..7], C, iRoiWidth, iRoiHeight of calls with 8 fresh errors)
        cap = np.load(os.path.join(os.path.dirname(OUT), "real_encode_416x240%s.npz" % ("_qp37" if qp == 37 else "")))["recs"]
        cap = cap[cap[:, 11] == 8]                      # esize == 8 (columns as tests/real_encode.py COLS)
        pick = cap[rng.choice(len(cap), N, replace=False)]
        err = np.concatenate([pick[:, 12:16], pick[:, 20:21], pick[:, 16:20]], 1).astype(np.uint32)   # raster 3x3, TES:88
        w, h = pick[:, 3], pick[:, 4]
        hi = torch.tensor([H_INDEX.get(int(v), 0) for v in h]); wi = torch.tensor([W_INDEX.get(int(v), 0) for v in w])
        F = torch.nn.functional

        def forward(dt):
            """the float32 parameters and inputs, evaluated in dtype dt (float32: the arithmetic the reference uses;
            float64: the arbiter both float32 implementations are measured against)"""
            p = {k: v.to(dt) for k, v in sd.items()}
            e = torch.from_numpy(err.astype(np.int64)).to(torch.float32).to(dt)                # uint -> float, TES:88
            x9 = ((e - mean_t.to(dt)) / stdev_t.to(dt)) * p["bn.weight"]                        # TES:89, 116
            x = torch.cat([p["embs.0.weight"][hi], p["embs.1.weight"][wi], x9], 1)              # TES:117
            x = torch.relu(F.linear(x, p["lins.0.weight"], p["lins.0.bias"])) * p["bns.0.weight"] + p["bns.0.bias"]
            x = torch.relu(F.linear(x, p["lins.1.weight"], p["lins.1.bias"])) * p["bns.1.weight"] + p["bns.1.bias"]
            return F.linear(x, p["outp.weight"], p["outp.bias"])                                # TES:130-131
        out["err_%d" % qp] = err
        out["hw_%d" % qp] = np.stack([h, w], 1).astype(np.int32)
        out["logits_%d" % qp] = forward(torch.float32).numpy().astype(np.float32)
        out["logits64_%d" % qp] = forward(torch.float64).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
