"""Write an FMNN weight blob back out in the reference's `DL/blowing/<qp>/{1..14}.*.csv` directory layout
(SURVEY.md A.4), so that fme_load_nn_csv_dir / FmeHmAdaptor::init can be exercised where /root/reference does not
exist (the GPU box).  float32 values are printed with repr precision: reading them back gives the same floats.

  python tools/fmnn_to_csv.py 22 tests/golden/blowing/22
"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fme_loader  # noqa: E402


def main():
    qp, out = int(sys.argv[1]), sys.argv[2]
    fme = fme_loader.load()
    blob = fme.nn_weights.load_blob(qp)
    h = fme.nn_weights.parse_header(blob)
    assert h["nEmb"] == 2 and h["nHidden"] == 2
    p = np.frombuffer(blob, "<f4", offset=64)
    pos = [0]

    def take(*shape):
        n = int(np.prod(shape))
        a = p[pos[0]:pos[0] + n].reshape(shape)
        pos[0] += n
        return a
    n_in = 9 + 2 * h["embDim"]
    mean, stdev, gamma_in = take(1, 9), take(1, 9), take(1, 9)
    emb0, emb1 = take(h["embRows"], h["embDim"]), take(h["embRows"], h["embDim"])
    h1, h2 = h["hidden"]
    w0, b0, g0, be0 = take(h1, n_in), take(1, h1), take(1, h1), take(1, h1)
    w1, b1, g1, be1 = take(h2, h1), take(1, h2), take(1, h2), take(1, h2)
    wo, bo = take(h["nOut"], h2), take(1, h["nOut"])
    assert pos[0] == len(p)
    files = {"1.emb0-weight": emb0, "2.emb1-weight": emb1, "3.lins0-weight": w0, "4.lins1-weight": w1,
             "5.outp-weight": wo, "6.lins0-bias": b0, "7.lins1-bias": b1, "8.outp-bias": bo, "9.bn-weight": gamma_in,
             "10.bns0-weight": g0, "11.bns1-weight": g1, "12.bns0-bias": be0, "13.bns1-bias": be1,
             "14.mapper_%d" % qp: np.concatenate([mean, stdev])}
    os.makedirs(out, exist_ok=True)
    for name, a in files.items():
        with open(os.path.join(out, name + ".csv"), "w") as f:
            for row in a:
                f.write("\t\t\t" + ",".join("%.9g" % float(v) for v in row) + ",\n")
    print("wrote", len(files), "files to", out)


if __name__ == "__main__":
    main()
