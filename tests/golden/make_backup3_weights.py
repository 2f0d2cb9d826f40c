"""Extract the reference's own 3-layer network (9 -> 40 -> 40 -> 40 -> 49, sigmoid output, no embeddings) from
`source/Lib/TLibEncoder/Backups/4. TEncSearch - SCR 3 layers - no normalization.cpp` -- the only 3-layer weight set in
the reference checkout -- into tests/golden/backup3_9_40_40_40_49.npz.

Weight arrays :65-288 (`in_h1`, `h1_h2`, `h2_h3`, `h3_out`, `b1..b3`, `bout`, `BN_gamma_in`, `BN_gamma_1..3`,
`BN_beta_1..3`), input normalisation constants :4427-4435 (`IN[i] = (E - mean) / stdev`).  The file computes in double;
the literals are kept as float64 exactly as a C++ compiler reads them.  Only numbers are extracted, no code.

  python tests/golden/make_backup3_weights.py        (dev container: needs /root/reference)
"""
import os
import re

import numpy as np

SRC = "/root/reference/source/Lib/TLibEncoder/Backups/4. TEncSearch - SCR 3 layers - no normalization.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "backup3_9_40_40_40_49.npz")


def main():
    text = open(SRC, "rb").read().replace(b"\r", b"").decode("latin-1")
    out = {}
    for m in re.finditer(r"double\s+(\w+)\s*((?:\[\d+\])+)\s*=\s*\{(.*?)\};", text, re.S):
        name, dims, body = m.group(1), [int(d) for d in re.findall(r"\[(\d+)\]", m.group(2))], m.group(3)
        nums = re.findall(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", body)
        if len(nums) != int(np.prod(dims)):
            continue   # zero-initialised scratch arrays (IN, X1, ...)
        out[name] = np.array([float(t) for t in nums], np.float64).reshape(dims)
    norm = re.findall(r"IN\[(\d)\]\s*=\s*\(\s*\w+\s*-\s*([\d.]+)\s*\)\s*/\s*([\d.]+)\s*;", text)
    assert [int(i) for i, _, _ in norm] == list(range(9)), norm
    out["mean"] = np.array([float(a) for _, a, _ in norm])
    out["stdev"] = np.array([float(b) for _, _, b in norm])
    want = {"in_h1": (40, 9), "h1_h2": (40, 40), "h2_h3": (40, 40), "h3_out": (49, 40), "b1": (40,), "b2": (40,),
            "b3": (40,), "bout": (49,), "BN_gamma_in": (9,), "BN_gamma_1": (40,), "BN_gamma_2": (40,),
            "BN_gamma_3": (40,), "BN_beta_1": (40,), "BN_beta_2": (40,), "BN_beta_3": (40,), "mean": (9,), "stdev": (9,)}
    for k, shp in want.items():
        assert out[k].shape == shp, (k, out[k].shape)
    np.savez_compressed(OUT, **{k: out[k] for k in want})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
