"""Pack the reference's per-QP NN_pred CSV weights (DL/blowing/<qp>/*.csv) into FMNN blobs.

Run in the dev container (needs /root/reference):  python tools/pack_weights.py
Writes hm16.9-nn_fme_b200/weights/qp{22,27,32,37}.fmnn (8.3 KB each; model weights are data the
engine needs at run time -- the reference hard-codes the same numbers in TEncSearch.cpp:470-1073).
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("nn_weights", os.path.join(ROOT, "hm16.9-nn_fme_b200", "nn_weights.py"))
nw = importlib.util.module_from_spec(spec)
spec.loader.exec_module(nw)

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/DL/blowing"
for qp in nw.QPS:
    blob = nw.blob_from_csv_dir(os.path.join(ref, str(qp)))
    out = os.path.join(nw.WEIGHTS_DIR, "qp%d.fmnn" % qp)
    with open(out, "wb") as f:
        f.write(blob)
    print(out, len(blob), nw.parse_header(blob), nw.flops_per_pu(blob))
