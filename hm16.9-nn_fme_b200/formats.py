"""On-disk formats either side of the path.

* raw planar YUV 4:2:0, 8 bit, as read/written by the reference's TLibVideoIO/TVideoIOYuv.cpp (Y plane, then Cb,
  then Cr, no header; `-i` / `-o` of TAppEncoder);
* the reference's trained models `DL/models/QP<qp>_blowing_200_train_acc*.h5` (torch state_dicts written by
  fastai 0.7: keys embs.{0,1}.weight, lins.{0,1}.{weight,bias}, bns.{0,1}.{weight,bias}, outp.{weight,bias},
  bn.weight) -> FMNN blob.  Only BN weight/bias are used, as NN_pred does (TEncSearch.cpp:122,127).
"""
import numpy as np

from . import nn_weights


def yuv420_frame_bytes(width, height):
    return width * height * 3 // 2


def read_yuv420_frame(path, width, height, index):
    """Returns (Y, Cb, Cr) uint8 arrays of frame `index`."""
    n = yuv420_frame_bytes(width, height)
    with open(path, "rb") as f:
        f.seek(n * index)
        buf = f.read(n)
    if len(buf) != n:
        raise EOFError("frame %d is beyond the end of %s" % (index, path))
    a = np.frombuffer(buf, np.uint8)
    y = a[:width * height].reshape(height, width)
    c = (width // 2) * (height // 2)
    cb = a[width * height:width * height + c].reshape(height // 2, width // 2)
    cr = a[width * height + c:].reshape(height // 2, width // 2)
    return y.copy(), cb.copy(), cr.copy()


def read_yuv420_raw(path, width, height, index):
    """Frame `index` as the flat byte buffer of the file (Y, Cb, Cr back to back): what fme_upload_ref_yuv420_u8 /
    fme_upload_org_yuv420_u8 take -- no per-plane copies on the host."""
    n = yuv420_frame_bytes(width, height)
    a = np.fromfile(path, np.uint8, count=n, offset=n * index)
    if a.size != n:
        raise EOFError("frame %d is beyond the end of %s" % (index, path))
    return a


def write_yuv420_frame(f, y, cb=None, cr=None):
    """Append one frame to an open binary file; missing chroma is written as mid-grey."""
    h, w = y.shape
    f.write(np.ascontiguousarray(y, np.uint8).tobytes())
    grey = np.full((h // 2, w // 2), 128, np.uint8)
    f.write(np.ascontiguousarray(grey if cb is None else cb, np.uint8).tobytes())
    f.write(np.ascontiguousarray(grey if cr is None else cr, np.uint8).tobytes())


def blob_from_state_dict(path, mapper_csv):
    """FMNN blob from a reference `.h5` torch state_dict plus the mean/stdev mapper CSV (14.mapper_<qp>.csv)."""
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=False)
    g = lambda k: sd[k].detach().cpu().numpy().astype(np.float64)
    rows = nn_weights._read_csv(mapper_csv)
    hidden = [(g("lins.0.weight"), g("lins.0.bias"), g("bns.0.weight"), g("bns.0.bias")),
              (g("lins.1.weight"), g("lins.1.bias"), g("bns.1.weight"), g("bns.1.bias"))]
    return nn_weights.pack_blob(rows[0], rows[1], g("bn.weight"), [g("embs.0.weight"), g("embs.1.weight")], hidden,
                                g("outp.weight"), g("outp.bias"))
