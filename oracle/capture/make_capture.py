"""TEST INFRASTRUCTURE ONLY.  Capture the reference encoder's own fractional-ME calls on a real encode.

Builds the full reference encoder (TAppEncoder) from /root/reference with ONE file replaced by a patched copy:
TEncSearch.cpp with a recorder injected around the xPatternSearchFracDIF / NN_pred call site of
TEncSearch::xMotionEstimation (TEncSearch.cpp:4534-4541).  The patched copy is generated under oracle/_ref/capture/
(git-ignored; reference sources are never copied into the repository) and compiled with the same flags as the
other reference objects.  The encoder is then run on a synthetic 416x240 4:2:0 sequence with
cfg/encoder_lowdelay_P_main.cfg at QP22 (BASELINE.json configs[0]), and the capture is stored as
tests/golden/real_encode_416x240.npz:

  per coded picture : source luma, the reference (reconstructed) luma pictures it searched, slice lambda
  per FME call      : PU x,y,w,h, reference index, integer MV, predictor, array_e[0..7] + C (as the reference had
                      them), bBi / lossless flags, and the reference's OUTPUTS: cMvHalf, cMvQter, ruiCost right after
                      xPatternSearchFracDIF, MVX/MVY_HALF/QRTER and NN_out right after NN_pred

Run in the dev container:  python oracle/capture/make_capture.py   (about 1-2 minutes)
"""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref", "capture")
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
sys.path.insert(0, ROOT)

RECORDER = r'''
// ---- fme capture (injected by oracle/capture/make_capture.py) ----
#include <cstdio>
#include <cstdlib>
#include <set>
namespace fmecap {
static FILE* f = NULL;
static std::set<long long> dumped;
static FILE* file() {
  if (!f) { const char* p = getenv("FME_CAPTURE_FILE"); f = fopen(p ? p : "fme_capture.bin", "wb"); }
  return f;
}
static void dumpPlane(int tag, int poc, const Pel* p, int stride, int w, int h) {
  int hdr[5] = {tag, poc, w, h, 0};
  fwrite(hdr, sizeof(int), 5, file());
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { unsigned char v = (unsigned char)p[y * stride + x]; fwrite(&v, 1, 1, file()); }
}
}
'''

PRE_CALL = r'''
  // fme capture: inputs as the reference has them in hand
  int fmecap_x = 0, fmecap_y = 0;
  std::vector<uint> fmecap_e;
  {
    TComPic* rp = pcCU->getSlice()->getRefPic( eRefPicList, iRefIdxPred );
    const Pel* ro = rp->getPicYuvRec()->getAddr(COMPONENT_Y);
    long off = long(piRefY - ro);
    fmecap_y = int(off / iRefStride); fmecap_x = int(off - (long)fmecap_y * iRefStride);
    int cur = pcCU->getSlice()->getPOC();
    long long key = ((long long)cur << 32) | (unsigned)(rp->getPOC() & 0xffff) | 0x10000;
    if (!fmecap::dumped.count(key)) { fmecap::dumped.insert(key);
      fmecap::dumpPlane(2, (cur << 16) | (rp->getPOC() & 0xffff), ro, iRefStride, rp->getPicYuvRec()->getWidth(COMPONENT_Y), rp->getPicYuvRec()->getHeight(COMPONENT_Y)); }
    long long okey = ((long long)cur << 32);
    if (!fmecap::dumped.count(okey)) { fmecap::dumped.insert(okey);
      TComPicYuv* og = pcCU->getPic()->getPicYuvOrg();
      fmecap::dumpPlane(1, cur, og->getAddr(COMPONENT_Y), og->getStride(COMPONENT_Y), og->getWidth(COMPONENT_Y), og->getHeight(COMPONENT_Y)); }
  }
'''

POST_FRAC = r'''
  unsigned fmecap_cost = ruiCost;
  fmecap_e = array_e;
'''

POST_NN = r'''
  {
    int hdr[5] = {3, pcCU->getSlice()->getPOC(), 0, 0, 0};
    fwrite(hdr, sizeof(int), 5, fmecap::file());
    int rec[32]; memset(rec, 0, sizeof(rec));
    rec[0] = fmecap_x; rec[1] = fmecap_y; rec[2] = iRoiWidth; rec[3] = iRoiHeight;
    rec[4] = pcCU->getSlice()->getRefPic( eRefPicList, iRefIdxPred )->getPOC(); rec[5] = (int)eRefPicList;
    rec[6] = (rcMv.getHor()); rec[7] = (rcMv.getVer());           // integer MV (before the << 2 below)
    rec[8] = pcMvPred->getHor(); rec[9] = pcMvPred->getVer();
    rec[10] = (int)fmecap_e.size();
    for (int i = 0; i < 8 && i < (int)fmecap_e.size(); i++) rec[11 + i] = (int)fmecap_e[i];
    rec[19] = (int)C;
    rec[20] = bBi ? 1 : 0; rec[21] = bIsLosslessCoded ? 1 : 0;
    rec[22] = cMvHalf.getHor(); rec[23] = cMvHalf.getVer(); rec[24] = cMvQter.getHor(); rec[25] = cMvQter.getVer();
    rec[26] = (int)fmecap_cost;
    rec[27] = MVX_HALF; rec[28] = MVY_HALF; rec[29] = MVX_QRTER; rec[30] = MVY_QRTER; rec[31] = (int)NN_out;
    fwrite(rec, sizeof(int), 32, fmecap::file());
    double lam = m_pcRdCost->getLambda();
    fwrite(&lam, sizeof(double), 1, fmecap::file());
    if (bBi) {
      // the other list's uni-prediction that removeHighFreq subtracted (TEncSearch.cpp:4462-4472): its picture and
      // the MV xPredInterUni used (clipped), plus a checksum of the search pattern 2*org - pred
      RefPicList eOther = RefPicList(1 - (Int)eRefPicList);
      TComMv om = pcCU->getCUMvField(eOther)->getMv(uiPartAddr);
      pcCU->clipMv(om);
      TComPic* op = pcCU->getSlice()->getRefPic(eOther, pcCU->getCUMvField(eOther)->getRefIdx(uiPartAddr));
      int cur = pcCU->getSlice()->getPOC();
      long long key = ((long long)cur << 32) | (unsigned)(op->getPOC() & 0xffff) | 0x10000;
      if (!fmecap::dumped.count(key)) { fmecap::dumped.insert(key);
        fmecap::dumpPlane(2, (cur << 16) | (op->getPOC() & 0xffff), op->getPicYuvRec()->getAddr(COMPONENT_Y),
                          op->getPicYuvRec()->getStride(COMPONENT_Y), op->getPicYuvRec()->getWidth(COMPONENT_Y),
                          op->getPicYuvRec()->getHeight(COMPONENT_Y)); }
      long long chk = 0;
      const Pel* pat = pcPatternKey->getROIY();
      for (int y = 0; y < iRoiHeight; y++)
        for (int x = 0; x < iRoiWidth; x++) chk += (long long)pat[y * pcPatternKey->getPatternLStride() + x] * (y * iRoiWidth + x + 1);
      int hdr4[5] = {4, cur, 0, 0, 0};
      fwrite(hdr4, sizeof(int), 5, fmecap::file());
      int r4[6] = {op->getPOC(), om.getHor(), om.getVer(), (int)eOther, (int)(chk & 0x7fffffff), (int)((chk >> 31) & 0x7fffffff)};
      fwrite(r4, sizeof(int), 6, fmecap::file());
    }
  }
'''


def sh(cmd, **kw):
    subprocess.check_call(cmd, shell=True, **kw)


def patch_source():
    src = open(os.path.join(REF, "source/Lib/TLibEncoder/TEncSearch.cpp"), "rb").read().replace(b"\r", b"").decode("latin-1")
    anchor_inc = '#include <iostream>\n'
    assert anchor_inc in src
    src = src.replace(anchor_inc, anchor_inc + RECORDER, 1)
    call = "  xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, cMvHalf, cMvQter, ruiCost );\n"
    assert src.count(call) == 1
    src = src.replace(call, PRE_CALL + call + POST_FRAC, 1)
    nn = "  //Run our ANN model\n  NN_pred();\n"
    assert src.count(nn) == 1
    src = src.replace(nn, nn + POST_NN, 1)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "TEncSearch_capture.cpp")
    open(path, "w", encoding="latin-1").write(src)
    return path


def build_encoder():
    sh("make -C %s -j8 ref > /dev/null" % os.path.join(ROOT, "oracle"))
    patched = patch_source()
    flags = "-std=gnu++11 -O2 -w -fPIC -I%s -I%s/source/Lib -I%s/source/Lib/TLibEncoder" % (
        os.path.join(ROOT, "oracle", "eigen_standin"), REF, REF)
    sh("g++ %s -c %s -o %s/TEncSearch_capture.o" % (flags, patched, OUT))
    app_objs = []
    for f in ("TAppEncCfg", "TAppEncTop", "encmain"):
        sh("g++ %s -c %s/source/App/TAppEncoder/%s.cpp -o %s/app.%s.o" % (flags, REF, f, OUT, f))
        app_objs.append("%s/app.%s.o" % (OUT, f))
    sh("g++ %s -c %s/source/Lib/TAppCommon/program_options_lite.cpp -o %s/pol.o" % (flags, REF, OUT))
    objs = [os.path.join(OBJ, o) for o in sorted(os.listdir(OBJ))
            if o.endswith(".o") and o not in ("TLibEncoder.TEncSearch.o", "ref_harness.o")]
    exe = os.path.join(OUT, "TAppEncoderCapture")
    sh("g++ -o %s %s %s/TEncSearch_capture.o %s/pol.o %s" % (exe, " ".join(app_objs), OUT, OUT, " ".join(objs)))
    return exe


def synth_yuv(path, w, h, frames, seed=7, scale=1.0):
    """Moving low-pass texture + noise, 4:2:0 planar 8-bit."""
    import fme_loader
    fme = fme_loader.load()
    rng = np.random.default_rng(seed)
    base = fme.pu_list._lowpass_noise(h + 64, w + 64, rng)
    base = np.clip(base + rng.normal(0, 5.0, base.shape), 0, 255)
    with open(path, "wb") as f:
        for t in range(frames):
            dx, dy = 1.3 * t * scale, 0.7 * t * scale
            ix, iy, fx, fy = int(dx), int(dy), dx - int(dx), dy - int(dy)
            a = base[16 + iy:16 + iy + h + 1, 16 + ix:16 + ix + w + 1]
            y = (a[:-1, :-1] * (1 - fx) + a[:-1, 1:] * fx) * (1 - fy) + (a[1:, :-1] * (1 - fx) + a[1:, 1:] * fx) * fy
            y = np.clip(np.rint(y + rng.normal(0, 1.5, y.shape)), 0, 255).astype(np.uint8)
            f.write(y.tobytes())
            c = np.full((h // 2, w // 2), 128, np.uint8)
            f.write(c.tobytes()); f.write(c.tobytes())


def parse_capture(path):
    data = open(path, "rb").read()
    pos = 0
    orgs, refs, recs, lams, others = {}, {}, [], [], []
    while pos < len(data):
        tag, a, w, h, _ = struct.unpack_from("<5i", data, pos)
        pos += 20
        if tag in (1, 2):
            plane = np.frombuffer(data, np.uint8, w * h, pos).reshape(h, w).copy()
            pos += w * h
            if tag == 1:
                orgs[a] = plane
            else:
                refs[(a >> 16, a & 0xffff)] = plane
        elif tag == 3:
            rec = struct.unpack_from("<32i", data, pos)
            pos += 128
            lam = struct.unpack_from("<d", data, pos)[0]
            pos += 8
            recs.append((a,) + rec)
            lams.append(lam)
            others.append((-1, 0, 0, 0, 0, 0))
        elif tag == 4:   # belongs to the tag-3 record just read
            others[-1] = struct.unpack_from("<6i", data, pos)
            pos += 24
        else:
            raise ValueError("bad tag %d at %d" % (tag, pos))
    return orgs, refs, np.array(recs, np.int64), np.array(lams), np.array(others, np.int64)


def main():
    # usage: make_capture.py [qp [frames [seed [motion_scale]]]]   (defaults reproduce real_encode_416x240.npz)
    a = sys.argv[1:]
    w, h = 416, 240
    qp = int(a[0]) if len(a) > 0 else 22
    frames = int(a[1]) if len(a) > 1 else 3
    seed = int(a[2]) if len(a) > 2 else 7
    scale = float(a[3]) if len(a) > 3 else 1.0
    cfg = a[4] if len(a) > 4 else "encoder_lowdelay_P_main.cfg"   # e.g. encoder_randomaccess_main.cfg (B slices)
    exe = build_encoder()
    yuv = os.path.join(OUT, "syn_416x240_%d.yuv" % seed)
    synth_yuv(yuv, w, h, frames, seed=seed, scale=scale)
    cap = os.path.join(OUT, "capture.bin")
    env = dict(os.environ, FME_CAPTURE_FILE=cap)
    cmd = [exe, "-c", REF + "/cfg/" + cfg, "-c", REF + "/cfg/per-sequence/BlowingBubbles.cfg",
           "-i", yuv, "-f", str(frames), "-q", str(qp), "-b", os.path.join(OUT, "str.bin"), "-o", os.path.join(OUT, "rec.yuv")]
    log = subprocess.run(cmd, env=env, capture_output=True, text=True)
    open(os.path.join(OUT, "encode.log"), "w").write(log.stdout + log.stderr)
    assert log.returncode == 0, log.stdout[-2000:] + log.stderr[-2000:]
    orgs, refs, recs, lams, others = parse_capture(cap)
    print("captured", len(recs), "FME calls;", len(orgs), "source pictures;", len(refs), "(cur,ref) reference pictures")
    out = {"recs": recs.astype(np.int32), "lambda": lams, "qp": np.array([qp], np.int32)}
    if (others[:, 0] >= 0).any():
        # per call: POC / quarter-pel MV / list of the other list's prediction, pattern checksum (bi calls; -1 else)
        out["other"] = others.astype(np.int32)
    for poc, p in orgs.items():
        out["org_%d" % poc] = p
    for (cur, rp), p in refs.items():
        out["ref_%d_%d" % (cur, rp)] = p
    tag = "" if "lowdelay_P" in cfg else "_" + cfg.split("_")[1]
    name = "real_encode_416x240.npz" if (qp, seed, tag) == (22, 7, "") else "real_encode_416x240%s_qp%d.npz" % (tag, qp)
    path = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
