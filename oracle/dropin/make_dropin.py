"""TEST INFRASTRUCTURE ONLY.  Executes the drop-in boundary: the reference encoder with its two fractional-ME call
sites bound to libfme_b200.so through the C++ adaptor, next to the stock encoder.

Builds under oracle/_ref/dropin/ (git-ignored; reference sources are never copied into the repository, the patched
translation unit is generated there):

  TAppEncoderStock   the reference encoder, every object compiled unmodified from /root/reference
  TAppEncoderFme     the same objects with ONE translation unit replaced: TEncSearch.cpp whose
                     TEncSearch::xMotionEstimation (TEncSearch.cpp:4531-4541) calls
                     FmeHmAdaptor::xPatternSearchFracDIF / FmeHmAdaptor::NN_pred (immediate mode, INTEGRATION.md
                     section 1) instead of the member function / NN_pred(); NN weights come from
                     fme_load_nn_csv_dir (FmeHmAdaptor::init).  Linked against hm16.9-nn_fme_b200/libfme_b200.so
                     with a relative rpath, so the pair travels to the GPU box with the snapshot.
  adaptor_check      a C++ program over the reference's own objects + the adaptor: batched enqueue/flush on a PU
                     list read from a file, slotOf, distFunc and filterHor/filterVer against the reference's
                     TComRdCost / TComInterpolationFilter in the same process (tests/test_dropin.py drives it).
  cfg/               the two encoder configuration files the test encodes with (copied: /root/reference does not
                     exist on the GPU box)

tests/test_dropin.py (-m gpu) encodes a synthetic 416x240 clip with both encoders and asserts that the bitstreams and
the reconstructions are byte-identical.

The binding differs from INTEGRATION.md section 1 in one respect only: the adaptor is a file-static object of the
patched TEncSearch.cpp, bound lazily per picture inside xMotionEstimation, so that TEncSearch.h (and with it every
other translation unit of the encoder) stays untouched.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref", "dropin")
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
PKG = os.path.join(ROOT, "hm16.9-nn_fme_b200")

BINDING = r'''
// ---- fme_b200 binding (injected by oracle/dropin/make_dropin.py; INTEGRATION.md section 1) ----
#include "fme_hm_adaptor.h"
namespace fmebind {
static FmeHmAdaptor g;
static bool   inited = false;
static int    lastPoc = -(1 << 30);
static double lastLambda = -1.0;
static long   nFrac = 0, nNN = 0;
struct Report { ~Report() { if (inited) fprintf(stderr, "fme_b200 binding: %ld xPatternSearchFracDIF + %ld NN_pred calls served by the engine\n", nFrac, nNN); } };
static Report report;
static void bind(TEncCfg* cfg, TComRdCost* rd, TComDataCU* cu)
{
  if (!inited)
  {
    const char* wd = getenv("FME_WEIGHTS_DIR");   // .../DL/blowing of a reference checkout
    g.init(cfg->getSourceWidth(), cfg->getSourceHeight(), 32, 64, cfg->getUseHADME(),
           cfg->getFastInterSearchMode() != FASTINTERSEARCH_DISABLED, cfg->getQP(), wd ? wd : "DL/blowing");
    inited = true;
  }
  TComSlice* s = cu->getSlice();
  if (s->getPOC() != lastPoc)
  { // once per coded picture: source picture and the reference pictures of both lists
    lastPoc = s->getPOC();
    g.setOrgPicture(cu->getPic()->getPicYuvOrg());
    Int slot = 0;
    for (Int l = 0; l < 2; l++)
      for (Int r = 0; r < s->getNumRefIdx(RefPicList(l)); r++)
        g.setRefPicture(slot++, s->getRefPic(RefPicList(l), r)->getPicYuvRec());
    lastLambda = -1.0;
  }
  if (rd->getLambda() != lastLambda) { lastLambda = rd->getLambda(); g.setSliceLambda(lastLambda); }
}
}
'''

FRAC_CALL = "  xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, cMvHalf, cMvQter, ruiCost );\n"
FRAC_BOUND = r'''  if ( bBi )
  { // the bi-predictive pattern is not a picture view: batched enqueueBi is the engine's interface for it
    xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, cMvHalf, cMvQter, ruiCost );
  }
  else
  {
    fmebind::bind( m_pcEncCfg, m_pcRdCost, pcCU );
    fmebind::g.xPatternSearchFracDIF( bIsLosslessCoded, pcPatternKey, piRefY, iRefStride, &rcMv, *pcMvPred, cMvHalf, cMvQter, ruiCost );
    fmebind::nFrac++;
  }
'''
NN_CALL = "  //Run our ANN model\n  NN_pred();\n"
NN_BOUND = r'''  //Run our ANN model
  {
    fmebind::bind( m_pcEncCfg, m_pcRdCost, pcCU );
    Int nnOut = 0;
    // NN_pred reads array_e[0..7] whatever its size (TEncSearch.cpp:88): passing the vector's storage hands the engine
    // the same (possibly stale) values the reference's own code would read
    fmebind::g.NN_pred( &array_e[0], C, PUHeight, PUWidth, MVX_HALF, MVX_QRTER, MVY_HALF, MVY_QRTER, nnOut );
    NN_out = nnOut;
    array_e.clear();
    fmebind::nNN++;
  }
'''


def sh(cmd):
    subprocess.check_call(cmd, shell=True)


def newer(target, *deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def patch_source():
    src = open(os.path.join(REF, "source/Lib/TLibEncoder/TEncSearch.cpp"), "rb").read().replace(b"\r", b"").decode("latin-1")
    anchor = '#include <iostream>\n'
    assert anchor in src
    src = src.replace(anchor, anchor + BINDING, 1)
    assert src.count(FRAC_CALL) == 1 and src.count(NN_CALL) == 1
    src = src.replace(FRAC_CALL, FRAC_BOUND, 1).replace(NN_CALL, NN_BOUND, 1)
    path = os.path.join(OUT, "TEncSearch_fme.cpp")
    open(path, "w", encoding="latin-1").write(src)
    return path


def build(verbose=False):
    """Idempotent; returns the output directory, or None when /root/reference is absent (GPU box: prebuilt files)."""
    if not os.path.isdir(REF):
        return None
    os.makedirs(os.path.join(OUT, "cfg"), exist_ok=True)
    sh("make -C %s -j8 ref > /dev/null" % os.path.join(ROOT, "oracle"))
    adaptor = os.path.join(PKG, "adaptor", "fme_hm_adaptor.h")
    header = os.path.join(ROOT, "include", "fme_b200.h")
    me = os.path.abspath(__file__)
    flags = ("-std=gnu++11 -O2 -w -fPIC -I%s -I%s/source/Lib -I%s/source/Lib/TLibEncoder -I%s -I%s"
             % (os.path.join(ROOT, "oracle", "eigen_standin"), REF, REF, os.path.join(ROOT, "include"),
                os.path.join(PKG, "adaptor")))
    link = "-L%s -lfme_b200 -Wl,-rpath,'$ORIGIN/../../../hm16.9-nn_fme_b200'" % PKG
    app = []
    for f in ("TAppEncCfg", "TAppEncTop", "encmain"):
        o = "%s/app.%s.o" % (OUT, f)
        if not newer(o, me):
            sh("g++ %s -c %s/source/App/TAppEncoder/%s.cpp -o %s" % (flags, REF, f, o))
        app.append(o)
    pol = "%s/pol.o" % OUT
    if not newer(pol, me):
        sh("g++ %s -c %s/source/Lib/TAppCommon/program_options_lite.cpp -o %s" % (flags, REF, pol))
    lib_objs = [os.path.join(OBJ, o) for o in sorted(os.listdir(OBJ)) if o.endswith(".o") and o != "ref_harness.o"]
    stock = os.path.join(OUT, "TAppEncoderStock")
    if not newer(stock, me, *lib_objs):
        sh("g++ -o %s %s %s %s" % (stock, " ".join(app), pol, " ".join(lib_objs)))
    fme_obj = os.path.join(OUT, "TEncSearch_fme.o")
    if not newer(fme_obj, me, adaptor, header):
        sh("g++ %s -c %s -o %s" % (flags, patch_source(), fme_obj))
    bound = os.path.join(OUT, "TAppEncoderFme")
    others = [o for o in lib_objs if os.path.basename(o) != "TLibEncoder.TEncSearch.o"]
    if not newer(bound, fme_obj, *others):
        sh("g++ -o %s %s %s %s %s %s" % (bound, " ".join(app), pol, fme_obj, " ".join(others), link))
    chk_src = os.path.join(HERE, "adaptor_check.cpp")
    chk = os.path.join(OUT, "adaptor_check")
    if not newer(chk, chk_src, adaptor, header, me):
        sh("g++ %s %s -o %s %s %s" % (flags, chk_src, chk, " ".join(lib_objs), link))
    for c in ("encoder_lowdelay_P_main.cfg", "per-sequence/BlowingBubbles.cfg"):
        shutil.copyfile(os.path.join(REF, "cfg", c), os.path.join(OUT, "cfg", os.path.basename(c)))
    if verbose:
        print("built", stock, bound, chk)
    return OUT


if __name__ == "__main__":
    if build(verbose=True) is None:
        sys.exit("no /root/reference here: nothing built")
