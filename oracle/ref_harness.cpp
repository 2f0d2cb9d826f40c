// TEST INFRASTRUCTURE ONLY (oracle/). C-callable harness around the REFERENCE's own objects.
//
// Linked with the unmodified TLibCommon / TLibEncoder objects compiled from /root/reference
// (see Makefile, target `ref`) into oracle/_ref/libhmref.so.  It is the ground truth used to
//   (1) pin the plain-C restatement in fme_oracle.c,
//   (2) generate the committed fixtures in tests/golden/ (tests/golden/make_golden.py),
//   (3) serve as the CPU baseline ("kind": "reference") in bench.py.
// Nothing in the product path links or loads this file.
//
// What is reference code and what is harness glue:
//   * filterHor/filterVer, every DistFunc (SAD/SSE/HADs), setDistParam, the MV-bit cost,
//     xPatternSearchFracDIF (+ xExtDIFUpSamplingH/Q, xPatternRefinement) and NN_pred run as
//     compiled from the reference sources;
//   * the set-up lines before each call follow TEncSearch::xMotionEstimation
//     (TEncSearch.cpp:4497-4500, 4529-4536) and SURVEY.md appendix D;
//   * NN_pred is the reference's code over oracle/eigen_standin (Eigen 3.3.7 is absent).

#include <eigen3/Eigen/Dense>
#include <cstdio>
#include <cstring>
#include <vector>

#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComYuv.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComInterpolationFilter.h"
#include "TLibEncoder/TEncCfg.h"
#include "TLibEncoder/TEncSearch.h"

// globals defined in TEncSearch.cpp:55-60
extern signed short MVX_HALF, MVX_QRTER, MVY_HALF, MVY_QRTER;
extern std::vector<uint> array_e;
extern uint PUHeight, PUWidth, C;
extern Eigen::MatrixXf::Index NN_out;
void NN_pred();

namespace {

struct Probe : public TEncSearch {
  using TEncSearch::xPatternSearchFracDIF;
  TComYuv& fb(int v, int h) { return m_filteredBlock[v][h]; }
  TComInterpolationFilter& interp() { return m_if; }
};

struct State {
  TEncCfg cfg;
  TComRdCost rd;
  Probe* search;
  BitDepths bd;
  bool romReady;
  int fen;
  State() : search(0), romReady(false), fen(1) {}
};
State g;

}  // namespace

extern "C" {

// (re)create the search object; TEncSearch::init picks the NN weight set from cfg QP (TEncSearch.cpp:472).
int hmref_init(int qp, int useHadME, int fen) {
  if (!g.romReady) { initROM(); g.romReady = true; }
  if (g.search) { delete g.search; g.search = 0; }
  g.cfg.setQP(qp);
  g.cfg.setChromaFormatIdc(CHROMA_420);
  g.cfg.setQuadtreeTULog2MaxSize(5);
  g.cfg.setQuadtreeTULog2MinSize(2);
  g.cfg.setUseHADME(useHadME != 0);
  g.cfg.setFastInterSearchMode(fen ? FASTINTERSEARCH_MODE1 : FASTINTERSEARCH_DISABLED);
  g.cfg.setMotionEstimationSearchMethod(MESEARCH_DIAMOND);
  g.cfg.setRestrictMESampling(false);
  g.fen = fen;
  g.bd.recon[CHANNEL_TYPE_LUMA] = g.bd.recon[CHANNEL_TYPE_CHROMA] = 8;
  g.rd.init();
  g.search = new Probe();
  g.search->init(&g.cfg, 0, 64, 4, MESEARCH_DIAMOND, 64, 64, 4, 0, &g.rd, 0, 0);
  return 0;
}

void hmref_set_lambda(double lambda) { g.rd.setLambda(lambda, g.bd); }

// TComInterpolationFilter::filterHor / filterVer (TComInterpolationFilter.cpp:341, 377)
void hmref_filter_hor(int comp, short* src, int srcStride, short* dst, int dstStride, int w, int h, int frac,
                      int isLast, int bitDepth) {
  TComInterpolationFilter f;
  f.filterHor(ComponentID(comp), src, srcStride, dst, dstStride, w, h, frac, isLast != 0, CHROMA_420, bitDepth);
}
void hmref_filter_ver(int comp, short* src, int srcStride, short* dst, int dstStride, int w, int h, int frac,
                      int isFirst, int isLast, int bitDepth) {
  TComInterpolationFilter f;
  f.filterVer(ComponentID(comp), src, srcStride, dst, dstStride, w, h, frac, isFirst != 0, isLast != 0, CHROMA_420,
              bitDepth);
}

// TComYuv::addAvg (TComYuv.cpp:354-409) on luma-only TComYuv buffers (CHROMA_400), partition 0: the bi-prediction
// average TComPrediction::xWeightedAverage applies to the two xPredInterUni(bi = true) outputs.
void hmref_add_avg(const short* src0, int stride0, const short* src1, int stride1, short* dst, int dstStride, int w,
                   int h) {
  TComYuv a, b, d;
  a.create(64, 64, CHROMA_400); b.create(64, 64, CHROMA_400); d.create(64, 64, CHROMA_400);
  Pel* pa = a.getAddr(COMPONENT_Y); Pel* pb = b.getAddr(COMPONENT_Y);
  const int sa = a.getStride(COMPONENT_Y), sb = b.getStride(COMPONENT_Y);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) { pa[y * sa + x] = src0[y * stride0 + x]; pb[y * sb + x] = src1[y * stride1 + x]; }
  d.addAvg(&a, &b, 0, w, h, g.bd);
  const Pel* pd = d.getAddr(COMPONENT_Y);
  const int sd = d.getStride(COMPONENT_Y);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) dst[y * dstStride + x] = pd[y * sd + x];
  a.destroy(); b.destroy(); d.destroy();
}

// kind 0: integer-ME metric chosen by setDistParam(pattern, ref, stride, dp) (TComRdCost.cpp:200-229)
//         i.e. SSE for widths 4/8/16/32/64 (this fork), SAD12/24/48 otherwise; subShift applies to SAD only.
// kind 1: HADs   (setDistParam(..., iStep=1, dp, bHADME=true),  TComRdCost.cpp:232-277)
// kind 2: SADs   (setDistParam(..., iStep=1, dp, bHADME=false))
unsigned hmref_dist(int kind, const short* org, int orgStride, const short* cur, int curStride, int w, int h,
                    int bitDepth, int subShift) {
  TComPattern key;
  key.initPattern(const_cast<short*>(org), w, h, orgStride, bitDepth);
  DistParam dp;
  if (kind == 0) g.rd.setDistParam(&key, cur, curStride, dp);
  else g.rd.setDistParam(&key, cur, curStride, 1, dp, kind == 1);
  dp.bitDepth = bitDepth;
  dp.iSubShift = subShift;
  dp.compIdx = COMPONENT_Y;
  return dp.DistFunc(&dp);
}

// getCostOfVectorWithPredictor (TComRdCost.h:165-174) after selectMotionLambda(true,0,false)
unsigned hmref_mv_cost(int x, int y, int scale, int predX, int predY) {
  TComMv pred(predX, predY);
  g.rd.selectMotionLambda(true, 0, false);
  g.rd.setPredictor(pred);
  g.rd.setCostScale(scale);
  unsigned c = g.rd.getCostOfVectorWithPredictor(x, y);
  g.rd.setCostScale(0);
  return c;
}

// TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:5232-5269) with the caller set-up of
// xMotionEstimation (TEncSearch.cpp:4497-4500, 4529-4536).  `ref` points at the PU's collocated
// sample in the padded reference plane (piRefY, TEncSearch.cpp:4481).
void hmref_frac_dif(short* org, int orgStride, int w, int h, short* ref, int refStride, int mvIntX, int mvIntY,
                    int predX, int predY, int lossless, short* halfXY, short* qterXY, unsigned* cost) {
  TComPattern key;
  key.initPattern(org, w, h, orgStride, 8);
  TComMv pred(predX, predY), mvInt(mvIntX, mvIntY), mvHalf, mvQter;
  g.rd.selectMotionLambda(true, 0, false);
  g.rd.setPredictor(pred);
  g.rd.setCostScale(1);
  Distortion c = 0;
  g.search->xPatternSearchFracDIF(lossless != 0, &key, ref, refStride, &mvInt, mvHalf, mvQter, c);
  g.rd.setCostScale(0);
  halfXY[0] = mvHalf.getHor(); halfXY[1] = mvHalf.getVer();
  qterXY[0] = mvQter.getHor(); qterXY[1] = mvQter.getVer();
  *cost = c;
}

// copy out m_filteredBlock[v][h] (TComPrediction.h:79) as left by the last hmref_frac_dif call
void hmref_get_filtered_block(int v, int h, short* dst, int dstStride, int w, int hgt) {
  TComYuv& y = g.search->fb(v, h);
  const Pel* p = y.getAddr(COMPONENT_Y);
  int s = y.getStride(COMPONENT_Y);
  for (int r = 0; r < hgt; ++r) memcpy(dst + r * dstStride, p + r * s, w * sizeof(short));
}

// NN_pred (TEncSearch.cpp:85-204) through its globals.  err9 is the raster 3x3 grid
// [TL,T,TR,L,C,R,BL,B,BR]; push order of the 8 neighbours follows TEncSearch.cpp:1341-1376.
void hmref_nn_pred(const unsigned* err9, int puHeight, int puWidth, int* cls, short* halfXY, short* qterXY) {
  array_e.clear();
  for (int i = 0; i < 9; ++i)
    if (i != 4) array_e.push_back(err9[i]);
  C = err9[4];
  PUHeight = puHeight;
  PUWidth = puWidth;
  NN_pred();
  *cls = int(NN_out);
  halfXY[0] = MVX_HALF; halfXY[1] = MVY_HALF;
  qterXY[0] = MVX_QRTER; qterXY[1] = MVY_QRTER;
}

// 3x3 integer error surface with the reference's distortion code: the metric of
// xTZSearchHelp (TEncSearch.cpp:1085-1090, 1156-1166): setDistParam(pattern, ref, stride, dp),
// iSubShift = 1 when FEN (FASTINTERSEARCH_MODE1/3) and rows > 8.  Raster order output.
void hmref_int_surface(short* org, int orgStride, int w, int h, short* refAtMv, int refStride, unsigned* err9) {
  TComPattern key;
  key.initPattern(org, w, h, orgStride, 8);
  int k = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      DistParam dp;
      g.rd.setDistParam(&key, refAtMv + dy * refStride + dx, refStride, dp);
      dp.bitDepth = 8;
      dp.compIdx = COMPONENT_Y;
      if (g.fen && dp.iRows > 8) dp.iSubShift = 1;
      err9[k++] = dp.DistFunc(&dp);
    }
}

// PU record shared with include/fme_b200.h (kept layout-identical; checked in tests)
struct hmref_pu {
  short x, y;
  unsigned char w, h, refSlot, flags;
  short mvIntX, mvIntY;
  short mvPredX, mvPredY;
  unsigned err[9];
};
struct hmref_result {
  signed char halfX, halfY, qterX, qterY;
  unsigned cost;
  signed char nnHalfX, nnHalfY, nnQterX, nnQterY;
  unsigned char nnClass, pad[3];
};

// The reference's per-PU loop over a PU list (what xMotionEstimation does at TEncSearch.cpp:4534,4541),
// used as the CPU baseline.  mode bit0 = standard FME, bit1 = NN_pred.
// org: source luma (Pel, stride orgStride, origin at picture (0,0));
// refs[s]: pointer to picture sample (0,0) of padded reference plane s, stride refStride.
void hmref_run_pu_list(short* org, int orgStride, short* const* refs, int refStride, const hmref_pu* pus, int n,
                       int mode, hmref_result* out) {
  for (int i = 0; i < n; ++i) {
    const hmref_pu& p = pus[i];
    hmref_result r;
    memset(&r, 0, sizeof(r));
    if (mode & 1) {
      short hxy[2], qxy[2];
      unsigned c;
      hmref_frac_dif(org + p.y * orgStride + p.x, orgStride, p.w, p.h, refs[p.refSlot] + p.y * refStride + p.x,
                     refStride, p.mvIntX, p.mvIntY, p.mvPredX, p.mvPredY, p.flags & 1 /* bIsLosslessCoded */, hxy, qxy, &c);
      r.halfX = hxy[0]; r.halfY = hxy[1]; r.qterX = qxy[0]; r.qterY = qxy[1]; r.cost = c;
    }
    if (mode & 2) {
      int cls; short hxy[2], qxy[2];
      hmref_nn_pred(p.err, p.h, p.w, &cls, hxy, qxy);
      r.nnHalfX = hxy[0]; r.nnHalfY = hxy[1]; r.nnQterX = qxy[0]; r.nnQterY = qxy[1]; r.nnClass = (unsigned char)cls;
    }
    out[i] = r;
  }
}

}  // extern "C"
