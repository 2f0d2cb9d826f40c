// fme_hm_adaptor.h -- header-only C++ adaptor: the reference's own signatures on top of the C ABI.
//
// Include it from a translation unit of HM-16.9-NN_FME (it needs the reference's TLibCommon headers for Pel,
// TComMv, TComPattern, TComPicYuv, DistParam) and link libfme_b200.so.  INTEGRATION.md shows the hooks in
// TEncSearch.cpp.  Two ways to use it:
//
//   immediate   FmeHmAdaptor::xPatternSearchFracDIF(...) / NN_pred(...) with the reference's argument lists
//               (TEncSearch.h:423-432, TEncSearch.cpp:85): a batch of one PU per call.  Bit-identical results,
//               useful for validation; throughput is bounded by the per-call round trip.
//   batched     enqueue(...) while the encoder walks a frame or CTU-row band, flush() once: the three GPU passes
//               run over every queued PU ("hard parts": the caller supplies the PU list).
//
// Failure behaviour mirrors the reference (assert/exit, TEncSearch.cpp:4330-4331): a failing fme_* call prints
// fme_last_error() and aborts.
#ifndef FME_HM_ADAPTOR_H
#define FME_HM_ADAPTOR_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "TLibCommon/CommonDef.h"
#include "TLibCommon/TComMv.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComRdCost.h"

#include "fme_b200.h"

class FmeHmAdaptor
{
public:
  FmeHmAdaptor() : m_ctx(NULL), m_orgPic(NULL) {}
  ~FmeHmAdaptor() { if (m_ctx) fme_destroy(m_ctx); }

  // TEncSearch::init (TEncSearch.cpp:377-1075): picture geometry + the QP that selects the NN weight set
  // (TEncSearch.cpp:472/625/775/925).  weightsDir = ".../DL/blowing" of the reference checkout.
  // biPred: also serve bi-predictive refinement calls (enqueueBi; random-access configurations)
  Void init(Int picWidth, Int picHeight, Int numRefSlots, Int maxPUsPerBatch, Bool useHadME, Bool fen, Int qp,
            const char* weightsDir, Int device = 0, Bool biPred = false)
  {
    fme_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = device; cfg.width = picWidth; cfg.height = picHeight; cfg.margin = 80; cfg.bitDepth = 8;
    cfg.numRefSlots = numRefSlots; cfg.maxPUs = maxPUsPerBatch; cfg.useHadME = useHadME; cfg.fen = fen;
    cfg.biPred = biPred;
    check(fme_create(&cfg, &m_ctx));
    const Int wq = (qp == 27 || qp == 32 || qp == 37) ? qp : 22;
    char dir[1024];
    snprintf(dir, sizeof(dir), "%s/%d", weightsDir, wq);
    check(fme_load_nn_csv_dir(m_ctx, dir));
    m_refPics.assign(numRefSlots, (const TComPicYuv*)NULL);
  }

  // once per slice: TComRdCost::setLambda (TComRdCost.cpp:104-117)
  Void setSliceLambda(Double lambda) { check(fme_set_slice(m_ctx, lambda)); }

  // once per coded picture: the source picture and every reference picture that changed
  Void setOrgPicture(const TComPicYuv* org)
  {
    m_orgPic = org;
    check(fme_upload_org(m_ctx, org->getAddr(COMPONENT_Y), org->getStride(COMPONENT_Y)));
  }
  Void setRefPicture(Int slot, const TComPicYuv* rec)
  {
    m_refPics[slot] = rec;
    check(fme_upload_ref(m_ctx, slot, rec->getAddr(COMPONENT_Y), rec->getStride(COMPONENT_Y)));
  }
  Int slotOf(const Pel* piRefY, Int iRefStride, Int& x, Int& y) const
  {
    for (size_t s = 0; s < m_refPics.size(); s++)
    {
      if (!m_refPics[s]) continue;
      const Pel* org = m_refPics[s]->getAddr(COMPONENT_Y);
      const ptrdiff_t off = piRefY - org;
      const Int w = m_refPics[s]->getWidth(COMPONENT_Y), h = m_refPics[s]->getHeight(COMPONENT_Y);
      // piRefY is the PU's collocated sample (TEncSearch.cpp:4481): inside the picture area of its plane
      ptrdiff_t yy = off >= 0 ? off / iRefStride : -1;
      ptrdiff_t xx = off - yy * iRefStride;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) { x = Int(xx); y = Int(yy); return Int(s); }
    }
    return -1;
  }

  // ---- batched mode --------------------------------------------------------------------------------------
  // What xMotionEstimation has in hand at TEncSearch.cpp:4534/4541: PU position and size, reference picture,
  // integer MV (rcMv), predictor (*pcMvPred), the 8 saved integer errors (array_e) and the centre error C.
  Int enqueue(Int puX, Int puY, Int width, Int height, Int refSlot, const TComMv& mvInt, const TComMv& mvPred,
              const UInt* arrayE8, UInt centreC, Bool lossless)
  {
    fme_pu p;
    memset(&p, 0, sizeof(p));
    p.x = Short(puX); p.y = Short(puY); p.w = UChar(width); p.h = UChar(height); p.refSlot = UChar(refSlot);
    p.flags = lossless ? FME_PU_LOSSLESS : 0;
    p.mvIntX = Short(mvInt.getHor()); p.mvIntY = Short(mvInt.getVer());
    p.mvPredX = Short(mvPred.getHor()); p.mvPredY = Short(mvPred.getVer());
    if (arrayE8)
    { // IN_errors << array_e[0..3], C, array_e[4..7]  (TEncSearch.cpp:88)
      for (Int i = 0; i < 4; i++) { p.err[i] = arrayE8[i]; p.err[5 + i] = arrayE8[4 + i]; }
      p.err[4] = centreC;
    }
    else
    {
      p.flags |= FME_PU_ERR_ON_GPU; // let the K0 pass compute the 3x3 surface (TEncSearch.cpp:5037-5050)
    }
    m_queue.push_back(p);
    return Int(m_queue.size()) - 1;
  }
  // xMotionEstimation with bBi (TEncSearch.cpp:4462-4472): the search pattern is 2*org - (other list's prediction).
  // Instead of the pattern buffer (m_cYuvPredTemp) the record names the other list's reference picture and the MV
  // xPredInterUni used for it (i.e. after pcCU->clipMv); the engine rebuilds the pattern on the device.
  Int enqueueBi(Int puX, Int puY, Int width, Int height, Int refSlot, const TComMv& mvInt, const TComMv& mvPred,
                Int otherRefSlot, const TComMv& otherMvClipped, Bool lossless)
  {
    fme_pu p;
    memset(&p, 0, sizeof(p));
    p.x = Short(puX); p.y = Short(puY); p.w = UChar(width); p.h = UChar(height); p.refSlot = UChar(refSlot);
    p.flags = UChar(FME_PU_BI | (lossless ? FME_PU_LOSSLESS : 0));
    p.mvIntX = Short(mvInt.getHor()); p.mvIntY = Short(mvInt.getVer());
    p.mvPredX = Short(mvPred.getHor()); p.mvPredY = Short(mvPred.getVer());
    p.err[0] = UInt(otherRefSlot) & 0xff;
    p.err[1] = (UInt(otherMvClipped.getHor()) & 0xffff) | (UInt(otherMvClipped.getVer()) << 16);
    m_queue.push_back(p);
    return Int(m_queue.size()) - 1;
  }
  Void flush(Int mode = FME_MODE_BOTH)
  {
    m_results.resize(m_queue.size());
    if (!m_queue.empty()) check(fme_submit(m_ctx, &m_queue[0], Int(m_queue.size()), &m_results[0], mode));
    m_queue.clear();
  }
  // The same flush over the bus-friendlier 44-byte records (fme_submit_compact): array_e / C travel as 24-bit values, the
  // PUs whose surface needs 32 bits (only possible above 256 luma samples) go in the full-grid list.  Uni-prediction
  // queues whose PUs all carry their surface; anything else falls back to flush().
  Void flushCompact(Int mode = FME_MODE_BOTH)
  {
    for (size_t i = 0; i < m_queue.size(); i++)
      if (m_queue[i].flags & (FME_PU_BI | FME_PU_ERR_ON_GPU)) { flush(mode); return; }
    m_results.resize(m_queue.size());
    m_compact.resize(m_queue.size());
    m_big.clear();
    for (size_t i = 0; i < m_queue.size(); i++)
      if (fme_pu_compact_pack(&m_queue[i], &m_compact[i]))
      {
        fme_err_grid g;
        g.pu = Int(i);
        for (Int k = 0; k < 9; k++) g.err[k] = m_queue[i].err[k];
        m_big.push_back(g);
      }
    if (!m_queue.empty())
      check(fme_submit_compact(m_ctx, &m_compact[0], Int(m_compact.size()), m_big.empty() ? NULL : &m_big[0], Int(m_big.size()),
                               &m_results[0], mode));
    m_queue.clear();
  }
  const fme_result& result(Int ticket) const { return m_results[ticket]; }

  // ---- immediate mode: the reference's own signatures ----------------------------------------------------
  // TEncSearch::xPatternSearchFracDIF (TEncSearch.h:423-432).  The caller has done what xMotionEstimation does
  // before the call (setPredictor, TEncSearch.cpp:4499); the predictor is passed explicitly because
  // TComRdCost::m_mvPredictor is private.  pcPatternKey must view the source picture set by setOrgPicture
  // (uni-prediction; bi-predictive refinement calls, whose pattern is not a picture view, go through enqueueBi).
  Void xPatternSearchFracDIF(Bool bIsLosslessCoded, TComPattern* pcPatternKey, Pel* piRefY, Int iRefStride,
                             TComMv* pcMvInt, const TComMv& mvPred, TComMv& rcMvHalf, TComMv& rcMvQter,
                             Distortion& ruiCost)
  {
    Int x = 0, y = 0;
    const Int slot = slotOf(piRefY, iRefStride, x, y);
    if (slot < 0) die("xPatternSearchFracDIF: piRefY does not point into a registered reference picture");
    const Int t = enqueue(x, y, pcPatternKey->getROIYWidth(), pcPatternKey->getROIYHeight(), slot, *pcMvInt, mvPred,
                          m_zeroErr, 0, bIsLosslessCoded);
    flush(FME_MODE_STD);
    const fme_result& r = result(t);
    rcMvHalf.set(r.halfX, r.halfY);
    rcMvQter.set(r.qterX, r.qterY);
    ruiCost = r.cost;
  }

  // NN_pred (TEncSearch.cpp:85-204) with its globals as arguments: array_e (8 values), C, PUHeight, PUWidth in;
  // MVX_HALF, MVX_QRTER, MVY_HALF, MVY_QRTER and NN_out out.
  Void NN_pred(const UInt* arrayE8, UInt centreC, UInt puHeight, UInt puWidth, Short& mvxHalf, Short& mvxQrter,
               Short& mvyHalf, Short& mvyQrter, Int& nnOut)
  {
    const TComMv zero(0, 0);
    const Int t = enqueue(0, 0, Int(puWidth), Int(puHeight), 0, zero, zero, arrayE8, centreC, false);
    flush(FME_MODE_NN);
    const fme_result& r = result(t);
    mvxHalf = r.nnHalfX; mvxQrter = r.nnQterX; mvyHalf = r.nnHalfY; mvyQrter = r.nnQterY; nnOut = r.nnClass;
  }

  // TComInterpolationFilter::filterHor / filterVer (TComInterpolationFilter.h:74-75)
  Void filterHor(const ComponentID compID, Pel* src, Int srcStride, Pel* dst, Int dstStride, Int width, Int height,
                 Int frac, Bool isLast, const ChromaFormat /*fmt = 4:2:0*/, const Int bitDepth)
  {
    check(fme_filter_hor(m_ctx, Int(compID), src, srcStride, dst, dstStride, width, height, frac, isLast, bitDepth));
  }
  Void filterVer(const ComponentID compID, Pel* src, Int srcStride, Pel* dst, Int dstStride, Int width, Int height,
                 Int frac, Bool isFirst, Bool isLast, const ChromaFormat /*fmt = 4:2:0*/, const Int bitDepth)
  {
    check(fme_filter_ver(m_ctx, Int(compID), src, srcStride, dst, dstStride, width, height, frac, isFirst, isLast,
                         bitDepth));
  }

  // FpDistFunc-compatible evaluation (TComRdCost.h:60): kind 0 = integer-ME metric (SSE / SAD12/24/48),
  // 1 = HADs, 2 = SADs; everything else is read from the DistParam exactly as the reference's functions do.
  Distortion distFunc(Int kind, DistParam* dp)
  {
    const Int w = dp->iCols, h = dp->iRows;
    std::vector<Pel> o(size_t(w) * h), c(size_t(w) * h);
    for (Int r = 0; r < h; r++)
    {
      memcpy(&o[size_t(r) * w], dp->pOrg + ptrdiff_t(r) * dp->iStrideOrg, w * sizeof(Pel));
      memcpy(&c[size_t(r) * w], dp->pCur + ptrdiff_t(r) * dp->iStrideCur, w * sizeof(Pel));
    }
    uint32_t out = 0;
    check(fme_dist(m_ctx, kind, &o[0], w, &c[0], w, w, h, dp->bitDepth, dp->iSubShift, 1, &out));
    return out;
  }

  fme_ctx* ctx() { return m_ctx; }

private:
  static Void die(const char* what)
  {
    fprintf(stderr, "fme_b200: %s\n", what);
    exit(1);
  }
  static Void check(int rc)
  {
    if (rc != FME_OK) die(fme_last_error());
  }
  fme_ctx*                       m_ctx;
  const TComPicYuv*              m_orgPic;
  std::vector<const TComPicYuv*> m_refPics;
  std::vector<fme_pu>            m_queue;
  std::vector<fme_result>        m_results;
  std::vector<fme_pu_compact>    m_compact;
  std::vector<fme_err_grid>      m_big;
  UInt                           m_zeroErr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

#endif
