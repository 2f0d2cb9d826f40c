// TEST INFRASTRUCTURE ONLY.  Executes hm16.9-nn_fme_b200/adaptor/fme_hm_adaptor.h against the reference's own objects
// (built by oracle/dropin/make_dropin.py, driven by tests/test_dropin.py on the GPU box).
//
//   adaptor_check <input.bin> <output.bin> <weights csv dir (parent of <qp>/)>
//
// input : int32 W, H, nRefs, nPUs, useHad, qp; double lambda; W*H u8 source; nRefs * W*H u8 references; nPUs fme_pu
// output: nPUs fme_result of the BATCHED path (enqueue + flush, mode BOTH) -- the caller compares them with the
//         reference encoder's captured outputs.
// In-process checks (exit code 1 on the first mismatch):
//   * slotOf() recovers (slot, x, y) from the piRefY pointer the reference passes (TEncSearch.cpp:4481)
//   * immediate xPatternSearchFracDIF / NN_pred (the reference's argument lists) == the batched results
//   * distFunc(kind, DistParam*) == the reference's DistParam::DistFunc (TComRdCost::setDistParam, HADs and SADs)
//   * filterHor / filterVer == TComInterpolationFilter::filterHor / filterVer on random blocks
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComInterpolationFilter.h"
#include "fme_hm_adaptor.h"

static void fail(const char* what, long i)
{
  fprintf(stderr, "adaptor_check: MISMATCH in %s at %ld\n", what, i);
  exit(1);
}

int main(int argc, char** argv)
{
  if (argc < 4) { fprintf(stderr, "usage: adaptor_check in.bin out.bin weightsDir\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int hdr[6];
  double lambda;
  if (fread(hdr, sizeof(int), 6, f) != 6 || fread(&lambda, sizeof(double), 1, f) != 1) return 2;
  const int W = hdr[0], H = hdr[1], nRefs = hdr[2], nPUs = hdr[3], useHad = hdr[4], qp = hdr[5];
  initROM();

  std::vector<TComPicYuv*> pics(nRefs + 1);
  std::vector<unsigned char> row(W);
  for (int p = 0; p <= nRefs; p++)
  { // picture 0 = source, 1.. = references; TComPicYuv as the encoder holds them (Pel planes with margins)
    pics[p] = new TComPicYuv;
    pics[p]->create(W, H, CHROMA_420, 64, 64, 4, true);
    Pel* y = pics[p]->getAddr(COMPONENT_Y);
    const Int s = pics[p]->getStride(COMPONENT_Y);
    for (int r = 0; r < H; r++)
    {
      if (fread(&row[0], 1, W, f) != (size_t)W) return 2;
      for (int x = 0; x < W; x++) y[r * s + x] = row[x];
    }
    if (p > 0) pics[p]->extendPicBorder();
  }
  std::vector<fme_pu> pus(nPUs);
  if (fread(&pus[0], sizeof(fme_pu), nPUs, f) != (size_t)nPUs) return 2;
  fclose(f);

  FmeHmAdaptor fme;
  fme.init(W, H, nRefs, nPUs, useHad != 0, true, qp, argv[3]);
  fme.setSliceLambda(lambda);
  fme.setOrgPicture(pics[0]);
  for (int s = 0; s < nRefs; s++) fme.setRefPicture(s, pics[s + 1]);

  // ---- batched: one ticket per PU, one flush ----
  std::vector<Int> ticket(nPUs);
  for (int i = 0; i < nPUs; i++)
  {
    const fme_pu& p = pus[i];
    UInt e8[8];
    for (int k = 0; k < 4; k++) { e8[k] = p.err[k]; e8[4 + k] = p.err[5 + k]; }
    ticket[i] = fme.enqueue(p.x, p.y, p.w, p.h, p.refSlot, TComMv(p.mvIntX, p.mvIntY), TComMv(p.mvPredX, p.mvPredY), e8,
                            p.err[4], (p.flags & FME_PU_LOSSLESS) != 0);
  }
  fme.flush(FME_MODE_BOTH);
  std::vector<fme_result> batched(nPUs);
  for (int i = 0; i < nPUs; i++) batched[i] = fme.result(ticket[i]);
  // ---- the same batch through the 44-byte compact records (flushCompact): identical results ----
  for (int i = 0; i < nPUs; i++)
  {
    const fme_pu& p = pus[i];
    UInt e8[8];
    for (int k = 0; k < 4; k++) { e8[k] = p.err[k]; e8[4 + k] = p.err[5 + k]; }
    ticket[i] = fme.enqueue(p.x, p.y, p.w, p.h, p.refSlot, TComMv(p.mvIntX, p.mvIntY), TComMv(p.mvPredX, p.mvPredY), e8,
                            p.err[4], (p.flags & FME_PU_LOSSLESS) != 0);
  }
  fme.flushCompact(FME_MODE_BOTH);
  for (int i = 0; i < nPUs; i++)
  {
    const fme_result& a = batched[i];
    const fme_result& b = fme.result(ticket[i]);
    if (a.halfX != b.halfX || a.halfY != b.halfY || a.qterX != b.qterX || a.qterY != b.qterY || a.cost != b.cost ||
        a.nnHalfX != b.nnHalfX || a.nnHalfY != b.nnHalfY || a.nnQterX != b.nnQterX || a.nnQterY != b.nnQterY || a.nnClass != b.nnClass)
      fail("flushCompact vs flush", i);
  }
  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 2; }
  fwrite(&batched[0], sizeof(fme_result), nPUs, o);
  fclose(o);

  // ---- slotOf + immediate mode on a sample of the list ----
  const int step = nPUs > 400 ? nPUs / 400 : 1;
  long nImm = 0;
  for (int i = 0; i < nPUs; i += step)
  {
    const fme_pu& p = pus[i];
    TComPicYuv* ref = pics[p.refSlot + 1];
    const Int rs = ref->getStride(COMPONENT_Y);
    Pel* piRefY = ref->getAddr(COMPONENT_Y) + p.y * rs + p.x;
    Int sx = -1, sy = -1;
    if (fme.slotOf(piRefY, rs, sx, sy) != p.refSlot || sx != p.x || sy != p.y) fail("slotOf", i);
    TComPattern key;
    key.initPattern(pics[0]->getAddr(COMPONENT_Y) + p.y * pics[0]->getStride(COMPONENT_Y) + p.x, p.w, p.h,
                    pics[0]->getStride(COMPONENT_Y), 8);
    TComMv mvInt(p.mvIntX, p.mvIntY), mvHalf, mvQter;
    Distortion cost = 0;
    fme.xPatternSearchFracDIF((p.flags & FME_PU_LOSSLESS) != 0, &key, piRefY, rs, &mvInt, TComMv(p.mvPredX, p.mvPredY),
                              mvHalf, mvQter, cost);
    const fme_result& b = batched[i];
    if (mvHalf.getHor() != b.halfX || mvHalf.getVer() != b.halfY || mvQter.getHor() != b.qterX ||
        mvQter.getVer() != b.qterY || cost != b.cost) fail("immediate xPatternSearchFracDIF vs batched", i);
    UInt e8[8];
    for (int k = 0; k < 4; k++) { e8[k] = p.err[k]; e8[4 + k] = p.err[5 + k]; }
    Short hx, qx, hy, qy;
    Int cls;
    fme.NN_pred(e8, p.err[4], p.h, p.w, hx, qx, hy, qy, cls);
    if (hx != b.nnHalfX || qx != b.nnQterX || hy != b.nnHalfY || qy != b.nnQterY || cls != b.nnClass)
      fail("immediate NN_pred vs batched", i);
    nImm++;
  }

  // ---- distFunc against the reference's DistParam::DistFunc ----
  TComRdCost rd;
  rd.init();
  long nDist = 0;
  static const int shapes[][2] = {{8, 8}, {16, 16}, {32, 32}, {64, 64}, {8, 4}, {4, 8}, {16, 4}, {16, 8}, {32, 16}, {64, 32}};
  for (size_t k = 0; k < sizeof(shapes) / sizeof(shapes[0]); k++)
  {
    const int w = shapes[k][0], h = shapes[k][1];
    const Pel* a = pics[0]->getAddr(COMPONENT_Y) + 16 * pics[0]->getStride(COMPONENT_Y) + 24 + (int)k;
    const Pel* b = pics[1]->getAddr(COMPONENT_Y) + 17 * pics[1]->getStride(COMPONENT_Y) + 21 + 2 * (int)k;
    for (int had = 0; had < 2; had++)
    {
      DistParam dp;
      rd.setDistParam(dp, 8, a, pics[0]->getStride(COMPONENT_Y), b, pics[1]->getStride(COMPONENT_Y), w, h, had != 0);
      const Distortion want = dp.DistFunc(&dp);
      const Distortion got = fme.distFunc(had ? 1 : 2, &dp);
      if (want != got) fail(had ? "distFunc HADs" : "distFunc SADs", (long)k);
      nDist++;
    }
  }

  // ---- filterHor / filterVer against TComInterpolationFilter ----
  TComInterpolationFilter filt;
  long nFilt = 0;
  Pel* src = pics[1]->getAddr(COMPONENT_Y) + 40 * pics[1]->getStride(COMPONENT_Y) + 40;
  const Int ss = pics[1]->getStride(COMPONENT_Y);
  std::vector<Pel> want(64 * 72), got(64 * 72), tmpW(64 * 80), tmpG(64 * 80);
  for (int frac = 0; frac < 4; frac++)
  {
    filt.filterHor(COMPONENT_Y, src, ss, &want[0], 64, 48, 24, frac, true, CHROMA_420, 8);
    fme.filterHor(COMPONENT_Y, src, ss, &got[0], 64, 48, 24, frac, true, CHROMA_420, 8);
    if (memcmp(&want[0], &got[0], sizeof(Pel) * 64 * 24)) fail("filterHor isLast", frac);
    // two-stage path as xExtDIFUpSamplingH runs it: horizontal first stage, vertical last stage
    filt.filterHor(COMPONENT_Y, src - 3 * ss, ss, &tmpW[0], 64, 32, 24 + 7, frac, false, CHROMA_420, 8);
    fme.filterHor(COMPONENT_Y, src - 3 * ss, ss, &tmpG[0], 64, 32, 24 + 7, frac, false, CHROMA_420, 8);
    if (memcmp(&tmpW[0], &tmpG[0], sizeof(Pel) * 64 * 31)) fail("filterHor first stage", frac);
    for (int fy = 0; fy < 4; fy++)
    {
      filt.filterVer(COMPONENT_Y, &tmpW[3 * 64], 64, &want[0], 64, 32, 24, fy, false, true, CHROMA_420, 8);
      fme.filterVer(COMPONENT_Y, &tmpG[3 * 64], 64, &got[0], 64, 32, 24, fy, false, true, CHROMA_420, 8);
      if (memcmp(&want[0], &got[0], sizeof(Pel) * 64 * 24)) fail("filterVer last stage", frac * 4 + fy);
      nFilt++;
    }
  }
  printf("adaptor_check ok: %d batched PUs, %ld immediate calls, %ld distFunc cases, %ld filter cases\n", nPUs, nImm, nDist,
         nFilt);
  return 0;
}
