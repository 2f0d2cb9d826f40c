/* TEST INFRASTRUCTURE ONLY -- see fme_oracle.h.  Plain-C restatement of the reference's
 * fractional-ME hot path, written from the algorithm (not from the reference's text); every
 * function cites the reference lines whose behaviour it restates.
 *
 * Reference files (relative to /root/reference/source/Lib):
 *   IF  = TLibCommon/TComInterpolationFilter.cpp      RD  = TLibCommon/TComRdCost.cpp / .h
 *   TES = TLibEncoder/TEncSearch.cpp                   SLC = TLibEncoder/TEncSlice.cpp
 */
#include "fme_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define INTERNAL_PREC 14 /* IF.h:49 IF_INTERNAL_PREC */
#define FILTER_PREC 6    /* IF.h:50 IF_FILTER_PREC   */
#define INTERNAL_OFFS (1 << (INTERNAL_PREC - 1)) /* IF.h:51 */

/* IF:57-75 */
static const int kLuma[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0},
                                {-1, 4, -10, 58, 17, -5, 1, 0},
                                {-1, 4, -11, 40, 40, -11, 4, -1},
                                {0, 1, -5, 17, 58, -10, 4, -1}};
static const int kChroma[8][4] = {{0, 64, 0, 0},   {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                  {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

static int imax(int a, int b) { return a > b ? a : b; }

/* IF:94-154 filterCopy.  Every store goes through a 16-bit Pel, as in the reference. */
static void filter_copy(int bitDepth, const orc_pel* src, int ss, orc_pel* dst, int ds, int w, int h, int isFirst,
                        int isLast) {
  int shift = imax(2, INTERNAL_PREC - bitDepth);
  for (int r = 0; r < h; ++r, src += ss, dst += ds)
    for (int c = 0; c < w; ++c) {
      if (isFirst == isLast) {
        dst[c] = src[c];
      } else if (isFirst) {
        orc_pel v = (orc_pel)((int)src[c] << shift);
        dst[c] = (orc_pel)(v - INTERNAL_OFFS);
      } else {
        int t = (int)src[c] + INTERNAL_OFFS;
        orc_pel v = (orc_pel)((t + (1 << (shift - 1))) >> shift);
        int maxVal = (1 << bitDepth) - 1;
        if (v < 0) v = 0;
        if (v > maxVal) v = (orc_pel)maxVal;
        dst[c] = v;
      }
    }
}

/* IF:172-257 filter<N,isVertical,isFirst,isLast> */
static void filter_fir(int ntaps, int isVertical, int isFirst, int isLast, int bitDepth, const orc_pel* src, int ss,
                       orc_pel* dst, int ds, int w, int h, const int* coeff) {
  int cStride = isVertical ? ss : 1;
  src -= (ntaps / 2 - 1) * cStride;
  int headRoom = imax(2, INTERNAL_PREC - bitDepth);
  int shift = FILTER_PREC, offset, maxVal;
  if (isLast) {
    shift += isFirst ? 0 : headRoom;
    offset = 1 << (shift - 1);
    offset += isFirst ? 0 : (INTERNAL_OFFS << FILTER_PREC);
    maxVal = (1 << bitDepth) - 1;
  } else {
    shift -= isFirst ? headRoom : 0;
    offset = isFirst ? -(INTERNAL_OFFS << shift) : 0;
    maxVal = 0;
  }
  for (int r = 0; r < h; ++r, src += ss, dst += ds)
    for (int c = 0; c < w; ++c) {
      int sum = 0;
      for (int k = 0; k < ntaps; ++k) sum += (int)src[c + k * cStride] * coeff[k];
      orc_pel v = (orc_pel)((sum + offset) >> shift); /* Pel val = (sum + offset) >> shift; IF:245 */
      if (isLast) {
        if (v < 0) v = 0;
        if (v > maxVal) v = (orc_pel)maxVal;
      }
      dst[c] = v;
    }
}

/* IF:341-358 (4:2:0: chroma phase index = frac, csx = 1) */
void orc_filter_hor(int isLuma, const orc_pel* src, int ss, orc_pel* dst, int ds, int w, int h, int frac, int isLast,
                    int bitDepth) {
  if (frac == 0) filter_copy(bitDepth, src, ss, dst, ds, w, h, 1, isLast);
  else if (isLuma) filter_fir(8, 0, 1, isLast, bitDepth, src, ss, dst, ds, w, h, kLuma[frac]);
  else filter_fir(4, 0, 1, isLast, bitDepth, src, ss, dst, ds, w, h, kChroma[frac]);
}

/* IF:377-394 */
void orc_filter_ver(int isLuma, const orc_pel* src, int ss, orc_pel* dst, int ds, int w, int h, int frac, int isFirst,
                    int isLast, int bitDepth) {
  if (frac == 0) filter_copy(bitDepth, src, ss, dst, ds, w, h, isFirst, isLast);
  else if (isLuma) filter_fir(8, 1, isFirst, isLast, bitDepth, src, ss, dst, ds, w, h, kLuma[frac]);
  else filter_fir(4, 1, isFirst, isLast, bitDepth, src, ss, dst, ds, w, h, kChroma[frac]);
}

/* TComYuv::addAvg, one component (TComYuv.cpp:354-409).  The reference unrolls by 2 or 4 columns; the per-sample
 * expression is the same in every branch.  rightShift(x, s) is an arithmetic >> for s >= 0 (CommonDef.h). */
void orc_add_avg(const orc_pel* src0, int st0, const orc_pel* src1, int st1, orc_pel* dst, int ds, int w, int h,
                 int bitDepth) {
  const int headRoom = 14 - bitDepth;
  const int shiftNum = (headRoom > 2 ? headRoom : 2) + 1;
  const int offset = (1 << (shiftNum - 1)) + 2 * 8192;
  const int maxv = (1 << bitDepth) - 1;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int v = ((int)src0[y * st0 + x] + (int)src1[y * st1 + x] + offset) >> shiftNum;
      dst[y * ds + x] = (orc_pel)(v < 0 ? 0 : v > maxv ? maxv : v);
    }
}

/* ---------------------------------------------------------------- distortion */

/* RD:359-855.  Widths 4..64 (and 12/24/48): rows subsampled by 1<<subShift, sum <<= subShift.
 * (The generic xGetSAD early exit, RD:380, cannot trigger on this path: the FME callers leave
 * m_maximumDistortionForEarlyExit at max when the fixed-width functions are selected.) */
uint32_t orc_sad(const orc_pel* org, int os, const orc_pel* cur, int cs, int w, int h, int bitDepth, int subShift) {
  uint32_t sum = 0;
  int step = 1 << subShift;
  for (int r = 0; r < h; r += step)
    for (int c = 0; c < w; ++c) sum += (uint32_t)abs((int)org[r * os + c] - (int)cur[r * cs + c]);
  sum <<= subShift;
  return sum >> (bitDepth - 8);
}

/* RD:861-1206: no subsampling, per-sample >> 2*(bitDepth-8) */
uint32_t orc_sse(const orc_pel* org, int os, const orc_pel* cur, int cs, int w, int h, int bitDepth) {
  uint32_t sum = 0;
  int sh = (bitDepth - 8) << 1;
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < w; ++c) {
      int d = (int)org[r * os + c] - (int)cur[r * cs + c];
      sum += (uint32_t)((d * d) >> sh);
    }
  return sum;
}

/* In-place unnormalised Hadamard of length n (n = 2,4,8) on a strided vector.  The reference's
 * butterflies (RD:1212-1425) produce the same multiset of coefficients up to sign/order, and
 * only sum |coef| is consumed. */
static void wht(int* v, int stride, int n) {
  for (int len = 1; len < n; len <<= 1)
    for (int i = 0; i < n; i += len << 1)
      for (int j = i; j < i + len; ++j) {
        int a = v[j * stride], b = v[(j + len) * stride];
        v[j * stride] = a + b;
        v[(j + len) * stride] = a - b;
      }
}

static uint32_t had_tile(const orc_pel* org, int os, const orc_pel* cur, int cs, int n) {
  int d[64];
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) d[r * n + c] = (int)org[r * os + c] - (int)cur[r * cs + c];
  for (int r = 0; r < n; ++r) wht(d + r * n, 1, n);
  for (int c = 0; c < n; ++c) wht(d + c, n, n);
  uint32_t s = 0;
  for (int i = 0; i < n * n; ++i) s += (uint32_t)abs(d[i]);
  if (n == 8) return (s + 2) >> 2; /* RD:1421 */
  if (n == 4) return (s + 1) >> 1; /* RD:1325 */
  return s;                        /* RD:1226-1231 */
}

/* RD:1428-1495 xGetHADs: tile size by divisibility, per-tile rounding before the sum */
uint32_t orc_hads(const orc_pel* org, int os, const orc_pel* cur, int cs, int w, int h, int bitDepth) {
  int n = (h % 8 == 0 && w % 8 == 0) ? 8 : (h % 4 == 0 && w % 4 == 0) ? 4 : 2;
  uint32_t sum = 0;
  for (int y = 0; y < h; y += n)
    for (int x = 0; x < w; x += n) sum += had_tile(org + y * os + x, os, cur + y * cs + x, cs, n);
  return sum >> (bitDepth - 8);
}

/* RD:200-229: this fork selects DF_SSE(+width) for the integer search; 12/24/48 keep SAD */
uint32_t orc_int_me_dist(const orc_pel* org, int os, const orc_pel* cur, int cs, int w, int h, int bitDepth,
                         int subShift) {
  if (w == 12 || w == 24 || w == 48) return orc_sad(org, os, cur, cs, w, h, bitDepth, subShift);
  return orc_sse(org, os, cur, cs, w, h, bitDepth);
}

uint32_t orc_dist(int kind, const orc_pel* org, int os, const orc_pel* cur, int cs, int w, int h, int bitDepth,
                  int subShift) {
  if (kind == 0) return orc_int_me_dist(org, os, cur, cs, w, h, bitDepth, subShift);
  if (kind == 1) return orc_hads(org, os, cur, cs, w, h, bitDepth);
  return orc_sad(org, os, cur, cs, w, h, bitDepth, subShift);
}

/* ---------------------------------------------------------------- MV-bit cost */

/* RD:172-185 */
uint32_t orc_exp_golomb_bits(int v) {
  uint32_t len = 1;
  uint32_t t = (v <= 0) ? (((uint32_t)(-v)) << 1) + 1 : ((uint32_t)v << 1);
  while (t != 1) {
    t >>= 1;
    len += 2;
  }
  return len;
}

/* RD:104-117: m_dLambdaMotionSAD[0] = 65536*sqrt(lambda); selectMotionLambda(true,0,false) picks it */
double orc_motion_lambda(double lambda) { return 65536.0 * sqrt(lambda); }

/* RD.h:165-174 */
uint32_t orc_mv_cost(double motionLambda, int x, int y, int scale, int predX, int predY) {
  uint32_t bits = orc_exp_golomb_bits((x << scale) - predX) + orc_exp_golomb_bits((y << scale) - predY);
  return (uint32_t)((motionLambda * bits) / 65536.0);
}

/* SLC:290-325 */
double orc_slice_lambda(int qp, double qpFactor, int depth, int hadME) {
  double qpTemp = (double)qp - 12.0;
  double lambda = qpFactor * pow(2.0, qpTemp / 3.0);
  if (depth > 0) {
    double f = qpTemp / 6.0;
    lambda *= f < 2.0 ? 2.0 : (f > 4.0 ? 4.0 : f);
  }
  if (!hadME) lambda *= 0.95;
  return lambda;
}

/* ---------------------------------------------------------------- fractional search */

#define FB_STRIDE 80 /* TComPrediction.cpp:136-146: (MAX_CU_SIZE+16) wide scratch */
static orc_pel g_fb[4][4][FB_STRIDE * 66];
static orc_pel g_tmp[4][FB_STRIDE * 73];

const orc_pel* orc_filtered_block(int v, int h) { return g_fb[v][h]; }

/* TES:6331-6365 */
static void up_sampling_h(const orc_pel* roi, int ss, int w, int h) {
  const orc_pel* src = roi - 4 * ss - 1;
  orc_filter_hor(1, src, ss, g_tmp[0], FB_STRIDE, w + 1, h + 8, 0, 0, 8);
  orc_filter_hor(1, src, ss, g_tmp[2], FB_STRIDE, w + 1, h + 8, 2, 0, 8);
  orc_filter_ver(1, g_tmp[0] + 4 * FB_STRIDE + 1, FB_STRIDE, g_fb[0][0], FB_STRIDE, w, h, 0, 0, 1, 8);
  orc_filter_ver(1, g_tmp[0] + 3 * FB_STRIDE + 1, FB_STRIDE, g_fb[2][0], FB_STRIDE, w, h + 1, 2, 0, 1, 8);
  orc_filter_ver(1, g_tmp[2] + 4 * FB_STRIDE, FB_STRIDE, g_fb[0][2], FB_STRIDE, w + 1, h, 0, 0, 1, 8);
  orc_filter_ver(1, g_tmp[2] + 3 * FB_STRIDE, FB_STRIDE, g_fb[2][2], FB_STRIDE, w + 1, h + 1, 2, 0, 1, 8);
}

/* TES:6378-6532 */
static void up_sampling_q(const orc_pel* roi, int ss, int w, int h, int hx, int hy) {
  int extH = (hy == 0) ? h + 8 : h + 7;
  const orc_pel* base = roi - 4 * ss - 1;
  const orc_pel* s1 = base + (hy > 0 ? ss : 0) + (hx >= 0 ? 1 : 0);
  const orc_pel* s3 = base + (hy > 0 ? ss : 0) + (hx > 0 ? 1 : 0);
  orc_filter_hor(1, s1, ss, g_tmp[1], FB_STRIDE, w, extH, 1, 0, 8);
  orc_filter_hor(1, s3, ss, g_tmp[3], FB_STRIDE, w, extH, 3, 0, 8);

  const int S = FB_STRIDE;
  /* (1,1) and (3,1) */
  orc_filter_ver(1, g_tmp[1] + 3 * S + (hy == 0 ? S : 0), S, g_fb[1][1], S, w, h, 1, 0, 1, 8);
  orc_filter_ver(1, g_tmp[1] + 3 * S, S, g_fb[3][1], S, w, h, 3, 0, 1, 8);
  if (hy != 0) { /* (2,1), (2,3) */
    orc_filter_ver(1, g_tmp[1] + 3 * S, S, g_fb[2][1], S, w, h, 2, 0, 1, 8);
    orc_filter_ver(1, g_tmp[3] + 3 * S, S, g_fb[2][3], S, w, h, 2, 0, 1, 8);
  } else { /* (0,1), (0,3) */
    orc_filter_ver(1, g_tmp[1] + 4 * S, S, g_fb[0][1], S, w, h, 0, 0, 1, 8);
    orc_filter_ver(1, g_tmp[3] + 4 * S, S, g_fb[0][3], S, w, h, 0, 0, 1, 8);
  }
  if (hx != 0) { /* (1,2), (3,2) from the half-pel horizontal intermediate */
    orc_filter_ver(1, g_tmp[2] + 3 * S + (hx > 0 ? 1 : 0) + (hy >= 0 ? S : 0), S, g_fb[1][2], S, w, h, 1, 0, 1, 8);
    orc_filter_ver(1, g_tmp[2] + 3 * S + (hx > 0 ? 1 : 0) + (hy > 0 ? S : 0), S, g_fb[3][2], S, w, h, 3, 0, 1, 8);
  } else { /* (1,0), (3,0) from the full-pel horizontal intermediate */
    orc_filter_ver(1, g_tmp[0] + 3 * S + 1 + (hy >= 0 ? S : 0), S, g_fb[1][0], S, w, h, 1, 0, 1, 8);
    orc_filter_ver(1, g_tmp[0] + 3 * S + 1 + (hy > 0 ? S : 0), S, g_fb[3][0], S, w, h, 3, 0, 1, 8);
  }
  /* (1,3) and (3,3) */
  orc_filter_ver(1, g_tmp[3] + 3 * S + (hy == 0 ? S : 0), S, g_fb[1][3], S, w, h, 1, 0, 1, 8);
  orc_filter_ver(1, g_tmp[3] + 3 * S, S, g_fb[3][3], S, w, h, 3, 0, 1, 8);
}

/* TES:212-236 */
static const int kRefineH[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
static const int kRefineQ[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

/* TES:1591-1645.  mvFrac: in = MV used for the bit cost, out = winning offset. */
static uint32_t pattern_refinement(const orc_pel* org, int os, int w, int h, int baseX, int baseY, int iFrac,
                                   int mvFrac[2], int useHad, double motionLambda, int costScale, int predX,
                                   int predY) {
  const int(*tab)[2] = (iFrac == 2) ? kRefineH : kRefineQ;
  uint32_t best = 0xffffffffu;
  int bestI = 0;
  for (int i = 0; i < 9; ++i) {
    int hor = (tab[i][0] + baseX) * iFrac, ver = (tab[i][1] + baseY) * iFrac;
    const orc_pel* p = g_fb[ver & 3][hor & 3];
    if (hor == 2 && (ver & 1) == 0) p += 1;
    if ((hor & 1) == 0 && ver == 2) p += FB_STRIDE;
    int tx = tab[i][0] + mvFrac[0], ty = tab[i][1] + mvFrac[1];
    uint32_t d = useHad ? orc_hads(org, os, p, FB_STRIDE, w, h, 8) : orc_sad(org, os, p, FB_STRIDE, w, h, 8, 0);
    d += orc_mv_cost(motionLambda, tx, ty, costScale, predX, predY);
    if (d < best) {
      best = d;
      bestI = i;
    }
  }
  mvFrac[0] = tab[bestI][0];
  mvFrac[1] = tab[bestI][1];
  return best;
}

/* TES:5232-5269 with the caller's cost-scale protocol (TES:4531 scale 1, TES:5260 scale 0) */
void orc_frac_dif(const orc_pel* org, int os, int w, int h, const orc_pel* ref, int rs, int mvIntX, int mvIntY,
                  int predX, int predY, double lambda, int useHad, int lossless, int16_t halfXY[2], int16_t qterXY[2],
                  uint32_t* cost) {
  const double ml = orc_motion_lambda(lambda);
  const orc_pel* roi = ref + mvIntX + mvIntY * rs;
  int had = useHad && !lossless; /* TES:1604, 5258 */

  up_sampling_h(roi, rs, w, h);
  int mvH[2] = {mvIntX << 1, mvIntY << 1};
  pattern_refinement(org, os, w, h, 0, 0, 2, mvH, had, ml, 1, predX, predY);

  up_sampling_q(roi, rs, w, h, mvH[0], mvH[1]);
  int mvQ[2] = {((mvIntX << 1) + mvH[0]) << 1, ((mvIntY << 1) + mvH[1]) << 1};
  *cost = pattern_refinement(org, os, w, h, mvH[0] << 1, mvH[1] << 1, 1, mvQ, had, ml, 0, predX, predY);
  halfXY[0] = (int16_t)mvH[0];
  halfXY[1] = (int16_t)mvH[1];
  qterXY[0] = (int16_t)mvQ[0];
  qterXY[1] = (int16_t)mvQ[1];
}

/* P[fy][fx] by the same two block-filter calls the reference makes
 * (filterHor isLast=false over h+7 rows, then filterVer isFirst=false isLast=true). */
void orc_subpel_plane(const orc_pel* ref, int rs, int x0, int y0, int w, int h, int fy, int fx, orc_pel* out, int outS) {
  orc_pel* tmp = (orc_pel*)malloc(sizeof(orc_pel) * (size_t)w * (size_t)(h + 7));
  orc_filter_hor(1, ref + (y0 - 3) * rs + x0, rs, tmp, w, w, h + 7, fx, 0, 8);
  orc_filter_ver(1, tmp + 3 * w, w, out, outS, w, h, fy, 0, 1, 8);
  free(tmp);
}

/* TES:1085-1090, 1156-1166 (FEN row subsampling), TES:1324-1377 (point order 1..8 = raster minus centre) */
void orc_int_surface(const orc_pel* org, int os, int w, int h, const orc_pel* refAtMv, int rs, int fen,
                     uint32_t err9[9]) {
  int k = 0;
  int sub = (fen && h > 8) ? 1 : 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) err9[k++] = orc_int_me_dist(org, os, refAtMv + dy * rs + dx, rs, w, h, 8, sub);
}

/* ---------------------------------------------------------------- NN_pred */

size_t orc_nn_blob_floats(const orc_nn_header* h) {
  size_t n = 3 * (size_t)h->nErr + (size_t)h->nEmb * h->embRows * h->embDim;
  int in = h->nErr + h->nEmb * h->embDim;
  for (int l = 0; l < h->nHidden; ++l) {
    n += (size_t)h->hidden[l] * in + 3 * (size_t)h->hidden[l];
    in = h->hidden[l];
  }
  n += (size_t)h->nOut * in + h->nOut;
  return n;
}

/* TES:93-113.  Height maps 16->3, 12->4; width maps 12->3, 16->4 (sic, master behaviour). */
static int emb_index(int v, int isHeight) {
  switch (v) {
    case 4: return 1;
    case 8: return 2;
    case 16: return isHeight ? 3 : 4;
    case 12: return isHeight ? 4 : 3;
    case 24: return 5;
    case 32: return 6;
    case 64: return 7;
    default: return 0;
  }
}

/* TES:134-193: class k -> (qx,qy) = (k%7-3, k/7-3); per axis -3..3 -> (half,quarter) */
void orc_nn_class_to_mv(int cls, int16_t halfXY[2], int16_t qterXY[2]) {
  static const int kHalf[7] = {-1, -1, 0, 0, 0, 1, 1};
  static const int kQter[7] = {-1, 0, -1, 0, 1, 0, 1};
  if (cls < 0 || cls > 48) { /* TES:193 default */
    halfXY[0] = halfXY[1] = qterXY[0] = qterXY[1] = 0;
    return;
  }
  halfXY[0] = (int16_t)kHalf[cls % 7];
  qterXY[0] = (int16_t)kQter[cls % 7];
  halfXY[1] = (int16_t)kHalf[cls / 7];
  qterXY[1] = (int16_t)kQter[cls / 7];
}

/* TES:85-134, float32 throughout, one rounding per operation, ascending-k dot products
 * (the order the eigen_standin build of the reference uses). */
int orc_nn_pred(const void* blob, const uint32_t err9[9], int puHeight, int puWidth, float* logits,
                int16_t halfXY[2], int16_t qterXY[2]) {
  const orc_nn_header* H = (const orc_nn_header*)blob;
  const float* p = (const float*)(H + 1);
  const float* mean = p; p += H->nErr;
  const float* stdev = p; p += H->nErr;
  const float* gin = p; p += H->nErr;
  float x[64 + 16], y[64 + 16];
  int n = 0;
  if (H->nEmb == 2) {
    const float* e0 = p + (size_t)emb_index(puHeight, 1) * H->embDim;
    const float* e1 = p + (size_t)H->embRows * H->embDim + (size_t)emb_index(puWidth, 0) * H->embDim;
    for (int i = 0; i < H->embDim; ++i) x[n++] = e0[i];
    for (int i = 0; i < H->embDim; ++i) x[n++] = e1[i];
  }
  p += (size_t)H->nEmb * H->embRows * H->embDim;
  for (int i = 0; i < H->nErr; ++i) {
    float e = (float)err9[i];      /* TES:88  uint -> float */
    e = (e - mean[i]) / stdev[i];  /* TES:89 */
    x[n++] = e * gin[i];           /* TES:116 */
  }
  for (int l = 0; l < H->nHidden; ++l) {
    int out = H->hidden[l];
    const float* W = p; p += (size_t)out * n;
    const float* b = p; p += out;
    const float* g = p; p += out;
    const float* be = p; p += out;
    for (int o = 0; o < out; ++o) {
      float acc = W[o * n] * x[0];
      for (int k = 1; k < n; ++k) acc = acc + W[o * n + k] * x[k];
      acc = acc + b[o];
      acc = acc < 0.0f ? 0.0f : acc;      /* TES:122 relu */
      y[o] = (acc * g[o]) + be[o];
    }
    memcpy(x, y, sizeof(float) * (size_t)out);
    n = out;
  }
  const float* W = p; p += (size_t)H->nOut * n;
  const float* b = p;
  int best = 0;
  float bestV = 0.0f;
  for (int o = 0; o < H->nOut; ++o) {
    float acc = W[o * n] * x[0];
    for (int k = 1; k < n; ++k) acc = acc + W[o * n + k] * x[k];
    acc = acc + b[o];
    /* outSigmoid (the 3-layer backup network, Backups/4...cpp:4476): the sigmoid is monotonic, so the class is decided
     * on the pre-activation (first maximum); the reported outputs carry the activation */
    if (logits) logits[o] = H->outSigmoid ? 1.0f / (1.0f + expf(-acc)) : acc;
    /* the reference evaluates that sigmoid in double, where it is exactly 1.0 from 53 ln 2 = 36.7368 on: saturated
     * outputs tie and std::max_element keeps the first of them -- the clamp reproduces that */
    if (H->outSigmoid && acc > 36.7368f) acc = 36.7368f;
    if (o == 0 || acc > bestV) { /* first maximum, TES:134 */
      bestV = acc;
      best = o;
    }
  }
  if (halfXY && qterXY) orc_nn_class_to_mv(best, halfXY, qterXY);
  return best;
}

/* The reference's own 3-layer network, restated as it computes: double precision throughout
 * (source/Lib/TLibEncoder/Backups/4. TEncSearch - SCR 3 layers - no normalization.cpp:4408-4490):
 *   IN[i] = (E_i - mean_i) / stdev_i (:4427-4435), IN_norm = IN * BN_gamma_in (:4439-4441),
 *   X = relu(W X + b) * gamma + beta per hidden layer (:4445-4471, accumulation in index order starting from 0),
 *   OUT = sigmoid(W X + b) (:4475-4481), class = first maximum (std::max_element, :4486).
 * `payload` holds the same sequence as an FMNN blob's payload, as doubles (tests/golden/backup3_*.npz). */
int orc_nn_pred_f64(const void* header, const double* payload, const uint32_t err9[9], double* outs) {
  const orc_nn_header* H = (const orc_nn_header*)header;
  const double* p = payload;
  const double* mean = p; p += H->nErr;
  const double* stdev = p; p += H->nErr;
  const double* gin = p; p += H->nErr;
  double x[64], y[64];
  int n = H->nErr;
  for (int i = 0; i < n; ++i) x[i] = (((double)err9[i] - mean[i]) / stdev[i]) * gin[i];
  for (int l = 0; l < H->nHidden; ++l) {
    int out = H->hidden[l];
    const double* W = p; p += (size_t)out * n;
    const double* b = p; p += out;
    const double* g = p; p += out;
    const double* be = p; p += out;
    for (int o = 0; o < out; ++o) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += W[o * n + k] * x[k];
      acc += b[o];
      y[o] = ((acc > 0.0 ? acc : 0.0) * g[o]) + be[o];
    }
    memcpy(x, y, sizeof(double) * (size_t)out);
    n = out;
  }
  const double* W = p; p += (size_t)H->nOut * n;
  const double* b = p;
  int best = 0;
  double bestV = 0.0;
  for (int o = 0; o < H->nOut; ++o) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += W[o * n + k] * x[k];
    acc += b[o];
    if (H->outSigmoid) acc = 1.0 / (1.0 + exp(-acc));
    if (outs) outs[o] = acc;
    if (o == 0 || acc > bestV) { bestV = acc; best = o; }
  }
  return best;
}

/* ---------------------------------------------------------------- PU-list runner */

/* fill err[] of every record with the 3x3 integer error surface around its integer MV */
void orc_fill_surface(const orc_pel* org, int os, const orc_pel* const* refs, int rs, orc_pu* pus, int n, int fen) {
  for (int i = 0; i < n; ++i) {
    orc_pu* p = &pus[i];
    orc_int_surface(org + p->y * os + p->x, os, p->w, p->h,
                    refs[p->refSlot] + (p->y + p->mvIntY) * rs + p->x + p->mvIntX, rs, fen, p->err);
  }
}

/* TComPrediction::xPredInterBlk for luma with bi = false (TComPrediction.cpp:661-680): three cases. */
void orc_mc_luma(const orc_pel* ref, int rs, orc_pel* dst, int ds, int w, int h, int xFrac, int yFrac) {
  if (yFrac == 0) {
    orc_filter_hor(1, ref, rs, dst, ds, w, h, xFrac, 1, 8);
  } else if (xFrac == 0) {
    orc_filter_ver(1, ref, rs, dst, ds, w, h, yFrac, 1, 1, 8);
  } else {
    orc_pel tmp[(64 + 7) * 64];
    orc_filter_hor(1, ref - 3 * rs, rs, tmp, 64, w, h + 7, xFrac, 0, 8);
    orc_filter_ver(1, tmp + 3 * 64, 64, dst, ds, w, h, yFrac, 0, 1, 8);
  }
}

/* TComYuv::removeHighFreq, bClipToBitDepths = false (TComYuv.cpp:443-452) */
void orc_bi_pattern(const orc_pel* org, int os, const orc_pel* pred, int ps, orc_pel* dst, int ds, int w, int h) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) dst[y * ds + x] = (orc_pel)(2 * org[y * os + x] - pred[y * ps + x]);
}

void orc_run_pu_list(const orc_pel* org, int os, const orc_pel* const* refs, int rs, const orc_pu* pus, int n, int mode,
                     double lambda, int useHad, const void* nnBlob, orc_result* out) {
  for (int i = 0; i < n; ++i) {
    const orc_pu* p = &pus[i];
    orc_result r;
    memset(&r, 0, sizeof(r));
    if ((mode & 1) && (p->flags & ORC_PU_BI)) {
      /* xMotionEstimation with bBi: the pattern is 2*org - (other list's prediction) (TEncSearch.cpp:4462-4472) */
      const int oslot = (int)(p->err[0] & 0xff);
      const int omx = (int16_t)(p->err[1] & 0xffff), omy = (int16_t)(p->err[1] >> 16);
      orc_pel pred[64 * 64], pat[64 * 64];
      int16_t hxy[2], qxy[2];
      uint32_t c;
      orc_mc_luma(refs[oslot] + (p->y + (omy >> 2)) * rs + p->x + (omx >> 2), rs, pred, 64, p->w, p->h, omx & 3, omy & 3);
      orc_bi_pattern(org + p->y * os + p->x, os, pred, 64, pat, 64, p->w, p->h);
      orc_frac_dif(pat, 64, p->w, p->h, refs[p->refSlot] + p->y * rs + p->x, rs, p->mvIntX, p->mvIntY, p->mvPredX,
                   p->mvPredY, lambda, useHad, p->flags & 1 /* bIsLosslessCoded */, hxy, qxy, &c);
      r.halfX = (int8_t)hxy[0]; r.halfY = (int8_t)hxy[1];
      r.qterX = (int8_t)qxy[0]; r.qterY = (int8_t)qxy[1];
      r.cost = c;
    } else if (mode & 1) {
      int16_t hxy[2], qxy[2];
      uint32_t c;
      orc_frac_dif(org + p->y * os + p->x, os, p->w, p->h, refs[p->refSlot] + p->y * rs + p->x, rs, p->mvIntX,
                   p->mvIntY, p->mvPredX, p->mvPredY, lambda, useHad, p->flags & 1 /* bIsLosslessCoded */, hxy,
                   qxy, &c);
      r.halfX = (int8_t)hxy[0]; r.halfY = (int8_t)hxy[1];
      r.qterX = (int8_t)qxy[0]; r.qterY = (int8_t)qxy[1];
      r.cost = c;
    }
    if ((mode & 2) && nnBlob) {
      int16_t hxy[2], qxy[2];
      int cls = orc_nn_pred(nnBlob, p->err, p->h, p->w, NULL, hxy, qxy);
      r.nnHalfX = (int8_t)hxy[0]; r.nnHalfY = (int8_t)hxy[1];
      r.nnQterX = (int8_t)qxy[0]; r.nnQterY = (int8_t)qxy[1];
      r.nnClass = (uint8_t)cls;
    }
    out[i] = r;
  }
}
