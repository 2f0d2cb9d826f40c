/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the HM-16.9-NN_FME fractional-ME hot path.
 *
 * Nothing under oracle/ is part of the shipped product.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only as the
 * checker.  The product (hm16.9-nn_fme_b200/libfme_b200.so) never links, loads or calls it.
 *
 * Parity status ("pinned" = checked against outputs of the reference's own compiled code,
 * oracle/_ref/libhmref.so, by tests/test_oracle_vs_reference.py here and against the committed
 * fixtures in tests/golden/ everywhere):
 *   filters, SAD/SSE/HADs, MV-bit cost, xPatternSearchFracDIF ........ pinned, bit-exact
 *   NN_pred ............ pinned against the reference code built over oracle/eigen_standin
 *                        (bit-exact); PARITY UNPINNED against real Eigen 3.3.7 (absent here;
 *                        contract is 1e-5 relative, BASELINE.json north_star).
 */
#ifndef FME_ORACLE_H
#define FME_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int16_t orc_pel; /* Pel = Short, TypeDef.h:228 */

/* TComInterpolationFilter::filterHor / filterVer (TComInterpolationFilter.cpp:341-394); 4:2:0 only */
void orc_filter_hor(int isLuma, const orc_pel* src, int srcStride, orc_pel* dst, int dstStride, int w, int h,
                    int frac, int isLast, int bitDepth);
void orc_filter_ver(int isLuma, const orc_pel* src, int srcStride, orc_pel* dst, int dstStride, int w, int h,
                    int frac, int isFirst, int isLast, int bitDepth);

/* TComYuv::addAvg for one component (TComYuv.cpp:354-409): bi-prediction average of two 14-bit intermediate blocks,
 * dst = ClipBD((s0 + s1 + (1 << (shift-1)) + 2*IF_INTERNAL_OFFS) >> shift), shift = max(2, 14 - bitDepth) + 1. */
void orc_add_avg(const orc_pel* src0, int stride0, const orc_pel* src1, int stride1, orc_pel* dst, int dstStride, int w,
                 int h, int bitDepth);

/* TComRdCost distortion functions (TComRdCost.cpp:359-855, 861-1206, 1212-1495) */
uint32_t orc_sad(const orc_pel* org, int orgStride, const orc_pel* cur, int curStride, int w, int h, int bitDepth,
                 int subShift);
uint32_t orc_sse(const orc_pel* org, int orgStride, const orc_pel* cur, int curStride, int w, int h, int bitDepth);
uint32_t orc_hads(const orc_pel* org, int orgStride, const orc_pel* cur, int curStride, int w, int h, int bitDepth);
/* the integer-ME metric of this fork: SSE for w in {4,8,16,32,64}, SAD12/24/48 otherwise (TComRdCost.cpp:200-229) */
uint32_t orc_int_me_dist(const orc_pel* org, int orgStride, const orc_pel* cur, int curStride, int w, int h,
                         int bitDepth, int subShift);
/* kind 0 = integer-ME metric, 1 = HADs, 2 = SADs (same numbering as hmref_dist) */
uint32_t orc_dist(int kind, const orc_pel* org, int orgStride, const orc_pel* cur, int curStride, int w, int h,
                  int bitDepth, int subShift);

/* MV-bit cost (TComRdCost.cpp:104-117, 172-185; TComRdCost.h:159-174) */
uint32_t orc_exp_golomb_bits(int v);
double orc_motion_lambda(double lambda); /* 65536*sqrt(lambda) */
uint32_t orc_mv_cost(double motionLambda, int x, int y, int scale, int predX, int predY);
/* slice lambda, TEncSlice.cpp:290-325 (non-hierarchical depth 0 unless depth>0 is given) */
double orc_slice_lambda(int qp, double qpFactor, int depth, int hadME);

/* TEncSearch::xPatternSearchFracDIF (TEncSearch.cpp:5232-5269) for one PU.
 * ref points at the PU's collocated sample of the padded reference plane. */
void orc_frac_dif(const orc_pel* org, int orgStride, int w, int h, const orc_pel* ref, int refStride, int mvIntX,
                  int mvIntY, int predX, int predY, double lambda, int useHad, int lossless, int16_t halfXY[2],
                  int16_t qterXY[2], uint32_t* cost);
/* m_filteredBlock[v][h] as left by the last orc_frac_dif call (stride 80) */
const orc_pel* orc_filtered_block(int v, int h);

/* Sub-pel plane P[fy][fx] of SURVEY A.1 computed with the block filters over a whole padded plane:
 * out(x,y) for x in [x0,x0+w), y in [y0,y0+h) relative to picture sample (0,0) at `ref`. */
void orc_subpel_plane(const orc_pel* ref, int refStride, int x0, int y0, int w, int h, int fy, int fx, orc_pel* out,
                      int outStride);

/* 3x3 integer error surface (TEncSearch.cpp:1085-1166, 1324-1377, 5037-5050), raster order */
void orc_int_surface(const orc_pel* org, int orgStride, int w, int h, const orc_pel* refAtMv, int refStride, int fen,
                     uint32_t err9[9]);

/* ---- NN_pred (TEncSearch.cpp:85-204) ---- */
#define ORC_NN_MAGIC 0x4e4e4d46 /* "FMNN" */
#define ORC_NN_MAX_HIDDEN 4
typedef struct {
  int32_t magic, version;
  int32_t nErr;     /* 9 */
  int32_t nEmb;     /* 0 or 2 embedding tables (height, width) */
  int32_t embRows;  /* 8 */
  int32_t embDim;   /* 4 */
  int32_t nHidden;  /* 2 (master), 3 (blowing40-style), up to 4 */
  int32_t hidden[ORC_NN_MAX_HIDDEN];
  int32_t nOut;     /* 49 */
  int32_t outSigmoid;
  int32_t reserved[3];
} orc_nn_header; /* 16 int32, followed by float32 payload:
   mean[nErr] stdev[nErr] gammaIn[nErr] emb[nEmb][embRows][embDim]
   per hidden layer: W[out][in] b[out] gamma[out] beta[out];  output: W[nOut][in] b[nOut] */

size_t orc_nn_blob_floats(const orc_nn_header* h);
/* returns class index; logits may be NULL */
int orc_nn_pred(const void* blob, const uint32_t err9[9], int puHeight, int puWidth, float* logits,
                int16_t halfXY[2], int16_t qterXY[2]);
void orc_nn_class_to_mv(int cls, int16_t halfXY[2], int16_t qterXY[2]);
/* the reference's 3-layer backup network in double precision (Backups/4...cpp:4408-4490); payload = the FMNN payload
 * sequence as doubles.  Returns the class (first maximum of the sigmoid outputs); outs may be NULL. */
int orc_nn_pred_f64(const void* header, const double* payload, const uint32_t err9[9], double* outs);

/* ---- PU-list runner (same record layout as include/fme_b200.h) ---- */
typedef struct {
  int16_t x, y;
  uint8_t w, h, refSlot, flags;
  int16_t mvIntX, mvIntY;
  int16_t mvPredX, mvPredY;
  uint32_t err[9];
} orc_pu;
typedef struct {
  int8_t halfX, halfY, qterX, qterY;
  uint32_t cost;
  int8_t nnHalfX, nnHalfY, nnQterX, nnQterY;
  uint8_t nnClass, pad[3];
} orc_result;
/* Bi-predictive refinement (bBi, TEncSearch.cpp:4462-4472): flags & ORC_PU_BI marks a record whose search pattern is
 * 2*org - P_other instead of org, P_other being the other list's uni-prediction (xPredInterBlk, bi = false) from
 * reference slot err[0] & 0xff at the quarter-pel MV packed in err[1] (x = low, y = high int16).  err[] carries no
 * error grid for such records: the reference's NN_pred runs on stale globals there (integer search is xPatternSearch,
 * which never refreshes array_e). */
#define ORC_PU_BI 0x04
/* xPredInterBlk, luma, uni-prediction (TComPrediction.cpp:661-680): ref -> the sample the integer part of mv points at */
void orc_mc_luma(const orc_pel* ref, int refStride, orc_pel* dst, int dstStride, int w, int h, int xFrac, int yFrac);
/* TComYuv::removeHighFreq without clipping (TComYuv.cpp:411-455; ClipForBiPredMEEnabled = 0): dst = 2*org - pred */
void orc_bi_pattern(const orc_pel* org, int orgStride, const orc_pel* pred, int predStride, orc_pel* dst, int dstStride,
                    int w, int h);
void orc_fill_surface(const orc_pel* org, int orgStride, const orc_pel* const* refs, int refStride, orc_pu* pus, int n,
                      int fen);
/* mode bit0 = standard FME, bit1 = NN_pred.  refs[s] -> picture sample (0,0) of padded plane s. */
void orc_run_pu_list(const orc_pel* org, int orgStride, const orc_pel* const* refs, int refStride, const orc_pu* pus,
                     int n, int mode, double lambda, int useHad, const void* nnBlob, orc_result* out);

#ifdef __cplusplus
}
#endif
#endif
