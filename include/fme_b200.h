/* fme_b200.h -- C ABI of the B200-native fractional-pel motion-estimation engine.
 *
 * Drop-in boundary for the fractional-ME hot path of HM-16.9-NN_FME.
 * The reference has no FFI; its "API" for this path is a set of C++ member functions.  Each
 * entry point below names the reference interface it replaces (paths relative to
 * /root/reference/source/Lib).  A header-only C++ adaptor that reproduces the reference's own
 * signatures on top of these calls is in hm16.9-nn_fme_b200/adaptor/fme_hm_adaptor.h;
 * INTEGRATION.md shows the binding a maintainer would add to TEncSearch.cpp.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every call returns FME_OK (0) or a
 * negative fme_status; fme_last_error() gives the text of the last failure on the calling thread.
 * All host buffers are caller-owned.  One fme_ctx per encoder instance; calls on one ctx are
 * serialised by the caller (the reference is single-threaded, TEncSearch.cpp:55-77).
 * There is NO CPU fallback: if no sm_100 device is usable, fme_create fails.
 *
 * Sample types: Pel = int16_t (TLibCommon/TypeDef.h:228), Distortion = uint32_t (TypeDef.h:239),
 * TComMv components = int16_t (TLibCommon/TComMv.h:53-54).  Frame-level passes are 8-bit 4:2:0
 * (RExt__HIGH_BIT_DEPTH_SUPPORT = 0, TypeDef.h:118); block-level filters accept bitDepth 8..12.
 */
#ifndef FME_B200_H
#define FME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fme_ctx fme_ctx;

typedef enum fme_status {
  FME_OK = 0,
  FME_ERR_INVALID = -1,   /* bad argument */
  FME_ERR_CUDA = -2,      /* CUDA runtime/driver failure (text in fme_last_error) */
  FME_ERR_NO_DEVICE = -3, /* no usable sm_100 GPU: the engine has no CPU path */
  FME_ERR_STATE = -4,     /* call order violated (e.g. submit before upload_ref / set_slice) */
  FME_ERR_IO = -5         /* weight file could not be read / parsed */
} fme_status;

/* fme_submit modes (SURVEY.md section 7 "both run" quirk, TEncSearch.cpp:4534-4597) */
#define FME_MODE_STD 1  /* standard interpolate-and-search FME: half/qter MV + cost              */
#define FME_MODE_NN 2   /* NN_pred only                                                            */
#define FME_MODE_BOTH 3 /* master behaviour: standard cost AND NN MV                               */

/* fme_pu.flags */
#define FME_PU_LOSSLESS 0x01 /* bIsLosslessCoded: SAD instead of Hadamard (TEncSearch.cpp:5258, 1604) */
#define FME_PU_BI 0x04 /* bi-predictive refinement call (bBi, TEncSearch.cpp:4462-4472): the search pattern is
                          2*org - P_other (TComYuv::removeHighFreq, no clipping), P_other = the other list's
                          uni-prediction (xPredInterBlk, bi = false) from reference slot err[0] & 0xff at the
                          quarter-pel MV packed in err[1] (x = low, y = high int16, after clipMv).  err[] carries no
                          error grid for such records: the reference's NN_pred runs on stale globals there (its
                          integer search is xPatternSearch, which never refreshes array_e), so the NN fields of the
                          result are meaningless.  Needs fme_config.biPred = 1. */
#define FME_PU_ERR_ON_GPU 0x02 /* err[] is ignored; the 3x3 integer error surface is computed on the
                                  device by the K0 pass (TEncSearch.cpp:5037-5050 semantics)         */

typedef struct fme_config {
  int32_t device;       /* CUDA device ordinal */
  int32_t width;        /* luma picture width  (TComPicYuv) */
  int32_t height;       /* luma picture height */
  int32_t margin;       /* luma margin of the padded reference planes; reference uses 80
                           (TComPicYuv.cpp:94-95).  Must be a multiple of 16, >= 16. */
  int32_t bitDepth;     /* 8 */
  int32_t numRefSlots;  /* reference pictures resident at once (lowdelay_P: 4) */
  int32_t maxPUs;       /* capacity of one fme_submit batch */
  int32_t useHadME;     /* HadamardME cfg flag (TEncCfg getUseHADME, TEncSearch.cpp:1604) */
  int32_t fen;          /* FEN / FastInterSearchMode 1|3: row-subsampled SAD in integer ME
                           (TEncSearch.cpp:1158-1164); only used by the K0 error-surface pass */
  int32_t nnFma;        /* 0 (default): NN_pred with the reference's mul-then-add rounding, bit-identical to the
                           oracle.  1: fused multiply-add in the dense layers of the compile-time fast paths -- half the
                           FP32 operations (measured: K3 0.153 -> 0.125 ms per 858 000 PUs), within BASELINE's NN
                           tolerance (class agreement >= 99.9 %), not bit-exact. */
  int32_t biPred;       /* 1: serve FME_PU_BI records with a second binning + refinement pass per submit (random
                           access configurations); 0 (default): such records are rejected / left untouched */
  int32_t k2Path;       /* how K2 computes the 8x8 SATD of uni-prediction Hadamard PUs; all paths are bit-identical:
                           FME_K2_PATH_AUTO (0, the fastest measured), _SWAR, _MMA_PACK, _MMA_GROUP, _UMMA */
  int32_t k1Path;       /* how K1 builds the 16 sub-pel planes; both paths are bit-identical: FME_K1_PATH_AUTO (0, the
                           fastest measured), _DP4A (integer dot products on the CUDA cores), _MMA (both filter stages
                           as exact mma.sync products against Toeplitz tap matrices), _UMMA (experimental: the vertical
                           stage on tcgen05.mma with the accumulator in TMEM; never picked by AUTO) */
  int32_t k3Fuse;       /* 0 (default): K3 is its own launch behind K2.  1 (experimental): when one submit runs both passes
                           over the same records with the shipped 17-22-20-49 net, K3's work items ride inside the persistent
                           K2 kernel as extra work items (same arithmetic, same results).  Measured slower: the two code
                           bodies do not fit the 32 KB instruction cache of an SM together */
  int32_t reserved[2];
} fme_config;
#define FME_K1_PATH_AUTO 0
#define FME_K1_PATH_DP4A 1
#define FME_K1_PATH_MMA 2
#define FME_K1_PATH_UMMA 3 /* experimental: vertical stage on tcgen05.mma / TMEM */
#define FME_K2_PATH_AUTO 0
#define FME_K2_PATH_SWAR 1      /* carry-tolerant 16-bit SWAR Hadamard in registers */
#define FME_K2_PATH_MMA_PACK 2  /* fp16-in / fp32-accumulate mma.sync Hadamard inside the 32-lane packs */
#define FME_K2_PATH_MMA_GROUP 3 /* the same on eight-tile groups x four candidates per evaluation */
#define FME_K2_PATH_UMMA 4      /* tcgen05.mma kind::i8 (u8 pixels x +-1 Hadamard matrix -> s32 in TMEM), one row per tile-candidate */

/* One PU's integer-ME hand-off: what xMotionEstimation has in hand at
 * TEncSearch.cpp:4534 / 4541.  err[] is the raster 3x3 integer error grid [TL,T,TR,L,C,R,BL,B,BR]:
 * array_e[0..3], C, array_e[4..7] (TEncSearch.cpp:88, 1341-1376, 5049-5050). */
typedef struct fme_pu {
  int16_t x, y;             /* PU top-left luma sample in the picture */
  uint8_t w, h;             /* PU size (iRoiWidth, iRoiHeight, TEncSearch.cpp:4460) */
  uint8_t refSlot;          /* which uploaded reference picture */
  uint8_t flags;            /* FME_PU_* */
  int16_t mvIntX, mvIntY;   /* best integer MV, full-pel units (rcMv at TEncSearch.cpp:4534) */
  int16_t mvPredX, mvPredY; /* predictor in quarter-pel units (m_mvPredictor, TComRdCost.h:163) */
  uint32_t err[9];
} fme_pu; /* 52 bytes */

/* The same hand-off without the error grid: the engine computes err[] itself with the integer-ME metric (K0 pass,
 * SURVEY.md "next" row f1), which removes 36 of the 52 bytes per PU from the host->device traffic. */
typedef struct fme_pu_head {
  int16_t x, y;
  uint8_t w, h;
  uint8_t refSlot;
  uint8_t flags;
  int16_t mvIntX, mvIntY;
  int16_t mvPredX, mvPredY;
} fme_pu_head; /* 16 bytes = the first 16 bytes of fme_pu */

typedef struct fme_result {
  int8_t halfX, halfY, qterX, qterY;         /* rcMvHalf, rcMvQter in {-1,0,1} (TEncSearch.h:423-432) */
  uint32_t cost;                             /* ruiCost: best quarter-stage cost incl. MV bits         */
  int8_t nnHalfX, nnHalfY, nnQterX, nnQterY; /* MVX_HALF.. (TEncSearch.cpp:55, 136-193)                */
  uint8_t nnClass;                           /* NN_out, 0..48                                          */
  uint8_t pad[3];
} fme_result; /* 16 bytes */

/* Compact result, 8 bytes instead of 16: OR FME_MODE_RESULT8 into the mode of any fme_submit* call and pass an
 * fme_result8 array as `out`.  Halves the device->host traffic of a batch (the end-to-end path is PCIe / host-fabric
 * bound when several GPUs share one host).  Fields hold value + 1 in two bits each; fme_result8_unpack() expands. */
#define FME_MODE_RESULT8 0x10
typedef struct fme_result8 {
  uint32_t cost; /* ruiCost */
  uint32_t mv;   /* bits 0-1 halfX+1, 2-3 halfY+1, 4-5 qterX+1, 6-7 qterY+1, 8-9 nnHalfX+1, 10-11 nnHalfY+1,
                    12-13 nnQterX+1, 14-15 nnQterY+1, 16-21 nnClass */
} fme_result8; /* 8 bytes */
static inline void fme_result8_unpack(const fme_result8* p, fme_result* r) {
  const uint32_t m = p->mv;
  r->halfX = (int8_t)((m & 3) - 1); r->halfY = (int8_t)(((m >> 2) & 3) - 1);
  r->qterX = (int8_t)(((m >> 4) & 3) - 1); r->qterY = (int8_t)(((m >> 6) & 3) - 1);
  r->cost = p->cost;
  r->nnHalfX = (int8_t)(((m >> 8) & 3) - 1); r->nnHalfY = (int8_t)(((m >> 10) & 3) - 1);
  r->nnQterX = (int8_t)(((m >> 12) & 3) - 1); r->nnQterY = (int8_t)(((m >> 14) & 3) - 1);
  r->nnClass = (uint8_t)((m >> 16) & 63);
  r->pad[0] = r->pad[1] = r->pad[2] = 0;
}

/* ---- lifetime -------------------------------------------------------------------------- */
int fme_create(const fme_config* cfg, fme_ctx** out);
void fme_destroy(fme_ctx* ctx);
const char* fme_last_error(void);
const char* fme_version(void);
/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work of this ctx; NULL
 * restores the ctx's own stream.  Lets a host framework time the kernels with its own events. */
int fme_set_stream(fme_ctx* ctx, void* cudaStream);
int fme_synchronize(fme_ctx* ctx);

/* ---- per-sequence / per-slice state ------------------------------------------------------ */
/* NN weights: replaces the Eigen comma initialisers of TEncSearch::init (TEncSearch.cpp:470-1073).
 * blob = "FMNN" container (see hm16.9-nn_fme_b200/nn_weights.py); csv dir = DL/blowing/<qp>. */
int fme_set_nn_weights(fme_ctx* ctx, const void* blob, size_t bytes);
int fme_load_nn_csv_dir(fme_ctx* ctx, const char* dir);
/* Motion lambda: replaces TComRdCost::setLambda + selectMotionLambda(true,0,false)
 * (TComRdCost.cpp:104-117, TComRdCost.h:159).  Builds cost[bits] = uint32((65536*sqrt(lambda)*bits)/65536.0)
 * on the host with the reference's exact double expression (TComRdCost.h:165-169). */
int fme_set_slice(fme_ctx* ctx, double lambda);
/* fme_set_slice never waits for the device: the table travels with every K2 launch as a kernel argument, so submits
 * already issued keep the lambda they were submitted under and the call may be made once per slice between
 * fme_submit_async calls without draining the pipeline.  FME_PU_LOSSLESS records use the same table, which is what
 * the reference does in the standard lossy and all-lossless cost modes (COST_MIXED_LOSSLESS_LOSSY_CODING, which
 * selects a second lambda for transquant-bypass PUs, TComRdCost.cpp:110-116, is not reproduced). */

/* ---- frame data --------------------------------------------------------------------------- */
/* Reference picture: y points at picture sample (0,0) of a TComPicYuv luma plane (Pel, stride in
 * samples).  Only the width x height picture area is read; the margin is regenerated on the device
 * by edge replication, which is what TComPicYuv::extendPicBorder (TComPicYuv.cpp:229-276) stores.
 * Runs the K1 pass: all 15 sub-pel planes of SURVEY.md A.1 (replaces xExtDIFUpSamplingH/Q,
 * TEncSearch.cpp:6331-6532, and their filterHor/filterVer calls). */
int fme_upload_ref(fme_ctx* ctx, int slot, const int16_t* y, int stride);
int fme_upload_ref_u8(fme_ctx* ctx, int slot, const uint8_t* y, int stride);
/* Source picture (pcYuvOrg / TComPattern key, TEncSearch.cpp:4474-4479). */
int fme_upload_org(fme_ctx* ctx, const int16_t* y, int stride);
int fme_upload_org_u8(fme_ctx* ctx, const uint8_t* y, int stride);

/* ---- the batched search -------------------------------------------------------------------- */
/* Replaces, for n PUs at once, xPatternSearchFracDIF (TEncSearch.cpp:5232-5269, mode bit 0) and
 * NN_pred (TEncSearch.cpp:85-204, mode bit 1).  Synchronous: host records in, host results out. */
int fme_submit(fme_ctx* ctx, const fme_pu* pus, int n, fme_result* out, int mode);
/* Asynchronous halves for pipelining: records/results in pinned host memory owned by the caller.
 * fme_submit_async enqueues H2D + kernels + D2H on the ctx stream; fme_synchronize completes it. */
int fme_submit_async(fme_ctx* ctx, const fme_pu* pus, int n, fme_result* out, int mode);
/* Contract of the asynchronous and device-resident entry points (fme_submit_async, fme_submit_heads_async,
 * fme_submit_device): the records are NOT validated on the host (they are still being DMA'd).  On the device
 *   - records whose shape HEVC cannot produce, and FME_PU_BI records on a ctx without biPred, are served by no pass:
 *     their result is the sentinel {half = qter = (0,0), cost = 0xffffffff} (the NN fields are still written);
 *   - reference slots beyond numRefSlots and integer MVs beyond the padded plane are clamped (memory safety only);
 *   - the K0 pass runs whenever mode includes FME_MODE_NN and fills err[] of exactly the records flagged
 *     FME_PU_ERR_ON_GPU, so flagged and unflagged records may be mixed freely in one batch;
 *   - FME_PU_BI is ignored on head records (a head has no err[] to name the other list's prediction). */
/* fme_submit / fme_submit_async for records without the error grid: equivalent to full records carrying
 * FME_PU_ERR_ON_GPU (the 3x3 surface of TEncSearch.cpp:5037-5050 is computed on the device before K2 / K3). */
int fme_submit_heads(fme_ctx* ctx, const fme_pu_head* heads, int n, fme_result* out, int mode);
int fme_submit_heads_async(fme_ctx* ctx, const fme_pu_head* heads, int n, fme_result* out, int mode);
/* Head records plus the caller's own error grids for SOME of them.  The integer search has array_e / C in hand for
 * every PU (TEncSearch.cpp:88, 1341-1376, 5049-5050); sending them costs 40 bytes per PU over the bus, computing them on
 * the device costs K0 time that is dominated by per-PU overhead, not by PU area.  The caller picks the split that its
 * bus allows (bench.py: grids for the PUs of 128 samples and more on one or two GPUs -- K0 0.173 -> 0.088 ms per 858 000
 * PUs for +8.4 MB per frame -- none on eight GPUs sharing one host fabric).  grids[j].pu indexes heads[]; a head named by
 * a grid is served exactly like a full record carrying that err[]; entries with pu outside [0, n) are ignored on the
 * asynchronous path and rejected on the synchronous one. */
typedef struct fme_err_grid {
  int32_t pu;      /* index into heads[] */
  uint32_t err[9]; /* raster 3x3 grid as in fme_pu.err */
} fme_err_grid; /* 40 bytes */
int fme_submit_heads_grids(fme_ctx* ctx, const fme_pu_head* heads, int n, const fme_err_grid* grids, int nGrids,
                           fme_result* out, int mode);
int fme_submit_heads_grids_async(fme_ctx* ctx, const fme_pu_head* heads, int n, const fme_err_grid* grids, int nGrids,
                                 fme_result* out, int mode);
/* Compact records: the head plus the caller's 3x3 error grid as nine 24-bit little-endian values, 44 bytes instead of 52.
 * The integer metric of this fork is the SSE (TComRdCost.cpp:212), so a grid value is below w*h*255^2: every PU of at most
 * 256 luma samples fits 24 bits exactly (256 * 65025 < 2^24) -- 94 % of the records of a 1080p list -- and so does almost
 * every larger one in practice; a PU with a value of 2^24 or more gets its grid through the fme_err_grid list instead
 * (big[j].pu indexes recs[]; its err24 bytes are then ignored).  No K0 pass runs: served exactly like full records
 * carrying the same err[].  One GPU moves full records at the PCIe limit (52.9 MB per 1080p frame at ~51 GB/s = 1.03 ms
 * against a 0.93 ms device step); 44-byte records put the end-to-end path back on the device step.  FME_PU_BI and
 * FME_PU_ERR_ON_GPU are ignored on compact records (a bi-predictive record needs the full err[] words). */
typedef struct fme_pu_compact {
  fme_pu_head head;
  uint8_t err24[27]; /* err[k] = err24[3k] | err24[3k+1] << 8 | err24[3k+2] << 16, raster order as fme_pu.err */
  uint8_t reserved;
} fme_pu_compact; /* 44 bytes */
/* Packs one full record; returns 1 when a grid value needs more than 24 bits (send that PU's grid in the big list). */
static inline int fme_pu_compact_pack(const fme_pu* p, fme_pu_compact* c) {
  int k, big = 0;
  const uint8_t* src = (const uint8_t*)p;
  uint8_t* dst = (uint8_t*)&c->head;
  for (k = 0; k < 16; ++k) dst[k] = src[k];
  for (k = 0; k < 9; ++k) {
    big |= p->err[k] >> 24 ? 1 : 0;
    c->err24[3 * k] = (uint8_t)p->err[k];
    c->err24[3 * k + 1] = (uint8_t)(p->err[k] >> 8);
    c->err24[3 * k + 2] = (uint8_t)(p->err[k] >> 16);
  }
  c->reserved = 0;
  return big;
}
int fme_submit_compact(fme_ctx* ctx, const fme_pu_compact* recs, int n, const fme_err_grid* big, int nBig, fme_result* out,
                       int mode);
int fme_submit_compact_async(fme_ctx* ctx, const fme_pu_compact* recs, int n, const fme_err_grid* big, int nBig,
                             fme_result* out, int mode);
/* With the ctx's own streams the copies of fme_upload_* / fme_submit_async run on dedicated copy streams and
 * overlap the kernels of neighbouring frames (staging rings of three, at most three submits in flight).
 * fme_wait_oldest blocks until the results of the oldest outstanding fme_submit_async are in `out`. */
int fme_wait_oldest(fme_ctx* ctx);
/* Device-resident variant: d_pus / d_out are device pointers; nothing is copied. */
int fme_submit_device(fme_ctx* ctx, const fme_pu* d_pus, int n, fme_result* d_out, int mode);
/* Re-run K1 on the reference picture already resident in `slot` (device-resident benchmarking): the picture is taken
 * from the slot's own integer-pel plane.  FME_ERR_STATE when the slot holds no picture. */
int fme_interp_slot(fme_ctx* ctx, int slot);
/* Reference picture already on the device (d_y: device pointer to 8-bit samples, pitch in bytes), then K1.  A 4-byte
 * aligned picture is read in place -- it must stay unchanged until the work issued by this call has run on the ctx
 * stream --, anything else is staged through a device copy.  Used with torch/NCCL buffers in the banded mode. */
int fme_upload_ref_device_u8(fme_ctx* ctx, int slot, const uint8_t* d_y, int pitch);
int fme_upload_org_device_u8(fme_ctx* ctx, const uint8_t* d_y, int pitch);
/* The same for a ROW RANGE of the picture (banded multi-GPU mode): only the sub-pel plane rows [picRowBegin, picRowEnd)
 * (picture rows; the range may reach into the margin) are produced, whole 16-row tiles covering it; the rest of the slot
 * keeps its previous content.  Exact as long as the range covers every row the submitted PUs reference: a PU at row y with
 * integer MV my and height h reads plane rows y + my - 1 .. y + my + h.  d_y must be 4-byte aligned, read in place. */
int fme_upload_ref_device_u8_rows(fme_ctx* ctx, int slot, const uint8_t* d_y, int pitch, int picRowBegin, int picRowEnd);
/* Source picture rows [rowBegin, rowEnd) only (clamped to the picture): in the banded mode a rank's PUs read nothing but
 * their own rows of the source (K2 and K0: rows y .. y + h - 1 of a PU at row y), so copying the whole picture on every
 * rank is work that does not shrink with the number of ranks.  The other rows of the ctx's source buffer are undefined. */
int fme_upload_org_device_u8_rows(fme_ctx* ctx, const uint8_t* d_y, int pitch, int rowBegin, int rowEnd);
/* K0: fill fme_pu.err[] of device-resident records from the 3x3 integer error surface
 * (xTZ8PointSquareSearch(save=true) metric, TEncSearch.cpp:1085-1166, 5037-5050). */
int fme_int_surface_device(fme_ctx* ctx, fme_pu* d_pus, int n);

/* ---- block-level parity entry points ------------------------------------------------------- */
/* TComInterpolationFilter::filterHor / filterVer (TComInterpolationFilter.cpp:341-394).
 * comp: 0 = luma (8-tap, frac 0..3), 1/2 = chroma 4:2:0 (4-tap, frac 0..7).  src must be readable
 * over the taps' support exactly as in the reference (3 left/4 right, 1/2 for chroma). */
int fme_filter_hor(fme_ctx* ctx, int comp, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w,
                   int h, int frac, int isLast, int bitDepth);
int fme_filter_ver(fme_ctx* ctx, int comp, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w,
                   int h, int frac, int isFirst, int isLast, int bitDepth);
/* DistParam::DistFunc (TComRdCost.h:60-108; TComRdCost.cpp:359-1495) for nBlocks independent block
 * pairs of one shape.  kind: 0 = integer-ME metric (SSE, or SAD12/24/48 with subShift), 1 = HADs,
 * 2 = SADs.  org/cur hold nBlocks blocks back to back, each h rows of `stride` samples. */
int fme_dist(fme_ctx* ctx, int kind, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w,
             int h, int bitDepth, int subShift, int nBlocks, uint32_t* out);
/* getCostOfVectorWithPredictor (TComRdCost.h:165-174) with the current slice lambda. */
int fme_mv_cost(fme_ctx* ctx, int x, int y, int scale, int predX, int predY, uint32_t* out);

/* ---- motion compensation ("next" row f2) ------------------------------------ */
/* Chroma reference planes for MC (Pel, picture sample (0,0), chroma stride; 4:2:0). */
int fme_upload_ref_chroma(fme_ctx* ctx, int slot, const int16_t* cb, const int16_t* cr, int stride);
int fme_upload_ref_chroma_u8(fme_ctx* ctx, int slot, const uint8_t* cb, const uint8_t* cr, int stride);
/* Raw planar 8-bit 4:2:0 frames as TVideoIOYuv reads and writes them (TLibVideoIO/TVideoIOYuv.cpp; Y, Cb, Cr back to
 * back, no padding; SURVEY.md "next" row f4): the frame buffer goes straight into the padded device planes (luma: K1 on
 * all 16 planes; chroma, if withChroma: the two padded planes fme_mc reads).  fme_upload_org_yuv420_u8 takes the luma
 * plane of a source frame. */
int fme_upload_ref_yuv420_u8(fme_ctx* ctx, int slot, const uint8_t* frame, int withChroma);
int fme_upload_org_yuv420_u8(fme_ctx* ctx, const uint8_t* frame);
/* xPredInterBlk, uni-prediction (TComPrediction.cpp:643-681) for n PUs: mv = quarter-pel luma MV.
 * dstY: n blocks back to back, each 64x64 Pel (stride 64); dstCb/dstCr: each 32x32 (stride 32). */
typedef struct fme_mc_pu {
  int16_t x, y;
  uint8_t w, h, refSlot, flags;
  int16_t mvX, mvY;
} fme_mc_pu; /* 12 bytes */
int fme_mc(fme_ctx* ctx, const fme_mc_pu* pus, int n, int16_t* dstY, int16_t* dstCb, int16_t* dstCr);

/* xPredInterBi with both lists valid (TComPrediction.cpp:575-621): two xPredInterUni(bi = true) blocks of 14-bit
 * intermediates (xPredInterBlk with isLast = false, TComPrediction.cpp:661-680) averaged by TComYuv::addAvg
 * (TComYuv.cpp:354-409).  Default weights only (weighted prediction is off in the reference configs).
 * Output layout as fme_mc.  Chroma needs fme_upload_ref_chroma for both slots. */
typedef struct fme_mc_bi_pu {
  int16_t x, y;
  uint8_t w, h, refSlot0, refSlot1;
  int16_t mv0X, mv0Y, mv1X, mv1Y; /* quarter-pel luma MVs of list 0 / list 1 */
} fme_mc_bi_pu; /* 16 bytes */
int fme_mc_bi(fme_ctx* ctx, const fme_mc_bi_pu* pus, int n, int16_t* dstY, int16_t* dstCb, int16_t* dstCr);

/* Prediction error of n uni-predicted PUs at quarter-pel MVs: luma MC + HADs (SAD when HadamardME is off or the PU is
 * lossless) against the source block -- TEncSearch::xGetInterPredictionError (TEncSearch.cpp:3576-3596, merge
 * estimation) and the distortion of xGetTemplateCost (TEncSearch.cpp:4397-4436; there the ctx is created with
 * useHadME = 0).  SURVEY.md "next" row f3. */
int fme_pred_error(fme_ctx* ctx, const fme_mc_pu* pus, int n, uint32_t* out);

/* ---- batched candidate costs and compact prediction output (SURVEY.md "next" rows f2 / f3) -------- */
/* One candidate of AMVP template matching or merge estimation: a uni-predicted PU at a quarter-pel MV (already clipped
 * by the caller as TComDataCU::clipMv does) plus the side-information bits the reference adds:
 *   xGetTemplateCost (TEncSearch.cpp:4397-4436): SAD of the MC block against the source + the cost of
 *     m_auiMVPIdxCost[iMVPIdx][iMVPNum] bits (xGetMvpIdxBits, :4258-4284) -> flags |= FME_CAND_SAD;
 *   xMergeEstimation (TEncSearch.cpp:3599-3655): xGetInterPredictionError (HADs when HadamardME is on, :3576-3596) +
 *     getCost(uiMergeCand + 1, one less for the last candidate), uni-predictive candidates.
 * cost = distortion + uint32((m_dLambdaMotionSAD[0] * bits) / 65536.0): TComRdCost::calcRdCost(bits, dist, DF_SAD) in
 * the standard lossy mode (TComRdCost.cpp:80-89) and getCost (TComRdCost.h:165) are the same expression, the table
 * of fme_set_slice.  The candidates of one PU are consecutive; groupStart = 1 on the first of them.  bestIndex[i]
 * (optional), at every group start i, is the index of the group's first minimum -- the candidate the reference's
 * strict-< loops keep (:3646, :4343-4349). */
#define FME_CAND_SAD 0x08 /* SAD distortion whatever HadamardME says (template matching) */
typedef struct fme_cand_pu {
  int16_t x, y;
  uint8_t w, h, refSlot, flags; /* FME_CAND_SAD, FME_PU_LOSSLESS (-> SAD) */
  int16_t mvX, mvY;             /* quarter-pel luma MV */
  uint16_t bits;                /* MVP-index / merge-index bits of this candidate */
  uint16_t groupStart;          /* 1: first candidate of a PU */
} fme_cand_pu; /* 16 bytes */
int fme_cand_cost(fme_ctx* ctx, const fme_cand_pu* cands, int n, uint32_t* cost, int32_t* bestIndex);
/* The same on device-resident arrays, stream-ordered on the ctx stream, no synchronisation. */
int fme_cand_cost_device(fme_ctx* ctx, const fme_cand_pu* d_cands, int n, uint32_t* d_cost, int32_t* d_bestIndex);
/* Luma motion compensation with compact output: block i is written as w*h 8-bit samples (row-major, pitch w) at byte
 * offset offsets[i] of out (offsets are multiples of 4; normally the running sum of w*h).  The uni-prediction of
 * xPredInterBlk (TComPrediction.cpp:643-681) at 8 bit IS an 8-bit block: nothing is lost against fme_mc, which pads
 * every block to 64x64 Pel.  Host (synchronous) and device-resident (stream-ordered) forms. */
int fme_mc_luma_compact(fme_ctx* ctx, const fme_mc_pu* pus, int n, const uint32_t* offsets, uint8_t* out, size_t outBytes);
int fme_mc_luma_compact_device(fme_ctx* ctx, const fme_mc_pu* d_pus, int n, const uint32_t* d_offsets, uint8_t* d_out);

/* ---- introspection (parity tests, profiling) ------------------------------------------------ */
/* Copy padded sub-pel plane P[fy][fx] of `slot` to host: (height+2*margin) rows of (width+2*margin) bytes. */
int fme_download_plane(fme_ctx* ctx, int slot, int fy, int fx, uint8_t* dst, int dstStride);
/* Milliseconds spent in the kernels of the last fme_submit* / fme_upload_ref* / fme_interp_slot call
 * (CUDA events on the ctx stream): k[0]=K1 interp, k[1]=K2 refine, k[2]=K3 NN, k[3]=K0 surface. */
int fme_last_kernel_ms(fme_ctx* ctx, float k[4]);
/* Number of kernel launches issued by this ctx since creation. */
int64_t fme_launch_count(fme_ctx* ctx);
/* Enable/disable per-kernel event timing (adds event records; default off). */
int fme_set_profiling(fme_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif
