// Shared declarations of the B200 fractional-ME engine (device + host side of libfme_b200.so).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fme_b200.h"

#define FME_NUM_PLANES 16     // P[fy][fx], fy,fx in 0..3; plane 0 = padded integer-pel copy
#define FME_COST_LUT_SIZE 160 // MV bit counts: 2 * (1 + 2*17) = 70 max for 16-bit components, padded
#define FME_MAX_CLASSES 64    // (w index) * 8 + (h index), w,h in {4,8,12,16,24,32,48,64}
// K2 bins PUs by (reference slot group, shape class): key = (slot & 7) * 64 + class.  Packs are scheduled slot-major so
// that the 16 planes of one reference picture (43 MB at 1080p) stay in the 126 MB L2 while every shape class of that
// slot is served, instead of all slots being swept once per shape class.
#define FME_K2_SLOT_GROUPS 8
#define FME_K2_KEYS (FME_K2_SLOT_GROUPS * FME_MAX_CLASSES)

// MV-bit cost table of one slice, cost[bits] (fme_set_slice).  It travels BY VALUE as a kernel argument of k2_refine:
// every launch carries the table of the slice it was submitted under, so changing lambda never has to wait for
// submits in flight (lowdelay_P changes lambda every frame, cfg/encoder_lowdelay_P_main.cfg:24-27).
struct FmeCostLut {
  uint32_t v[FME_COST_LUT_SIZE];
};

// Geometry of one padded plane set (all slots share it).
struct FmeGeom {
  int W, H;        // picture
  int M;           // margin
  int Wp, Hp;      // padded: W + 2M, H + 2M
  int pitch;       // bytes per padded row (multiple of 128)
  size_t planeBytes;  // Hp * pitch
  size_t slotBytes;   // 16 * planeBytes
  int orgPitch;    // bytes per row of the source picture
  int Wc, Hc, Mc, Wcp, Hcp, cPitch;  // chroma (4:2:0) padded geometry
  size_t cPlaneBytes;
  int numSlots;     // reference slots allocated (records with a larger refSlot are clamped)
};

// NN weight header as laid out in the FMNN blob (see nn_weights.py / oracle header).
struct FmeNnHeader {
  int32_t magic, version, nErr, nEmb, embRows, embDim, nHidden, hidden[4], nOut, outSigmoid, reserved[3];
};
#define FME_NN_MAGIC 0x4e4e4d46

__host__ __device__ inline int fme_dim_index(int v) {
  // {4,8,12,16,24,32,48,64} -> 0..7, anything else -> -1
  switch (v) {
    case 4: return 0;
    case 8: return 1;
    case 12: return 2;
    case 16: return 3;
    case 24: return 4;
    case 32: return 5;
    case 48: return 6;
    case 64: return 7;
    default: return -1;
  }
}
__host__ __device__ inline int fme_index_dim(int i) {
  // {4, 8, 12, 16, 24, 32, 48, 64} without a (local-memory) table
  i &= 7;
  return i < 4 ? 4 * (i + 1) : ((i & 1) ? 32 : 24) << ((i - 4) >> 1);
}

// The inter PU shapes HEVC can produce (2Nx2N, 2NxN, Nx2N and the AMP splits of CUs 8..64, TypeDef.h PartSize):
// one side is the CU size S in {8,16,32,64} and the other is S, S/2, S/4 or 3S/4 (quarter splits from S = 16 up;
// 4x4 does not exist).  Anything else (12x12, 24x8, ...) is rejected on the host and skipped by the kernels.
__host__ __device__ inline bool fme_hevc_pu_shape(int w, int h) {
  const int S = w > h ? w : h, m = w > h ? h : w;
  if (S != 8 && S != 16 && S != 32 && S != 64) return false;
  if (m == S || 2 * m == S) return true;
  return S >= 16 && (4 * m == S || 4 * m == 3 * S);
}

// ---- launchers implemented in the kernel translation units ----------------------------------
struct FmeK2Scratch {
  int* classCount;   // [FME_K2_KEYS]
  int* classCursor;  // [FME_K2_KEYS]
  int* classOffset;  // [FME_K2_KEYS + 1] in schedule order v = key ^ 63
  int* packOffset;   // [FME_K2_KEYS + 1] cumulative number of packs, schedule order
  int* order;        // [maxPUs] PU indices grouped by class
  short* keys;       // [maxPUs] binning key of every record (-1: not served by this pass), written by k2_count so that
                     // k2_scatter reads 2 bytes per record instead of striding through the 52-byte records twice more
  int* workCounter;  // [1] dynamic pack scheduler
};

// d_tileCounter: two zero-initialised ints owned by the ctx (dynamic tile hand-out; the kernel re-arms them itself)
// Grid of a grid-striding kernel: at most one resident wave (numSMs x the CTAs of this kernel an SM holds).  More CTAs
// than that run after the wave at a fraction of the occupancy (K3: 6 per SM against 5 resident cost 4 %).
template <typename Kernel>
inline int fme_one_wave(Kernel kernel, int threads, size_t smemBytes, int wanted) {
  int dev = 0, sms = 148, perSM = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, threads, smemBytes) != cudaSuccess || perSM < 1) perSM = 4;
  return wanted < sms * perSM ? wanted : sms * perSM;
}

cudaError_t fme_launch_k1(const FmeGeom& g, const uint8_t* d_pic, int picPitch, uint8_t* d_planes, int* d_tileCounter,
                          int numSMs, int rowBegin, int rowEnd, int path, cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_pad_chroma(const FmeGeom& g, const uint8_t* d_pic, int picPitch, uint8_t* d_plane,
                                  cudaStream_t s, int64_t* launches);
// K3's work folded into the K2 kernel (k2_refine.cu): the FMNN blob on the device (shipped 17-22-20-49 shape only).
struct FmeK2NnFuse {
  const float* d_blob;
  int fma;         // fme_config.nnFma
  float outClamp;  // 53 ln 2 for sigmoid-output nets, +inf otherwise (as in k3_nn_fixed)
};
cudaError_t fme_launch_k2(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_pu* d_pus, int n,
                          fme_result* d_res, const FmeCostLut& costLut, int useHad, int biPred, int k2Path,
                          const FmeK2Scratch& sc, int numSMs, cudaStream_t s, int64_t* launches,
                          const FmeK2NnFuse* nn = nullptr, bool* nnDone = nullptr);
cudaError_t fme_launch_k3(const fme_pu* d_pus, int n, fme_result* d_res, const float* d_nn, size_t nnBytes,
                          const FmeNnHeader& h, int fma, cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_k0(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, fme_pu* d_pus, int n,
                          int fen, cudaStream_t s, int64_t* launches, const int* d_runIf = nullptr);
cudaError_t fme_launch_clear_results(fme_result* d_res, int n, cudaStream_t s, int64_t* launches);

cudaError_t fme_launch_filter(int isVertical, int ntaps, int isFirst, int isLast, int bitDepth, const int16_t* d_src,
                              int srcStride, int16_t* d_dst, int dstStride, int w, int h, int frac, int isLuma,
                              cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_dist(int kind, const int16_t* d_org, int orgStride, const int16_t* d_cur, int curStride, int w,
                            int h, int bitDepth, int subShift, int nBlocks, uint32_t* d_out, cudaStream_t s,
                            int64_t* launches);
cudaError_t fme_launch_pel_to_u8(const int16_t* d_src, int srcStride, uint8_t* d_dst, int dstPitch, int w, int h,
                                 cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_pack_results(const fme_result* d_res, int n, fme_result8* d_out, cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_expand_heads(const fme_pu_head* d_heads, int n, fme_pu* d_pus, cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_expand_compact(const fme_pu_compact* d_recs, int n, fme_pu* d_pus, cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_apply_grids(const fme_err_grid* d_grids, int nGrids, fme_pu* d_pus, int n, cudaStream_t s,
                                   int64_t* launches);
cudaError_t fme_launch_mc_bi(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_cb, const uint8_t* d_cr,
                             const fme_mc_bi_pu* d_pus, int n, int16_t* d_y, int16_t* d_cbOut, int16_t* d_crOut,
                             cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_mc(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_cb, const uint8_t* d_cr,
                          const fme_mc_pu* d_pus, int n, int16_t* d_y, int16_t* d_cbOut, int16_t* d_crOut,
                          cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_cand_cost(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_cand_pu* d_cands,
                                 int n, const FmeCostLut& lut, int useHad, uint32_t* d_cost, int32_t* d_best, cudaStream_t s,
                                 int64_t* launches);
cudaError_t fme_launch_mc_luma_compact(const FmeGeom& g, const uint8_t* d_planes, const fme_mc_pu* d_pus, int n,
                                       const uint32_t* d_offsets, uint8_t* d_out, cudaStream_t s, int64_t* launches);
cudaError_t fme_launch_pred_error(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_mc_pu* d_pus,
                                  int n, int useHad, uint32_t* d_out, cudaStream_t s, int64_t* launches);
