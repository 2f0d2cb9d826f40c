// Helpers shared by the K2 translation units (k2_refine.cu: SWAR / mma.sync paths and the binning prepass; k2_umma.cu: the
// tcgen05 path): shape classes and pack geometry, the refinement tables, candidate-tile access in the staged regions,
// SAD, cp.async staging.  Everything has internal linkage (each translation unit gets its own copy of the tables).
#pragma once
#include "fme_common.cuh"

namespace {

#ifndef FME_K2_WARPS
#define FME_K2_WARPS 12
#endif
constexpr int K2_WARPS = FME_K2_WARPS;
constexpr int K2_STAGE_BYTES = 10496;  // two staging buffers of max P*(h+1)*RW + 16 bytes (64x64: 65*80 = 5200; 8x8: 32*152 = 4864)
constexpr int K2_ORG2_BYTES = 2048;    // second source tile of every lane for PUs with more than 32 tiles (8 rows x 32 lanes x 8 B)
constexpr int K2_SMEM_PER_WARP = K2_STAGE_BYTES + K2_ORG2_BYTES;

// TEncSearch.cpp:212-236
__constant__ int8_t c_refineH[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
// candidates served by staging step s, in table order: first index and count (see the step list in k2_pack)
__constant__ int8_t c_stepFirst[12] = {0, 3, 1, 5, 1, 2, 3, 4, 5, 6, 7, 8};
__constant__ int8_t c_stepCount[12] = {1, 2, 2, 4, 1, 1, 1, 1, 1, 1, 1, 1};
__constant__ int8_t c_refineQ[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

// table index of the q-th candidate a pack evaluates: the half-pel candidates in staging order (c_stepFirst / c_stepCount:
// H0 | H3 H4 | H1 H2 | H5..H8), then Q1..Q8
__constant__ int8_t c_seqI[17] = {0, 3, 4, 1, 2, 5, 6, 7, 8, 1, 2, 3, 4, 5, 6, 7, 8};
// the q-th candidate's offset (ox, oy) in {-1, 0, 1}^2 (c_refineH for q < 9, c_refineQ afterwards), as the two PRMT
// selectors that pick byte ox + 1 of bitsX and byte oy + 1 of bitsY: 0x444b | 0x444b << 16
#define K2_SEL(ox, oy) (0x4440u | (unsigned)((ox) + 1) | ((0x4440u | (unsigned)((oy) + 1)) << 16))
__constant__ unsigned c_seqSel[17] = {
    K2_SEL(0, 0), K2_SEL(-1, 0), K2_SEL(1, 0), K2_SEL(0, -1), K2_SEL(0, 1), K2_SEL(-1, -1), K2_SEL(1, -1), K2_SEL(-1, 1), K2_SEL(1, 1),
    K2_SEL(0, -1), K2_SEL(0, 1), K2_SEL(-1, -1), K2_SEL(1, -1), K2_SEL(-1, 0), K2_SEL(1, 0), K2_SEL(-1, 1), K2_SEL(1, 1)};
#undef K2_SEL

struct ClassInfo {
  int w, h;
  int ts;        // tile size 8 or 4
  int tilesX;    // tiles per PU row
  int tiles;     // tiles per PU
  int units;     // lane work units per PU: one 8x8 tile, or a pair of 4x4 tiles
  int lanes;     // lanes per PU: units rounded up to a power of two (<= 32) so that per-PU sums are xor-shuffles
  int P;         // PUs per pack
};

__host__ __device__ inline ClassInfo class_info(int cls) {
  ClassInfo c;
  c.w = fme_index_dim(cls >> 3);
  c.h = fme_index_dim(cls & 7);
  c.ts = ((c.w & 7) == 0 && (c.h & 7) == 0) ? 8 : 4;
  c.tilesX = c.w / c.ts;
  c.tiles = c.tilesX * (c.h / c.ts);
  c.units = c.ts == 8 ? c.tiles : c.tiles / 2;  // 4x4-tiled PUs always have an even tile count
  c.lanes = 1;
  while (c.lanes < c.units && c.lanes < 32) c.lanes <<= 1;
  c.P = 32 / c.lanes;
  return c;
}

// PUs per pack: the binning is the same for every K2 path (k2_group_mma walks a whole pack as sub-items of eight tiles).
__host__ __device__ inline int pack_pus(const ClassInfo& c, int packMode) {
  return packMode == 2 ? 4 * c.P : c.P;  // 2: one warp-load for each of the four worker warps of a k2_refine_umma CTA
}

// A candidate tile in shared memory: rows are `pitchWords` 32-bit words apart (the staged row pitch is a multiple of
// 4 bytes), so the byte misalignment of the tile is the same in every row and is resolved once.
struct CandTile {
  const unsigned* base;  // word containing the first byte of row 0
  unsigned shift;        // 8 * (byte address & 3)
  int pitchWords;
};
__device__ __forceinline__ CandTile cand_tile(const uint8_t* cand, int candPitch) {
  CandTile t;
  unsigned addr = (unsigned)(size_t)cand;
  t.base = reinterpret_cast<const unsigned*>(cand - (addr & 3u));
  t.shift = (addr & 3u) * 8u;
  t.pitchWords = candPitch >> 2;
  return t;
}
__device__ __forceinline__ void cand_row8(const CandTile& t, int r, unsigned& lo, unsigned& hi) {
  const unsigned* p = t.base + r * t.pitchWords;
  unsigned w0 = p[0], w1 = p[1], w2 = p[2];
  lo = __funnelshift_r(w0, w1, t.shift);
  hi = __funnelshift_r(w1, w2, t.shift);
}
__device__ __forceinline__ unsigned cand_row4(const CandTile& t, int r) {
  const unsigned* p = t.base + r * t.pitchWords;
  return __funnelshift_r(p[0], p[1], t.shift);
}

template <typename OrgRow>
__device__ __forceinline__ unsigned sad8x8(OrgRow orgRow, const uint8_t* cand, int candPitch) {
  unsigned s = 0;
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned c0, c1, o0, o1;
    cand_row8(ct, r, c0, c1);
    orgRow(r, o0, o1);
    s = __vsadu4(o0, c0) + s;
    s = __vsadu4(o1, c1) + s;
  }
  return s;
}
__device__ __forceinline__ unsigned sad4x4(const unsigned (&o)[4], const uint8_t* cand, int candPitch) {
  unsigned s = 0;
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 4; ++r) s += __vsadu4(o[r], cand_row4(ct, r));
  return s;
}

// 8 / 4 bytes at an arbitrary byte address in global memory (aligned 32-bit loads + funnel shift)
__device__ __forceinline__ void ldg_row8(const uint8_t* base, unsigned& lo, unsigned& hi) {
  size_t addr = (size_t)base;
  const unsigned* p = reinterpret_cast<const unsigned*>(addr & ~(size_t)3);
  unsigned sh = (unsigned)(addr & 3) * 8u;
  unsigned w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh);
}
__device__ __forceinline__ unsigned ldg_row4(const uint8_t* base) {
  size_t addr = (size_t)base;
  const unsigned* p = reinterpret_cast<const unsigned*>(addr & ~(size_t)3);
  unsigned sh = (unsigned)(addr & 3) * 8u;
  return __funnelshift_r(__ldg(p), __ldg(p + 1), sh);
}

// TComRdCost.cpp:172-185
__device__ __forceinline__ int golomb_bits(int v) {
  unsigned u = (v <= 0) ? (((unsigned)(-v)) << 1) + 1u : ((unsigned)v << 1);
  return 1 + 2 * (31 - __clz(u));
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
template <int A>
__device__ __forceinline__ void cp_async_g(void* smemDst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
  if constexpr (A == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
  else if constexpr (A == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
template <int A>
__device__ __forceinline__ void cp_async_s(unsigned sa, const void* gsrc) {
  if constexpr (A == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
  else if constexpr (A == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-pack staging geometry (uniform across the warp).
struct StageGeom {
  int RW;   // staged row bytes (multiple of the copy granule A)
  int G;    // granules per row
  int RB;   // bytes per staged region = (h + 1) * RW
};
// One staging step for one lane.  A PU's region is copied by the lanes that serve that PU (ci.lanes, a power of two):
// lane `sub` of the group owns granule column gi = sub % Gp of the rows rowSub, rowSub + rowStep, ... where
// Gp = min(pow2ceil(G), lanes) and rowStep = lanes / Gp.  Lanes of one PU read neighbouring granules of the same rows
// (sector-coalesced); the walk is one pointer bump per row, no index arithmetic.  Single-lane groups (8x8, 8x4, 4x8)
// have G == 2 and copy both granules of every row themselves (second = true); elsewhere Gp >= G.
template <int A>
__device__ __forceinline__ void stage_rows(unsigned dst, const uint8_t* src, int rowSub, int rowStep, int rows, int RW,
                                           int pitch, bool second) {
  const long long srcStep = (long long)rowStep * pitch;
  const int dstStep = rowStep * RW;
#pragma unroll 4
  for (int r = rowSub; r < rows; r += rowStep) {
    cp_async_s<A>(dst, src);
    if (second) cp_async_s<A>(dst + A, src + A);
    src += srcStep;
    dst += dstStep;
  }
}

}  // namespace

// Binning prepass (k2_refine.cu): memset + k2_count + k2_scatter on stream s.  packMode 0: a pack is the P PUs that fill
// one warp; packMode 2: a pack is 4 P PUs, one warp-load for each of the four worker warps of a k2_refine_umma CTA.
cudaError_t fme_k2_bin(const fme_pu* d_pus, int n, fme_result* d_res, int wantBi, int biServed, const FmeK2Scratch& sc,
                       int packMode, int numSMs, cudaStream_t s, int64_t* launches);
// The tcgen05 path (k2_umma.cu): binning with packMode 2 + k2_refine_umma, uni-prediction records, Hadamard distortion.
cudaError_t fme_launch_k2_umma(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_pu* d_pus, int n,
                               fme_result* d_res, const FmeCostLut& costLut, int biServed, const FmeK2Scratch& sc, int numSMs,
                               cudaStream_t s, int64_t* launches);
