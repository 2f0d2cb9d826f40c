// K2: batched half-/quarter-pel refinement with SATD or SAD + MV-bit cost
// (replaces TEncSearch::xPatternRefinement TEncSearch.cpp:1591-1645 and the distortion functions it
// calls, TComRdCost::xGetHADs / xCalcHADs8x8 / xCalcHADs4x4 TComRdCost.cpp:1234-1495, xGetSAD* :359-855,
// driven as in xPatternSearchFracDIF TEncSearch.cpp:5232-5269).
//
// Candidate addressing (derived from TEncSearch.cpp:1613-1623, 6343-6531): with
// (X,Y) = PU origin + integer MV and total quarter-pel offset (qx,qy) in [-3,3]^2 the candidate block is
// plane P[qy&3][qx&3] at integer origin (X + (qx>>2), Y + (qy>>2)).
//
// Work decomposition (v1):
//   * a prepass groups PU indices by shape class (w,h) so that a warp always works on same-shape PUs;
//   * one warp = one "pack" of P PUs whose tiles (8x8, or 4x4 when w or h is not a multiple of 8,
//     TComRdCost.cpp:1446-1492) fill the 32 lanes: lane = (PU in pack, tile in PU);
//   * per refinement stage the warp stages the candidate regions of its PUs from the HBM/L2-resident
//     planes into shared memory with 4-byte cp.async (rows stay source-aligned), then every lane
//     computes the SATD of its own tile in registers: two 16-bit residuals per 32-bit register (SWAR),
//     butterflies as plain 32-bit adds, |a+b|+|a-b| = 2 max(|a|,|b|) for the intra-register stage;
//   * per-PU sums by __reduce_add_sync over the PU's lane group, MV cost from an exact host-built LUT,
//     first-minimum argmin in the reference's table order (TEncSearch.cpp:212-236, strict < at :1634).
#include "fme_common.cuh"

namespace {

constexpr int K2_WARPS = 8;
constexpr int K2_THREADS = K2_WARPS * 32;
constexpr int K2_SMEM_PER_WARP = 17920;  // max over shape classes of one staging step (64x64 half stage: 4*65*68)
constexpr int K2_MAX_PACK = 32;

// TEncSearch.cpp:212-236
__constant__ int8_t c_refineH[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
__constant__ int8_t c_refineQ[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

struct ClassInfo {
  int w, h;
  int ts;        // tile size 8 or 4
  int tilesX;    // tiles per PU row
  int tiles;     // tiles per PU
  int P;         // PUs per pack
};

__host__ __device__ inline ClassInfo class_info(int cls) {
  ClassInfo c;
  c.w = fme_index_dim(cls >> 3);
  c.h = fme_index_dim(cls & 7);
  c.ts = ((c.w & 7) == 0 && (c.h & 7) == 0) ? 8 : 4;
  c.tilesX = c.w / c.ts;
  c.tiles = c.tilesX * (c.h / c.ts);
  c.P = c.tiles >= 32 ? 1 : 32 / c.tiles;
  return c;
}

// ------------------------------------------------------------------------------------------------
// prepass: counting sort of PU indices by shape class
// ------------------------------------------------------------------------------------------------
__global__ void k2_count(const fme_pu* __restrict__ pus, int n, int* __restrict__ classCount) {
  __shared__ int s_cnt[FME_MAX_CLASSES];
  for (int i = threadIdx.x; i < FME_MAX_CLASSES; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int wi = fme_dim_index(pus[i].w), hi = fme_dim_index(pus[i].h);
    if (wi >= 0 && hi >= 0) atomicAdd(&s_cnt[wi * 8 + hi], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FME_MAX_CLASSES; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&classCount[i], s_cnt[i]);
}

__global__ void k2_scan(const int* __restrict__ classCount, int* __restrict__ classOffset, int* __restrict__ packOffset,
                        int* __restrict__ classCursor, int* __restrict__ workCounter) {
  if (threadIdx.x == 0) {
    int off = 0, packs = 0;
    for (int c = 0; c < FME_MAX_CLASSES; ++c) {
      classOffset[c] = off;
      packOffset[c] = packs;
      classCursor[c] = 0;
      int cnt = classCount[c];
      off += cnt;
      if (cnt) packs += (cnt + class_info(c).P - 1) / class_info(c).P;
    }
    classOffset[FME_MAX_CLASSES] = off;
    packOffset[FME_MAX_CLASSES] = packs;
    *workCounter = 0;
  }
}

__global__ void k2_scatter(const fme_pu* __restrict__ pus, int n, const int* __restrict__ classOffset,
                           int* __restrict__ classCursor, int* __restrict__ order) {
  __shared__ int s_cnt[FME_MAX_CLASSES];
  __shared__ int s_base[FME_MAX_CLASSES];
  // one contiguous chunk per block so that local ranks are well defined
  int chunk = (n + gridDim.x - 1) / gridDim.x;
  int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  for (int i = threadIdx.x; i < FME_MAX_CLASSES; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    int wi = fme_dim_index(pus[i].w), hh = fme_dim_index(pus[i].h);
    if (wi >= 0 && hh >= 0) atomicAdd(&s_cnt[wi * 8 + hh], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FME_MAX_CLASSES; i += blockDim.x) {
    s_base[i] = s_cnt[i] ? classOffset[i] + atomicAdd(&classCursor[i], s_cnt[i]) : 0;
    s_cnt[i] = 0;
  }
  __syncthreads();
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    int wi = fme_dim_index(pus[i].w), hh = fme_dim_index(pus[i].h);
    if (wi >= 0 && hh >= 0) {
      int c = wi * 8 + hh;
      order[s_base[c] + atomicAdd(&s_cnt[c], 1)] = i;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tile distortion in registers
// ------------------------------------------------------------------------------------------------
// x264-style packed |.| on two 16-bit lanes held carry-tolerantly in one 32-bit word
// (value = hi*65536 + lo with signed lo); returns true unsigned fields (|hi| , |lo|).
__device__ __forceinline__ unsigned abs2(unsigned a) {
  unsigned s = ((a >> 15) & 0x10001u) * 0xffffu;
  return (a + s) ^ s;
}
// max of the two 16-bit fields of m, zero-extended
__device__ __forceinline__ unsigned hmax2(unsigned m) {
  unsigned sw = __byte_perm(m, 0, 0x1032);
  return __vmaxu2(m, sw) & 0xffffu;
}

// Read 8 bytes at arbitrary byte address from shared memory (3 aligned words + 2 funnel shifts).
__device__ __forceinline__ void lds_row8(const uint8_t* base, unsigned& lo, unsigned& hi) {
  unsigned addr = (unsigned)(size_t)base;  // only the low bits matter for alignment
  const unsigned* p = reinterpret_cast<const unsigned*>(base - (addr & 3u));
  unsigned sh = (addr & 3u) * 8u;
  unsigned w0 = p[0], w1 = p[1], w2 = p[2];
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh);
}
__device__ __forceinline__ unsigned lds_row4(const uint8_t* base) {
  unsigned addr = (unsigned)(size_t)base;
  const unsigned* p = reinterpret_cast<const unsigned*>(base - (addr & 3u));
  unsigned sh = (addr & 3u) * 8u;
  return __funnelshift_r(p[0], p[1], sh);
}

// SATD of one 8x8 tile: xCalcHADs8x8 (TComRdCost.cpp:1330-1425).  o[] holds the source tile as
// 16 words of u8x4 (row r -> o[2r], o[2r+1]); cand points at the candidate tile's row 0 in smem.
__device__ __forceinline__ unsigned satd8x8(const unsigned (&o)[16], const uint8_t* cand, int candPitch) {
  unsigned d[32];  // d[4r + j] = (res(r,2j), res(r,2j+1)) packed lo/hi
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned c0, c1;
    lds_row8(cand + r * candPitch, c0, c1);
    d[4 * r + 0] = __byte_perm(o[2 * r], 0, 0x4140) - __byte_perm(c0, 0, 0x4140);
    d[4 * r + 1] = __byte_perm(o[2 * r], 0, 0x4342) - __byte_perm(c0, 0, 0x4342);
    d[4 * r + 2] = __byte_perm(o[2 * r + 1], 0, 0x4140) - __byte_perm(c1, 0, 0x4140);
    d[4 * r + 3] = __byte_perm(o[2 * r + 1], 0, 0x4342) - __byte_perm(c1, 0, 0x4342);
  }
  // horizontal: column-index bits 1 and 2 (bit 0 lives inside a word and is folded into the final max)
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned a0 = d[4 * r] + d[4 * r + 1], a1 = d[4 * r] - d[4 * r + 1];
    unsigned a2 = d[4 * r + 2] + d[4 * r + 3], a3 = d[4 * r + 2] - d[4 * r + 3];
    d[4 * r] = a0 + a2; d[4 * r + 2] = a0 - a2;
    d[4 * r + 1] = a1 + a3; d[4 * r + 3] = a1 - a3;
  }
  // vertical: three stages over the row index
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unsigned v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = d[4 * r + j];
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int i = 0; i < 8; i += 2 * len)
#pragma unroll
        for (int k = i; k < i + len; ++k) {
          unsigned a = v[k], b = v[k + len];
          v[k] = a + b;
          v[k + len] = a - b;
        }
#pragma unroll
    for (int r = 0; r < 8; ++r) d[4 * r + j] = v[r];
  }
  // last horizontal stage + abs: |lo+hi| + |lo-hi| = 2 max(|lo|,|hi|)
  unsigned sum = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) sum += hmax2(abs2(d[i]));
  return (2 * sum + 2) >> 2;  // TComRdCost.cpp:1421
}

// SATD of one 4x4 tile: xCalcHADs4x4 (TComRdCost.cpp:1234-1328).  o[r] = source row r (u8x4).
__device__ __forceinline__ unsigned satd4x4(const unsigned (&o)[4], const uint8_t* cand, int candPitch) {
  unsigned d[8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    unsigned c = lds_row4(cand + r * candPitch);
    d[2 * r + 0] = __byte_perm(o[r], 0, 0x4140) - __byte_perm(c, 0, 0x4140);
    d[2 * r + 1] = __byte_perm(o[r], 0, 0x4342) - __byte_perm(c, 0, 0x4342);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    unsigned a = d[2 * r], b = d[2 * r + 1];
    d[2 * r] = a + b;
    d[2 * r + 1] = a - b;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    unsigned v0 = d[j], v1 = d[2 + j], v2 = d[4 + j], v3 = d[6 + j];
    unsigned a0 = v0 + v1, a1 = v0 - v1, a2 = v2 + v3, a3 = v2 - v3;
    d[j] = a0 + a2; d[4 + j] = a0 - a2; d[2 + j] = a1 + a3; d[6 + j] = a1 - a3;
  }
  unsigned sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += hmax2(abs2(d[i]));
  return (2 * sum + 1) >> 1;  // TComRdCost.cpp:1325
}

__device__ __forceinline__ unsigned sad8x8(const unsigned (&o)[16], const uint8_t* cand, int candPitch) {
  unsigned s = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned c0, c1;
    lds_row8(cand + r * candPitch, c0, c1);
    s = __vsadu4(o[2 * r], c0) + s;
    s = __vsadu4(o[2 * r + 1], c1) + s;
  }
  return s;
}
__device__ __forceinline__ unsigned sad4x4(const unsigned (&o)[4], const uint8_t* cand, int candPitch) {
  unsigned s = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) s += __vsadu4(o[r], lds_row4(cand + r * candPitch));
  return s;
}

template <int TS>
__device__ __forceinline__ unsigned tile_dist(const unsigned* o, const uint8_t* cand, int candPitch, bool had) {
  if constexpr (TS == 8) {
    const unsigned(&oo)[16] = *reinterpret_cast<const unsigned(*)[16]>(o);
    return had ? satd8x8(oo, cand, candPitch) : sad8x8(oo, cand, candPitch);
  } else {
    const unsigned(&oo)[4] = *reinterpret_cast<const unsigned(*)[4]>(o);
    return had ? satd4x4(oo, cand, candPitch) : sad4x4(oo, cand, candPitch);
  }
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

__device__ __forceinline__ void cp_async4(void* smemDst, const void* gsrc) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct PackPu {      // per-PU state of the pack, kept in shared memory (one per PU in the pack)
  int pu;            // global PU index, -1 = padding
  int X, Y;          // PU origin + integer MV, picture coordinates
  int planeBase;     // byte offset of the slot's plane 0 (as 32-bit units of 16 B to fit: see below)
};

template <int TS>
__device__ __forceinline__ void k2_pack(const ClassInfo ci, const int* __restrict__ order, int first, int count,
                                        const fme_pu* __restrict__ pus, fme_result* __restrict__ res,
                                        const uint8_t* __restrict__ planes, const uint8_t* __restrict__ org,
                                        const FmeGeom g, const uint32_t* __restrict__ costLut, int useHad,
                                        uint8_t* smem, int* s_pu, int* s_X, int* s_Y, long long* s_slotOff, int* s_win) {
  const int lane = threadIdx.x & 31;
  const int w = ci.w, h = ci.h;
  const int T = ci.tiles;
  const int rounds = (T + 31) / 32;           // > 1 only when a single PU has more than 32 tiles (P == 1)
  const int lanesPerPu = T >= 32 ? 32 : T;
  const int myPu = lane / lanesPerPu;         // PU slot in the pack served by this lane
  const bool laneActive = myPu < count && myPu < ci.P;
  const int tile0 = lane - myPu * lanesPerPu;  // tile index in round 0
  const unsigned groupMask =
      lanesPerPu == 32 ? 0xffffffffu : (((1u << lanesPerPu) - 1u) << (myPu * lanesPerPu));

  // ---- pack table ----
  if (lane < ci.P) {
    int idx = lane < count ? order[first + lane] : -1;
    s_pu[lane] = idx;
    if (idx >= 0) {
      fme_pu p = pus[idx];
      // The reference clips MVs so that reads stay inside its padded planes (TComDataCU.cpp:2773-2786:
      // X in [-71, W+7]); the clamp below only guards device memory against out-of-contract records.
      s_X[lane] = min(max(p.x + p.mvIntX, -(g.M - 8)), g.W + g.M - 8 - w);
      s_Y[lane] = min(max(p.y + p.mvIntY, -(g.M - 8)), g.H + g.M - 8 - h);
      s_slotOff[lane] = (long long)min((int)p.refSlot, g.numSlots - 1) * (long long)g.slotBytes;
    }
  }
  __syncwarp();

  // ---- source tile(s) of this lane into registers ----
  constexpr int OW = TS == 8 ? 16 : 4;  // words per tile
  unsigned o[2][OW];
  int predX = 0, predY = 0, mvIntX = 0, mvIntY = 0, lossless = 0, puIdx = -1;
  if (laneActive) {
    puIdx = s_pu[myPu];
    fme_pu p = pus[puIdx];
    predX = p.mvPredX; predY = p.mvPredY; mvIntX = p.mvIntX; mvIntY = p.mvIntY;
    lossless = p.flags & FME_PU_LOSSLESS;
#pragma unroll
    for (int rd = 0; rd < 2; ++rd) {
      int t = tile0 + 32 * rd;
      if (rd < rounds && t < T) {
        int tx = t % ci.tilesX, ty = t / ci.tilesX;
        int ox = min(max((int)p.x, 0), g.W - w), oy = min(max((int)p.y, 0), g.H - h);
        const uint8_t* src = org + (size_t)(oy + ty * TS) * g.orgPitch + ox + tx * TS;
#pragma unroll
        for (int r = 0; r < TS; ++r) {
          if constexpr (TS == 8) ldg_row8(src + (size_t)r * g.orgPitch, o[rd][2 * r], o[rd][2 * r + 1]);
          else o[rd][r] = ldg_row4(src + (size_t)r * g.orgPitch);
        }
      }
    }
  }
  const bool had = useHad && !lossless;

  // region geometry for staging
  const int RW = w + 4;                 // staged row bytes: covers [x0 & ~3, x0 + w + 1)
  const int wordsPerRow = RW >> 2;

  // ================= half-pel stage: planes (0,0) (0,2) (2,0) (2,2), origin (X-1, Y-1), (w+1) x (h+1)
  {
    const int rows = h + 1;
    const int regionBytes = rows * RW;
    const int perPu = 4 * regionBytes;
    const int wordsPerRegion = rows * wordsPerRow;
    const int total = count * 4 * wordsPerRegion;
    for (int i = lane; i < total; i += 32) {
      int j = i / (4 * wordsPerRegion);
      int rem = i - j * 4 * wordsPerRegion;
      int k = rem / wordsPerRegion;
      rem -= k * wordsPerRegion;
      int row = rem / wordsPerRow, wd = rem - row * wordsPerRow;
      int plane = (k & 1) * 2 + (k >> 1) * 8;  // k: 0->(0,0) 1->(0,2) 2->(2,0) 3->(2,2); index fy*4+fx
      int x0 = s_X[j] - 1, y0 = s_Y[j] - 1;
      const uint8_t* src = planes + s_slotOff[j] + (size_t)plane * g.planeBytes +
                           (size_t)(y0 + g.M + row) * g.pitch + ((x0 + g.M) & ~3) + wd * 4;
      cp_async4(smem + j * perPu + k * regionBytes + row * RW + wd * 4, src);
    }
    cp_async_wait_all();
    __syncwarp();

    unsigned best = 0xffffffffu;
    int bestI = 0;
    const int align = laneActive ? ((s_X[myPu] - 1 + g.M) & 3) : 0;
#pragma unroll 1
    for (int i = 0; i < 9; ++i) {
      int hx = c_refineH[i][0], hy = c_refineH[i][1];
      int qx = 2 * hx, qy = 2 * hy;
      int k = ((qx & 3) ? 1 : 0) + ((qy & 3) ? 2 : 0);
      int dx = 1 + (qx >> 2), dy = 1 + (qy >> 2);  // offset inside the staged region (origin X-1, Y-1)
      unsigned dist = 0;
      if (laneActive) {
#pragma unroll
        for (int rd = 0; rd < 2; ++rd) {
          int t = tile0 + 32 * rd;
          if (rd < rounds && t < T) {
            int tx = t % ci.tilesX, ty = t / ci.tilesX;
            const uint8_t* c = smem + myPu * perPu + k * regionBytes + (dy + ty * TS) * RW + align + dx + tx * TS;
            dist += tile_dist<TS>(o[rd], c, RW, had);
          }
        }
      }
      if (laneActive) {
        dist = __reduce_add_sync(groupMask, dist);
        // cost scale 1 (TEncSearch.cpp:4531): ((int<<1)+h) << 1 - pred
        int bits = golomb_bits((((mvIntX << 1) + hx) << 1) - predX) + golomb_bits((((mvIntY << 1) + hy) << 1) - predY);
        dist += costLut[bits];
        if (dist < best) { best = dist; bestI = i; }
      }
    }
    __syncwarp();

    // ================= quarter-pel stage: 8 candidates around the half-pel winner, two steps of 4
    const int bhx = c_refineH[bestI][0], bhy = c_refineH[bestI][1];
    unsigned qBest = best;  // candidate 0 of the quarter stage is the half-pel winner itself (same block, same bits)
    int qBestI = 0;
    const int qrows = h;
    const int qRegionBytes = qrows * RW;
    const int qPerPu = 4 * qRegionBytes;
    const int qWordsPerRegion = qrows * wordsPerRow;
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
      // publish each PU's half-pel winner for the cooperative staging loop
      if (laneActive && tile0 == 0) s_win[myPu] = (bhx & 0xff) | ((bhy & 0xff) << 8);
      __syncwarp();
      const int totalQ = count * 4 * qWordsPerRegion;
      for (int i = lane; i < totalQ; i += 32) {
        int j = i / (4 * qWordsPerRegion);
        int rem = i - j * 4 * qWordsPerRegion;
        int k = rem / qWordsPerRegion;
        rem -= k * qWordsPerRegion;
        int row = rem / wordsPerRow, wd = rem - row * wordsPerRow;
        int wv = s_win[j];
        int jhx = (int)(int8_t)(wv & 0xff), jhy = (int)(int8_t)((wv >> 8) & 0xff);
        int ci9 = 1 + step * 4 + k;  // candidate index in s_acMvRefineQ
        int qx = 2 * jhx + c_refineQ[ci9][0], qy = 2 * jhy + c_refineQ[ci9][1];
        int plane = (qy & 3) * 4 + (qx & 3);
        int x0 = s_X[j] + (qx >> 2), y0 = s_Y[j] + (qy >> 2);
        const uint8_t* src = planes + s_slotOff[j] + (size_t)plane * g.planeBytes +
                             (size_t)(y0 + g.M + row) * g.pitch + ((x0 + g.M) & ~3) + wd * 4;
        cp_async4(smem + j * qPerPu + k * qRegionBytes + row * RW + wd * 4, src);
      }
      cp_async_wait_all();
      __syncwarp();
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        int ci9 = 1 + step * 4 + k;
        int qx = 2 * bhx + c_refineQ[ci9][0], qy = 2 * bhy + c_refineQ[ci9][1];
        unsigned dist = 0;
        if (laneActive) {
          int al = (s_X[myPu] + (qx >> 2) + g.M) & 3;
#pragma unroll
          for (int rd = 0; rd < 2; ++rd) {
            int t = tile0 + 32 * rd;
            if (rd < rounds && t < T) {
              int tx = t % ci.tilesX, ty = t / ci.tilesX;
              const uint8_t* c = smem + myPu * qPerPu + k * qRegionBytes + (ty * TS) * RW + al + tx * TS;
              dist += tile_dist<TS>(o[rd], c, RW, had);
            }
          }
          dist = __reduce_add_sync(groupMask, dist);
          // cost scale 0 (TEncSearch.cpp:5260): (((int<<1)+half)<<1) + q - pred
          int vx = (((mvIntX << 1) + bhx) << 1) + c_refineQ[ci9][0];
          int vy = (((mvIntY << 1) + bhy) << 1) + c_refineQ[ci9][1];
          dist += costLut[golomb_bits(vx - predX) + golomb_bits(vy - predY)];
          if (dist < qBest) { qBest = dist; qBestI = ci9; }
        }
      }
      __syncwarp();
    }

    if (laneActive && tile0 == 0) {
      fme_result* r = &res[puIdx];
      r->halfX = (int8_t)bhx; r->halfY = (int8_t)bhy;
      r->qterX = c_refineQ[qBestI][0]; r->qterY = c_refineQ[qBestI][1];
      r->cost = qBest;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(K2_THREADS, 1)
k2_refine(const fme_pu* __restrict__ pus, fme_result* __restrict__ res, const uint8_t* __restrict__ planes,
          const uint8_t* __restrict__ org, const FmeGeom g, const uint32_t* __restrict__ costLutG, int useHad,
          const int* __restrict__ order, const int* __restrict__ classOffset, const int* __restrict__ packOffset,
          int* __restrict__ workCounter) {
  extern __shared__ __align__(16) uint8_t dynSmem[];
  __shared__ uint32_t s_lut[FME_COST_LUT_SIZE];
  __shared__ int s_packOff[FME_MAX_CLASSES + 1];
  __shared__ int s_classOff[FME_MAX_CLASSES + 1];
  __shared__ int s_pu[K2_WARPS][K2_MAX_PACK];
  __shared__ int s_X[K2_WARPS][K2_MAX_PACK];
  __shared__ int s_Y[K2_WARPS][K2_MAX_PACK];
  __shared__ long long s_slot[K2_WARPS][K2_MAX_PACK];
  __shared__ int s_win[K2_WARPS][K2_MAX_PACK];

  for (int i = threadIdx.x; i < FME_COST_LUT_SIZE; i += blockDim.x) s_lut[i] = costLutG[i];
  for (int i = threadIdx.x; i <= FME_MAX_CLASSES; i += blockDim.x) {
    s_packOff[i] = packOffset[i];
    s_classOff[i] = classOffset[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = dynSmem + warp * K2_SMEM_PER_WARP;
  const int totalPacks = s_packOff[FME_MAX_CLASSES];

  while (true) {
    int pack = 0;
    if (lane == 0) pack = atomicAdd(workCounter, 1);
    pack = __shfl_sync(0xffffffffu, pack, 0);
    if (pack >= totalPacks) break;
    // class of this pack: last c with packOff[c] <= pack (skipping empty classes)
    int cls = 0;
    for (int c = 0; c < FME_MAX_CLASSES; ++c)
      if (s_packOff[c] <= pack && s_packOff[c + 1] > pack) { cls = c; break; }
    ClassInfo ci = class_info(cls);
    int inClass = s_classOff[cls + 1] - s_classOff[cls];
    int first = (pack - s_packOff[cls]) * ci.P;
    int count = min(ci.P, inClass - first);
    first += s_classOff[cls];
    if (ci.ts == 8)
      k2_pack<8>(ci, order, first, count, pus, res, planes, org, g, s_lut, useHad, smem, s_pu[warp], s_X[warp],
                 s_Y[warp], s_slot[warp], s_win[warp]);
    else
      k2_pack<4>(ci, order, first, count, pus, res, planes, org, g, s_lut, useHad, smem, s_pu[warp], s_X[warp],
                 s_Y[warp], s_slot[warp], s_win[warp]);
  }
}

__global__ void k_clear_results(fme_result* res, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint4 z = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&res[i]) = z;
  }
}

}  // namespace

cudaError_t fme_launch_clear_results(fme_result* d_res, int n, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  k_clear_results<<<(n + 255) / 256, 256, 0, s>>>(d_res, n);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_k2(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_pu* d_pus, int n,
                          fme_result* d_res, const uint32_t* d_costLut, int useHad, const FmeK2Scratch& sc,
                          int numSMs, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(sc.classCount, 0, sizeof(int) * FME_MAX_CLASSES, s);
  if (e != cudaSuccess) return e;
  int blocks = min(numSMs * 4, (n + 255) / 256);
  k2_count<<<blocks, 256, 0, s>>>(d_pus, n, sc.classCount);
  k2_scan<<<1, 32, 0, s>>>(sc.classCount, sc.classOffset, sc.packOffset, sc.classCursor, sc.workCounter);
  k2_scatter<<<blocks, 256, 0, s>>>(d_pus, n, sc.classOffset, sc.classCursor, sc.order);
  *launches += 3;
  static bool attrSet = false;
  const int smemBytes = K2_WARPS * K2_SMEM_PER_WARP;
  if (!attrSet) {
    e = cudaFuncSetAttribute(k2_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
    if (e != cudaSuccess) return e;
    attrSet = true;
  }
  k2_refine<<<numSMs, K2_THREADS, smemBytes, s>>>(d_pus, d_res, d_planes, d_org, g, d_costLut, useHad, sc.order,
                                                  sc.classOffset, sc.packOffset, sc.workCounter);
  ++*launches;
  return cudaGetLastError();
}
