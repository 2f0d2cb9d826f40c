// K0 (integer 3x3 error surface), block-level filter / distortion parity kernels and motion compensation.
#include "fme_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// K0: 3x3 integer error surface around the best integer MV.
// Metric of xTZSearchHelp (TEncSearch.cpp:1085-1090, 1156-1166) as selected by
// TComRdCost::setDistParam(pattern, ref, stride, dp) (TComRdCost.cpp:200-229): SSE for widths
// 4/8/16/32/64, SAD12/24/48 otherwise, the SADs on every second row (<< 1) when FEN is on and rows > 8.
// Raster order [TL,T,TR,L,C,R,BL,B,BR] = array_e[0..3], C, array_e[4..7] (TEncSearch.cpp:88, 1341-1376).
// Plane 0 is the padded integer-pel copy.
// ------------------------------------------------------------------------------------------------
// 8 lanes per PU (4 PUs per warp), arranged as lanesX x lanesY over (4-sample column groups, rows) with lanesX the
// power-of-two factor of w/4 (no divisions).  A lane's item is one 4-sample group of one source row against the 3x3
// neighbourhood: three reference rows, each read as three aligned words, aligned once (2 funnel shifts) and then
// shifted to the three horizontal offsets; SSE per word = VABSDIFF4 + dp4a(d, d), SAD per word = VABSDIFF4.ACC.
template <bool SAD>
__device__ __forceinline__ void k0_accumulate(unsigned (&acc)[9], const uint8_t* __restrict__ src, int orgPitch,
                                              const uint8_t* __restrict__ ref, int pitch, int w, int rows, int step,
                                              int sub) {
  const int groups = w >> 2;
  const int lanesX = min(groups & -groups, 8);  // 1,2,1,4,2,8,4,8 for w/4 = 1,2,3,4,6,8,12,16
  const int lxShift = 31 - __clz(lanesX);
  const int lanesY = 8 >> lxShift;
  const int cg0 = sub & (lanesX - 1), r0 = sub >> lxShift;
  // source words: aligned in contract (PU x is a multiple of 4, the picture base and pitch are 128-byte aligned)
  const unsigned so = (unsigned)((size_t)src & 3) * 8u;
  // reference window of column group 0, row -1: bytes -1 .. 6; the alignment is the same for every item of the PU
  const uint8_t* rp0 = ref - pitch - 1;
  const unsigned a8 = (unsigned)((size_t)rp0 & 3) * 8u;
  const unsigned* rw0 = reinterpret_cast<const unsigned*>((size_t)rp0 & ~(size_t)3);
  const int pitchW = pitch >> 2, orgPitchW = orgPitch >> 2;
  const unsigned* sw0 = reinterpret_cast<const unsigned*>((size_t)src & ~(size_t)3);
  for (int rr = r0; rr < rows; rr += lanesY) {
    const int r = rr * step;
    for (int cg = cg0; cg < groups; cg += lanesX) {
      const unsigned* spw = sw0 + r * orgPitchW + cg;
      const unsigned o = so ? __funnelshift_r(__ldg(spw), __ldg(spw + 1), so) : __ldg(spw);
      const unsigned* rpw = rw0 + r * pitchW + cg;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const unsigned w0 = __ldg(rpw + dy * pitchW), w1 = __ldg(rpw + dy * pitchW + 1), w2 = __ldg(rpw + dy * pitchW + 2);
        const unsigned x0 = __funnelshift_r(w0, w1, a8), x1 = __funnelshift_r(w1, w2, a8);  // bytes -1..2, 3..6
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const unsigned v = dx == 0 ? x0 : __funnelshift_r(x0, x1, 8 * dx);
          if (SAD) {
            acc[dy * 3 + dx] = __vsadu4(o, v) + acc[dy * 3 + dx];
          } else {
            const unsigned d = __vabsdiffu4(o, v);
            acc[dy * 3 + dx] = __dp4a(d, d, acc[dy * 3 + dx]);
          }
        }
      }
    }
  }
}

// SSE widths 4 / 8 / 16 / 32 / 64 (the common case): a lane owns a strip of NW words (4 NW samples) and a block of
// consecutive source rows, and slides a three-row window down the reference: every reference row is loaded and aligned
// ONCE (NW + 2 words) and serves the three vertical offsets of three source rows -- 2 NW + 2 loads per 4 NW samples instead
// of 11 per 4 samples in the per-item form above (K0 was bound by L1 wavefronts, not by arithmetic).
template <int NW>
__device__ __forceinline__ void k0_sse_strip(unsigned (&acc)[9], const uint8_t* __restrict__ src, int orgPitch,
                                             const uint8_t* __restrict__ ref, int pitch, int nrows) {
  const uint8_t* rp = ref - pitch - 1;  // row -1, byte -1 of the strip
  const unsigned a8 = (unsigned)((size_t)rp & 3) * 8u;
  const unsigned* rw = reinterpret_cast<const unsigned*>((size_t)rp & ~(size_t)3);
  const unsigned* sw = reinterpret_cast<const unsigned*>(src);  // 4-byte aligned (checked by the caller)
  const int pitchW = pitch >> 2, orgPitchW = orgPitch >> 2;
  unsigned A[NW + 1], B[NW + 1], C[NW + 1];
  auto load_row = [&](const unsigned* p, unsigned (&x)[NW + 1]) {
    unsigned w[NW + 2];
#pragma unroll
    for (int j = 0; j < NW + 2; ++j) w[j] = __ldg(p + j);
#pragma unroll
    for (int j = 0; j <= NW; ++j) x[j] = __funnelshift_r(w[j], w[j + 1], a8);  // bytes 4j-1 .. 4j+2 of the row
  };
  auto row_sse = [&](const unsigned (&o)[NW], const unsigned (&x)[NW + 1], int dy) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const unsigned v = dx == 0 ? x[j] : __funnelshift_r(x[j], x[j + 1], 8 * dx);
        const unsigned d = __vabsdiffu4(o[j], v);
        acc[dy * 3 + dx] = __dp4a(d, d, acc[dy * 3 + dx]);
      }
  };
  load_row(rw, A);
  load_row(rw + pitchW, B);
#pragma unroll 1
  for (int r = 0; r < nrows; ++r) {
    load_row(rw + (r + 2) * pitchW, C);
    unsigned o[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) o[j] = __ldg(sw + r * orgPitchW + j);
    row_sse(o, A, 0);
    row_sse(o, B, 1);
    row_sse(o, C, 2);
#pragma unroll
    for (int j = 0; j <= NW; ++j) { A[j] = B[j]; B[j] = C[j]; }
  }
}

// runIf != nullptr: the number of flagged records counted by k2_count earlier on the stream; zero -> nothing to do.
__global__ void __launch_bounds__(256) k0_int_surface(fme_pu* __restrict__ pus, int n, const uint8_t* __restrict__ planes,
                                                      const uint8_t* __restrict__ org, const FmeGeom g, int fen,
                                                      const int* __restrict__ runIf) {
  if (runIf && *runIf == 0) return;
  const int sub = threadIdx.x & 7;                                      // lane within the PU's 8-lane group
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;        // PU group index
  const int nGrp = (gridDim.x * blockDim.x) >> 3;
  for (int i0 = grp; i0 < ((n + 3) & ~3); i0 += nGrp) {               // all 4 groups of a warp iterate together
    const int i = i0;
    const bool valid = i < n;
    // the first 12 bytes of the record (position, size, slot, flags, integer MV); records are 4-byte aligned
    int px = 0, py = 0, w = 0, h = 0, slot = 0, flags = 0, mvx = 0, mvy = 0;
    if (valid) {
      const unsigned* rec = reinterpret_cast<const unsigned*>(&pus[i]);
      const unsigned r0 = rec[0], r1 = rec[1], r2 = rec[2];
      px = (int)(short)(r0 & 0xffff); py = (int)(short)(r0 >> 16);
      w = r1 & 0xff; h = (r1 >> 8) & 0xff; slot = (r1 >> 16) & 0xff; flags = r1 >> 24;
      mvx = (int)(short)(r2 & 0xffff); mvy = (int)(short)(r2 >> 16);
    }
    const bool doit = valid && (flags & FME_PU_ERR_ON_GPU) && !(flags & FME_PU_BI) &&
                      fme_hevc_pu_shape(w, h);  // FME_PU_BI records carry the other list's prediction in err[]
    unsigned acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0;
    int step = 1;
    if (doit) {
      const bool useSad = (w == 12 || w == 24 || w == 48);
      step = (useSad && fen && h > 8) ? 2 : 1;
      const int X = min(max(px + mvx, -(g.M - 8)), g.W + g.M - 8 - w);
      const int Y = min(max(py + mvy, -(g.M - 8)), g.H + g.M - 8 - h);
      const int ox = min(max(px, 0), g.W - w), oy = min(max(py, 0), g.H - h);
      const uint8_t* ref = planes + (size_t)min(slot, g.numSlots - 1) * g.slotBytes +
                           (size_t)((Y + g.M) * g.pitch + (X + g.M));
      const uint8_t* src = org + (size_t)(oy * g.orgPitch + ox);
      if (useSad) {
        k0_accumulate<true>(acc, src, g.orgPitch, ref, g.pitch, w, step == 2 ? h >> 1 : h, step, sub);
      } else if (((size_t)src & 3) == 0) {
        // 8 lanes = strips x row blocks: strips of 16 samples (w >= 16), else one strip of w samples
        const int strips = w >= 16 ? w >> 4 : 1;                 // 1, 2, 4
        const int sShift = strips >> 1;                          // log2
        const int lanesY = 8 >> sShift;
        const int strip = sub & (strips - 1), rb = sub >> sShift;
        const int rowsPer = (h + lanesY - 1) >> (3 - sShift);
        const int r0 = rb * rowsPer;
        const int nrows = min(rowsPer, h - r0);
        if (nrows > 0) {
          const uint8_t* s2 = src + (size_t)r0 * g.orgPitch + strip * 16;
          const uint8_t* f2 = ref + (size_t)r0 * g.pitch + strip * 16;
          if (w >= 16) k0_sse_strip<4>(acc, s2, g.orgPitch, f2, g.pitch, nrows);
          else if (w == 8) k0_sse_strip<2>(acc, s2, g.orgPitch, f2, g.pitch, nrows);
          else k0_sse_strip<1>(acc, s2, g.orgPitch, f2, g.pitch, nrows);
        }
      } else {
        k0_accumulate<false>(acc, src, g.orgPitch, ref, g.pitch, w, h, 1, sub);  // out-of-contract (unaligned) source position
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      unsigned v = acc[k];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (step == 2) v <<= 1;  // uiSum <<= iSubShift (TComRdCost.cpp, xGetSAD12/24/48)
      acc[k] = v;
    }
    if (doit) {
      // every lane of the group holds all nine totals: lane k stores err[k], lane 0 also err[8]
      unsigned v = acc[0];
#pragma unroll
      for (int k = 1; k < 8; ++k)
        if (sub == k) v = acc[k];
      pus[i].err[sub] = v;
      if (sub == 0) pus[i].err[8] = acc[8];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Block-level TComInterpolationFilter::filter / filterCopy (TComInterpolationFilter.cpp:94-257)
// ------------------------------------------------------------------------------------------------
__constant__ int8_t c_luma[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0},
                                    {-1, 4, -10, 58, 17, -5, 1, 0},
                                    {-1, 4, -11, 40, 40, -11, 4, -1},
                                    {0, 1, -5, 17, 58, -10, 4, -1}};
__constant__ int8_t c_chroma[8][4] = {{0, 64, 0, 0},   {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                      {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

// src points at the block's first sample; the halo (N/2-1 before, N/2 after) must be present.
__global__ void k_filter_block(int isVertical, int ntaps, int isFirst, int isLast, int bitDepth,
                               const int16_t* __restrict__ src, int srcStride, int16_t* __restrict__ dst,
                               int dstStride, int w, int h, int frac, int isLuma) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int headRoom = max(2, 14 - bitDepth);
  const int16_t* s = src + (ptrdiff_t)y * srcStride + x;
  int16_t out;
  if (frac == 0) {  // filterCopy, TComInterpolationFilter.cpp:94-154
    int v = s[0];
    if (isFirst == isLast) {
      out = (int16_t)v;
    } else if (isFirst) {
      int16_t t = (int16_t)(v << headRoom);
      out = (int16_t)(t - 8192);
    } else {
      int16_t t = (int16_t)((v + 8192 + (1 << (headRoom - 1))) >> headRoom);
      int maxVal = (1 << bitDepth) - 1;
      if (t < 0) t = 0;
      if (t > maxVal) t = (int16_t)maxVal;
      out = t;
    }
  } else {  // filter<N,...>, TComInterpolationFilter.cpp:172-257
    int cStride = isVertical ? srcStride : 1;
    s -= (ntaps / 2 - 1) * cStride;
    int shift = 6, offset, maxVal;
    if (isLast) {
      shift += isFirst ? 0 : headRoom;
      offset = 1 << (shift - 1);
      offset += isFirst ? 0 : (8192 << 6);
      maxVal = (1 << bitDepth) - 1;
    } else {
      shift -= isFirst ? headRoom : 0;
      offset = isFirst ? -(8192 << shift) : 0;
      maxVal = 0;
    }
    int sum = 0;
    for (int k = 0; k < ntaps; ++k) {
      int c = isLuma ? c_luma[frac][k] : c_chroma[frac][k];
      sum += (int)s[(ptrdiff_t)k * cStride] * c;
    }
    int16_t v = (int16_t)((sum + offset) >> shift);
    if (isLast) {
      if (v < 0) v = 0;
      if (v > maxVal) v = (int16_t)maxVal;
    }
    out = v;
  }
  dst[(ptrdiff_t)y * dstStride + x] = out;
}

// ------------------------------------------------------------------------------------------------
// Block-level distortion (TComRdCost.cpp:359-1495) on Pel buffers, one warp per block pair.
// kind 0: SSE (or SAD12/24/48 with subShift), 1: HADs, 2: SADs with subShift.
// ------------------------------------------------------------------------------------------------
__device__ int had_tile_generic(const int16_t* org, int os, const int16_t* cur, int cs, int n) {
  int d[64];
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) d[r * n + c] = (int)org[r * os + c] - (int)cur[r * cs + c];
  for (int len = 1; len < n; len <<= 1) {
    for (int r = 0; r < n; ++r)
      for (int i = 0; i < n; i += 2 * len)
        for (int j = i; j < i + len; ++j) {
          int a = d[r * n + j], b = d[r * n + j + len];
          d[r * n + j] = a + b;
          d[r * n + j + len] = a - b;
        }
  }
  for (int len = 1; len < n; len <<= 1) {
    for (int c = 0; c < n; ++c)
      for (int i = 0; i < n; i += 2 * len)
        for (int j = i; j < i + len; ++j) {
          int a = d[j * n + c], b = d[(j + len) * n + c];
          d[j * n + c] = a + b;
          d[(j + len) * n + c] = a - b;
        }
  }
  int s = 0;
  for (int i = 0; i < n * n; ++i) s += abs(d[i]);
  if (n == 8) return (s + 2) >> 2;
  if (n == 4) return (s + 1) >> 1;
  return s;
}

__global__ void __launch_bounds__(128) k_dist_blocks(int kind, const int16_t* __restrict__ org, int os,
                                                     const int16_t* __restrict__ cur, int cs, int w, int h,
                                                     int bitDepth, int subShift, int nBlocks,
                                                     uint32_t* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nBlocks) return;
  const int16_t* o = org + (size_t)warp * h * os;
  const int16_t* c = cur + (size_t)warp * h * cs;
  bool sadIntMe = (kind == 0) && (w == 12 || w == 24 || w == 48);
  unsigned acc = 0;
  if (kind == 1) {
    int n = ((h % 8 == 0) && (w % 8 == 0)) ? 8 : ((h % 4 == 0) && (w % 4 == 0)) ? 4 : 2;
    int tx = w / n, tiles = tx * (h / n);
    for (int t = lane; t < tiles; t += 32) {
      int x = (t % tx) * n, y = (t / tx) * n;
      acc += (unsigned)had_tile_generic(o + y * os + x, os, c + y * cs + x, cs, n);
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    acc >>= (bitDepth - 8);
  } else if (kind == 2 || sadIntMe) {
    int step = 1 << subShift;
    int rows = h / step;
    for (int t = lane; t < rows * w; t += 32) {
      int r = (t / w) * step, x = t % w;
      acc += (unsigned)abs((int)o[r * os + x] - (int)c[r * cs + x]);
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    acc <<= subShift;
    acc >>= (bitDepth - 8);
  } else {
    int sh = (bitDepth - 8) << 1;
    for (int t = lane; t < h * w; t += 32) {
      int r = t / w, x = t % w;
      int d = (int)o[r * os + x] - (int)c[r * cs + x];
      acc += (unsigned)((d * d) >> sh);
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
  }
  if (lane == 0) out[warp] = acc;
}

// ------------------------------------------------------------------------------------------------
// Motion compensation, uni-prediction (TComPrediction::xPredInterBlk, TComPrediction.cpp:643-681).
// Luma: the clipped MC sample at quarter-pel MV equals plane P[mvY&3][mvX&3] at the integer offset
// (SURVEY.md A.1), so luma MC is a gather.  Chroma (4:2:0): 4-tap at 1/8 pel with the reference's
// two-stage rounding; chroma planes are padded copies.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chroma_sample(const uint8_t* plane, int pitch, int x, int y, int xFrac, int yFrac) {
  const uint8_t* p = plane + (ptrdiff_t)y * pitch + x;
  if (yFrac == 0) {
    if (xFrac == 0) return p[0];
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c_chroma[xFrac][k] * (int)p[k - 1];
    return min(max((s + 32) >> 6, 0), 255);
  }
  if (xFrac == 0) {
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c_chroma[yFrac][k] * (int)p[(ptrdiff_t)(k - 1) * pitch];
    return min(max((s + 32) >> 6, 0), 255);
  }
  int acc = 2048 + (8192 << 6);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint8_t* q = p + (ptrdiff_t)(j - 1) * pitch;
    int t = -8192;
#pragma unroll
    for (int k = 0; k < 4; ++k) t += c_chroma[xFrac][k] * (int)q[k - 1];
    acc += c_chroma[yFrac][j] * t;
  }
  return min(max(acc >> 12, 0), 255);
}

__global__ void __launch_bounds__(256) k_mc(const fme_mc_pu* __restrict__ pus, int n, const uint8_t* __restrict__ planes,
                                            const uint8_t* __restrict__ cb, const uint8_t* __restrict__ cr,
                                            const FmeGeom g, int16_t* __restrict__ dstY, int16_t* __restrict__ dstCb,
                                            int16_t* __restrict__ dstCr) {
  int i = blockIdx.x;
  if (i >= n) return;
  fme_mc_pu p = pus[i];
  int w = p.w, h = p.h;
  // luma
  {
    int fx = p.mvX & 3, fy = p.mvY & 3;
    int X = min(max(p.x + (p.mvX >> 2), -(g.M - 8)), g.W + g.M - 8 - w);
    int Y = min(max(p.y + (p.mvY >> 2), -(g.M - 8)), g.H + g.M - 8 - h);
    const uint8_t* src = planes + (size_t)min((int)p.refSlot, g.numSlots - 1) * g.slotBytes + (size_t)(fy * 4 + fx) * g.planeBytes +
                         (size_t)(Y + g.M) * g.pitch + (X + g.M);
    int16_t* d = dstY + (size_t)i * 64 * 64;
    for (int t = threadIdx.x; t < w * h; t += blockDim.x) {
      int r = t / w, c = t % w;
      d[r * 64 + c] = src[(size_t)r * g.pitch + c];
    }
  }
  // chroma
  if (cb && cr && dstCb && dstCr) {
    int cw = w >> 1, ch = h >> 1;
    int xFrac = p.mvX & 7, yFrac = p.mvY & 7;
    int X = min(max((p.x >> 1) + (p.mvX >> 3), -(g.Mc - 4)), g.Wc + g.Mc - 4 - cw) + g.Mc;
    int Y = min(max((p.y >> 1) + (p.mvY >> 3), -(g.Mc - 4)), g.Hc + g.Mc - 4 - ch) + g.Mc;
    const uint8_t* pcb = cb + (size_t)min((int)p.refSlot, g.numSlots - 1) * g.cPlaneBytes;
    const uint8_t* pcr = cr + (size_t)min((int)p.refSlot, g.numSlots - 1) * g.cPlaneBytes;
    int16_t* dcb = dstCb + (size_t)i * 32 * 32;
    int16_t* dcr = dstCr + (size_t)i * 32 * 32;
    for (int t = threadIdx.x; t < cw * ch; t += blockDim.x) {
      int r = t / cw, c = t % cw;
      dcb[r * 32 + c] = (int16_t)chroma_sample(pcb, g.cPitch, X + c, Y + r, xFrac, yFrac);
      dcr[r * 32 + c] = (int16_t)chroma_sample(pcr, g.cPitch, X + c, Y + r, xFrac, yFrac);
    }
  }
}

// Records without the error grid -> full records flagged for the K0 pass (which fills err[]).
static_assert(sizeof(fme_pu_head) == 16 && sizeof(fme_pu) == 52, "record layouts");
__global__ void k_expand_heads(const fme_pu_head* __restrict__ heads, int n, fme_pu* __restrict__ pus) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint4 v = reinterpret_cast<const uint4*>(heads)[i];
  // flags is byte 7 of the record.  FME_PU_BI is dropped: a bi-predictive record names the other list's prediction in
  // err[], which a head does not have (the synchronous entry points reject such heads).
  v.y = (v.y & ~((unsigned)FME_PU_BI << 24)) | ((unsigned)FME_PU_ERR_ON_GPU << 24);
  unsigned* d = reinterpret_cast<unsigned*>(&pus[i]);  // 52-byte records are 4-byte aligned
  d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}

// Compact records (head + nine 24-bit grid values) -> full records; no K0 flag, no FME_PU_BI (see fme_b200.h).
static_assert(sizeof(fme_pu_compact) == 44, "compact record layout");
// A block serves 256 records: their 11 x 256 words come in and their 13 x 256 words go out as coalesced streams through
// shared memory (a thread reading its own 44-byte record and writing its own 52-byte one touches 4x the sectors: 0.064 ms
// per 858 000 records against 0.02 ms).
__global__ void __launch_bounds__(256) k_expand_compact(const fme_pu_compact* __restrict__ recs, int n, fme_pu* __restrict__ pus) {
  __shared__ unsigned s_in[256 * 11];
  __shared__ unsigned s_out[256 * 13];
  const int first = blockIdx.x * 256, cnt = min(256, n - first);
  const unsigned* src = reinterpret_cast<const unsigned*>(recs) + (size_t)first * 11;  // 44-byte records are 4-byte aligned
  for (int k = threadIdx.x; k < cnt * 11; k += 256) s_in[k] = __ldg(src + k);
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    const unsigned* w = s_in + threadIdx.x * 11;   // 11 is odd: conflict-free
    unsigned* d = s_out + threadIdx.x * 13;        // 13 is odd: conflict-free
    d[0] = w[0];
    d[1] = w[1] & ~((unsigned)(FME_PU_BI | FME_PU_ERR_ON_GPU) << 24);  // flags is byte 7 of the record
    d[2] = w[2];
    d[3] = w[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) {   // value k = bytes 3k .. 3k+2 of the 28-byte tail
      const int b = 3 * k, q = b >> 2, r = b & 3;
      const unsigned lo = w[4 + q], hi = q + 1 < 7 ? w[5 + q] : 0u;
      d[4 + k] = __funnelshift_r(lo, hi, 8 * r) & 0xffffffu;
    }
  }
  __syncthreads();
  unsigned* dst = reinterpret_cast<unsigned*>(pus) + (size_t)first * 13;   // 52-byte records are 4-byte aligned
  for (int k = threadIdx.x; k < cnt * 13; k += 256) dst[k] = s_out[k];
}

// Host-supplied error grids for some of the expanded heads: err[] filled, FME_PU_ERR_ON_GPU cleared (K0 skips them).
static_assert(sizeof(fme_err_grid) == 40, "grid layout");
__global__ void k_apply_grids(const fme_err_grid* __restrict__ grids, int nGrids, fme_pu* __restrict__ pus, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nGrids) return;
  const unsigned* gw = reinterpret_cast<const unsigned*>(&grids[j]);  // 40-byte entries are 4-byte aligned
  const int i = (int)gw[0];
  if (i < 0 || i >= n) return;
  unsigned* d = reinterpret_cast<unsigned*>(&pus[i]);
  d[1] &= ~((unsigned)FME_PU_ERR_ON_GPU << 24);  // flags is byte 7 of the record
#pragma unroll
  for (int k = 0; k < 9; ++k) d[4 + k] = gw[1 + k];
}

// ------------------------------------------------------------------------------------------------
// Motion compensation, bi-prediction: xPredInterBi with both lists valid (TComPrediction.cpp:575-621) =
// two xPredInterUni(bi = true) blocks of 14-bit intermediates (xPredInterBlk with isLast = !bi = false,
// TComPrediction.cpp:661-680) averaged by TComYuv::addAvg (TComYuv.cpp:354-409).  The intermediates are not
// recoverable from the clipped 8-bit planes, so they are filtered here from the integer plane P[0][0]:
//   yFrac == 0: sum c_x * s - 8192 (copy: (s << 6) - 8192) | xFrac == 0: same vertically |
//   else      : (sum_j c_y[j] * (sum_k c_x[k] * s - 8192)) >> 6      (first stage shift 0 at 8 bit, IF.cpp:94-257)
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ int bi_sample(const uint8_t* plane, int pitch, int x, int y, const int8_t* cx, const int8_t* cy,
                                         bool fracX, bool fracY) {
  const uint8_t* p = plane + (ptrdiff_t)y * pitch + x;
  constexpr int L = N / 2 - 1;  // taps to the left / above
  if (!fracY) {
    if (!fracX) return ((int)p[0] << 6) - 8192;
    int s = -8192;
#pragma unroll
    for (int k = 0; k < N; ++k) s += cx[k] * (int)p[k - L];
    return s;
  }
  if (!fracX) {
    int s = -8192;
#pragma unroll
    for (int k = 0; k < N; ++k) s += cy[k] * (int)p[(ptrdiff_t)(k - L) * pitch];
    return s;
  }
  int acc = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const uint8_t* q = p + (ptrdiff_t)(j - L) * pitch;
    int t = -8192;
#pragma unroll
    for (int k = 0; k < N; ++k) t += cx[k] * (int)q[k - L];
    acc += cy[j] * t;
  }
  return acc >> 6;
}
__device__ __forceinline__ int add_avg8(int a, int b) { return min(max((a + b + 64 + 2 * 8192) >> 7, 0), 255); }

__global__ void __launch_bounds__(256) k_mc_bi(const fme_mc_bi_pu* __restrict__ pus, int n, const uint8_t* __restrict__ planes,
                                               const uint8_t* __restrict__ cb, const uint8_t* __restrict__ cr,
                                               const FmeGeom g, int16_t* __restrict__ dstY, int16_t* __restrict__ dstCb,
                                               int16_t* __restrict__ dstCr) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const fme_mc_bi_pu p = pus[i];
  const int w = p.w, h = p.h;
  const int slot[2] = {min((int)p.refSlot0, g.numSlots - 1), min((int)p.refSlot1, g.numSlots - 1)};
  const int mvx[2] = {p.mv0X, p.mv1X}, mvy[2] = {p.mv0Y, p.mv1Y};
  {  // luma: 8-tap at quarter pel from plane P[0][0] of each slot
    const uint8_t* src[2];
    int X[2], Y[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      X[l] = min(max(p.x + (mvx[l] >> 2), -(g.M - 8)), g.W + g.M - 8 - w) + g.M;
      Y[l] = min(max(p.y + (mvy[l] >> 2), -(g.M - 8)), g.H + g.M - 8 - h) + g.M;
      src[l] = planes + (size_t)slot[l] * g.slotBytes;
    }
    int16_t* d = dstY + (size_t)i * 64 * 64;
    for (int t = threadIdx.x; t < w * h; t += blockDim.x) {
      const int r = t / w, c = t % w;
      const int a = bi_sample<8>(src[0], g.pitch, X[0] + c, Y[0] + r, c_luma[mvx[0] & 3], c_luma[mvy[0] & 3], mvx[0] & 3, mvy[0] & 3);
      const int b = bi_sample<8>(src[1], g.pitch, X[1] + c, Y[1] + r, c_luma[mvx[1] & 3], c_luma[mvy[1] & 3], mvx[1] & 3, mvy[1] & 3);
      d[r * 64 + c] = (int16_t)add_avg8(a, b);
    }
  }
  if (cb && cr && dstCb && dstCr) {  // chroma 4:2:0: 4-tap at 1/8 pel
    const int cw = w >> 1, ch = h >> 1;
    int X[2], Y[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      X[l] = min(max((p.x >> 1) + (mvx[l] >> 3), -(g.Mc - 4)), g.Wc + g.Mc - 4 - cw) + g.Mc;
      Y[l] = min(max((p.y >> 1) + (mvy[l] >> 3), -(g.Mc - 4)), g.Hc + g.Mc - 4 - ch) + g.Mc;
    }
    int16_t* dcb = dstCb + (size_t)i * 32 * 32;
    int16_t* dcr = dstCr + (size_t)i * 32 * 32;
    for (int t = threadIdx.x; t < cw * ch; t += blockDim.x) {
      const int r = t / cw, c = t % cw;
      int v[2][2];
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const int8_t* cx = c_chroma[mvx[l] & 7];
        const int8_t* cy = c_chroma[mvy[l] & 7];
        v[l][0] = bi_sample<4>(cb + (size_t)slot[l] * g.cPlaneBytes, g.cPitch, X[l] + c, Y[l] + r, cx, cy, mvx[l] & 7, mvy[l] & 7);
        v[l][1] = bi_sample<4>(cr + (size_t)slot[l] * g.cPlaneBytes, g.cPitch, X[l] + c, Y[l] + r, cx, cy, mvx[l] & 7, mvy[l] & 7);
      }
      dcb[r * 32 + c] = (int16_t)add_avg8(v[0][0], v[1][0]);
      dcr[r * 32 + c] = (int16_t)add_avg8(v[0][1], v[1][1]);
    }
  }
}

}  // namespace

cudaError_t fme_launch_k0(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, fme_pu* d_pus, int n,
                          int fen, cudaStream_t s, int64_t* launches, const int* d_runIf) {
  if (n <= 0) return cudaSuccess;
  int blocks = (n + 31) / 32;  // 32 PUs per 256-thread CTA
  if (blocks > 148 * 16) blocks = 148 * 16;
  k0_int_surface<<<blocks, 256, 0, s>>>(d_pus, n, d_planes, d_org, g, fen, d_runIf);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_filter(int isVertical, int ntaps, int isFirst, int isLast, int bitDepth, const int16_t* d_src,
                              int srcStride, int16_t* d_dst, int dstStride, int w, int h, int frac, int isLuma,
                              cudaStream_t s, int64_t* launches) {
  dim3 grid((w + 63) / 64, h);
  k_filter_block<<<grid, 64, 0, s>>>(isVertical, ntaps, isFirst, isLast, bitDepth, d_src, srcStride, d_dst, dstStride,
                                     w, h, frac, isLuma);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_dist(int kind, const int16_t* d_org, int orgStride, const int16_t* d_cur, int curStride, int w,
                            int h, int bitDepth, int subShift, int nBlocks, uint32_t* d_out, cudaStream_t s,
                            int64_t* launches) {
  int blocks = (nBlocks + 3) / 4;
  k_dist_blocks<<<blocks, 128, 0, s>>>(kind, d_org, orgStride, d_cur, curStride, w, h, bitDepth, subShift, nBlocks,
                                       d_out);
  ++*launches;
  return cudaGetLastError();
}

// fme_result (16 bytes) -> fme_result8 (include/fme_b200.h): every field of the MV group is in {-1, 0, 1}
__global__ void k_pack_results(const fme_result* __restrict__ res, int n, fme_result8* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 v = *reinterpret_cast<const uint4*>(&res[i]);  // x: half/qter bytes, y: cost, z: nn bytes, w: class
  auto two = [](unsigned bytes, int k) { return (unsigned)((int)(signed char)(bytes >> (8 * k)) + 1) & 3u; };
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) m |= two(v.x, k) << (2 * k) | two(v.z, k) << (8 + 2 * k);
  m |= (v.w & 63u) << 16;
  *reinterpret_cast<uint2*>(&out[i]) = make_uint2(v.y, m);
}
cudaError_t fme_launch_pack_results(const fme_result* d_res, int n, fme_result8* d_out, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  k_pack_results<<<(n + 255) / 256, 256, 0, s>>>(d_res, n, d_out);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_expand_heads(const fme_pu_head* d_heads, int n, fme_pu* d_pus, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  k_expand_heads<<<(n + 255) / 256, 256, 0, s>>>(d_heads, n, d_pus);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_expand_compact(const fme_pu_compact* d_recs, int n, fme_pu* d_pus, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  k_expand_compact<<<(n + 255) / 256, 256, 0, s>>>(d_recs, n, d_pus);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_apply_grids(const fme_err_grid* d_grids, int nGrids, fme_pu* d_pus, int n, cudaStream_t s,
                                   int64_t* launches) {
  if (nGrids <= 0) return cudaSuccess;
  k_apply_grids<<<(nGrids + 255) / 256, 256, 0, s>>>(d_grids, nGrids, d_pus, n);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_mc_bi(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_cb, const uint8_t* d_cr,
                             const fme_mc_bi_pu* d_pus, int n, int16_t* d_y, int16_t* d_cbOut, int16_t* d_crOut,
                             cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  k_mc_bi<<<n, 256, 0, s>>>(d_pus, n, d_planes, d_cb, d_cr, g, d_y, d_cbOut, d_crOut);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_mc(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_cb, const uint8_t* d_cr,
                          const fme_mc_pu* d_pus, int n, int16_t* d_y, int16_t* d_cbOut, int16_t* d_crOut,
                          cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  k_mc<<<n, 256, 0, s>>>(d_pus, n, d_planes, d_cb, d_cr, g, d_y, d_cbOut, d_crOut);
  ++*launches;
  return cudaGetLastError();
}
