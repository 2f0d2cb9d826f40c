// K1: sub-pel plane generation (replaces xExtDIFUpSamplingH/Q + filterHor/filterVer for whole
// reference pictures; reference: TEncSearch.cpp:6331-6532, TComInterpolationFilter.cpp:172-257).
//
// For an 8-bit picture R (edge-replicated to infinity, which is what TComPicYuv::extendPicBorder
// stores in the 80-sample margin) and HEVC luma taps c_f, the 16 planes are
//   T_fx(x,y)    = sum_k c_fx[k] R(x+k-3, y) - 8192            (first stage, shift 0 at 8 bit, int16)
//   P[fy][fx]    = clip255((sum_k c_fy[k] T_fx(x, y+k-3) + 2048 + (8192<<6)) >> 12)      fy != 0
//   P[0][fx]     = clip255((T_fx + 8192 + 32) >> 6)
// which is bit-identical to what the reference's per-PU two-stage filtering produces at every
// position.  Layout in HBM: planes[p = fy*4+fx][Hp][pitch] u8, picture sample (0,0) at [M][M].
//
// Persistent CTAs walk 128 x 16 output tiles (all 16 planes per tile), prefetching the next tile's input:
//   stage A  (136 x 39) u8 input tile -> shared memory, coordinates clamped to the picture
//   stage B  horizontal 8-tap for fx = 0..3 with dp4a (u8 samples x s8 taps) -> int16 in shared memory
//   stage C  vertical 8-tap for fy = 1..3 with dp2a (s16 x s8 tap pairs) on 4-row register blocks, fy = 0 by
//            shift; cvt.pack.sat (I2IP) narrowing, u8x4 stores
// Algorithmic HBM bytes: 1 B read + 16 B written per padded sample (plane 0 is the padded copy).
#include "fme_common.cuh"
#include <type_traits>

namespace {

constexpr int TW = 128;           // output tile width  (one 128-byte line per plane row)
constexpr int TH = 16;            // output tile height
constexpr int IN_W = TW + 8;      // input columns x0-4 .. x0+TW+3 (word aligned)
constexpr int IN_H = TH + 7;      // input rows    y0-3 .. y0+TH+3
constexpr int K1_THREADS = 256;

__device__ __forceinline__ int dp4a_u8s8(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_lo(int a, int b, int c) {
  int d;
  asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi(int a, int b, int c) {
  int d;
  asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

#define PACK4(a, b, c, d) \
  ((int)(((unsigned)(a)&0xffu) | (((unsigned)(b)&0xffu) << 8) | (((unsigned)(c)&0xffu) << 16) | (((unsigned)(d)&0xffu) << 24)))
// TComInterpolationFilter.cpp:57-63 luma taps, packed as s8x4 (low half, high half)
__constant__ int c_lumaLo[4] = {PACK4(0, 0, 0, 64), PACK4(-1, 4, -10, 58), PACK4(-1, 4, -11, 40), PACK4(0, 1, -5, 17)};
__constant__ int c_lumaHi[4] = {PACK4(0, 0, 0, 0), PACK4(17, -5, 1, 0), PACK4(40, -11, 4, -1), PACK4(58, -10, 4, -1)};

// cvt.pack.sat.u8.s32: d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)   (SASS: I2IP.U8.S32.SAT)
__device__ __forceinline__ unsigned pack_sat_u8x2(int lo, int hi, unsigned upper) {
  unsigned d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
  return d;
}
__device__ __forceinline__ unsigned pack_sat_u8x4(int p0, int p1, int p2, int p3) {
  unsigned hi16, d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi16) : "r"(p3), "r"(p2), "r"(0));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(p1), "r"(p0), "r"(hi16));
  return d;
}

// Input-tile word `i` (row-major over IN_H x IN_W/4) of the tile at (x0, y0), coordinates clamped to the picture.
__device__ __forceinline__ unsigned k1_load_word(const uint8_t* __restrict__ pic, int picPitch, int W, int H, int pxBase,
                                                 int pyBase, int i) {
  const int r = i / (IN_W / 4), wq = i - r * (IN_W / 4);
  const int py = min(max(pyBase + r, 0), H - 1);
  const int px = pxBase + wq * 4;
  const uint8_t* row = pic + (size_t)py * picPitch;
  if (px >= 0 && px + 3 <= W - 1) return *reinterpret_cast<const unsigned*>(row + px);
  unsigned b0 = row[min(max(px + 0, 0), W - 1)], b1 = row[min(max(px + 1, 0), W - 1)];
  unsigned b2 = row[min(max(px + 2, 0), W - 1)], b3 = row[min(max(px + 3, 0), W - 1)];
  return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

constexpr int K1_IN_WORDS = IN_H * (IN_W / 4);
constexpr int K1_PREF = (K1_IN_WORDS + K1_THREADS - 1) / K1_THREADS;  // input words per thread

// Persistent: gridDim.x CTAs walk the tile list; the next tile's input words are fetched into registers before the
// current tile's filtering starts and parked in shared memory after it, so global latency hides behind stages B/C.
#ifndef FME_K1_CTAS
#define FME_K1_CTAS 4
#endif
__global__ void __launch_bounds__(K1_THREADS, FME_K1_CTAS)
k1_interp_planes(const uint8_t* __restrict__ pic, int picPitch, int W, int H, int M, int Wp, int Hp, int pitch,
                 size_t planeBytes, uint8_t* __restrict__ planes, int tilesX, unsigned tilesXRcp, int nTiles, int* __restrict__ tileCounter,
                 int tyBegin) {  // first tile row to produce (row-range launches of the banded multi-GPU mode)
  __shared__ __align__(16) uint8_t s_in[IN_H][IN_W];
  __shared__ __align__(16) int16_t s_t[4][IN_H][TW];

  const int tid = threadIdx.x;
  unsigned pref[K1_PREF];
  auto fetch = [&](int tile) {
    // tilesXRcp = ceil(2^32 / tilesX); 0 stands for tilesX == 1 (2^32 does not fit)
    const int tr = tilesXRcp ? (int)__umulhi((unsigned)tile, tilesXRcp) : tile, tx = tile - tr * tilesX;
    const int ty = tr + tyBegin;
    const int pxBase = tx * TW - M - 4, pyBase = ty * TH - M - 3;  // picture coords of input (0,0); pxBase % 4 == 0
#pragma unroll
    for (int k = 0; k < K1_PREF; ++k) {
      const int i = tid + k * K1_THREADS;
      pref[k] = i < K1_IN_WORDS ? k1_load_word(pic, picPitch, W, H, pxBase, pyBase, i) : 0u;
    }
  };
  auto park = [&]() {
#pragma unroll
    for (int k = 0; k < K1_PREF; ++k) {
      const int i = tid + k * K1_THREADS;
      if (i < K1_IN_WORDS) reinterpret_cast<unsigned*>(&s_in[0][0])[i] = pref[k];
    }
  };

  // Tiles beyond the first one per CTA are handed out dynamically (tileCounter[0]), so that all resident CTAs of an
  // SM finish together whatever the ratio of tiles to CTAs; the last CTA to leave re-arms the counters
  // (tileCounter[1] counts finished CTAs), so no memset is needed between launches.
  __shared__ int s_next;
  int tile = blockIdx.x;
  if (tile < nTiles) fetch(tile);
  while (tile < nTiles) {
    const int trI = tilesXRcp ? (int)__umulhi((unsigned)tile, tilesXRcp) : tile, txI = tile - trI * tilesX;
    const int x0 = txI * TW, y0 = (trI + tyBegin) * TH;  // padded-plane coordinates of the tile
    if (tid == 0) s_next = (int)gridDim.x + atomicAdd(&tileCounter[0], 1);
    __syncthreads();  // every thread is done with s_in / s_t of the previous tile; s_next is published
    park();
    const int nextTile = s_next;
    __syncthreads();
    if (nextTile < nTiles) fetch(nextTile);  // in flight during stages B and C

  // ---- stage B: horizontal filters -> int16 T_fx ---------------------------------------------
  // item = (row, quad of 4 output columns).  Output column x uses input columns x+1 .. x+8
  // (input column index = picture x + 4, taps reach x-3 .. x+4).
  for (int i = tid; i < IN_H * (TW / 4); i += K1_THREADS) {
    int r = i / (TW / 4), q = i - r * (TW / 4);
    const unsigned* src = reinterpret_cast<const unsigned*>(&s_in[r][q * 4]);
    unsigned w0 = src[0], w1 = src[1], w2 = src[2];
    unsigned lo[4], hi[4];
    lo[0] = __funnelshift_r(w0, w1, 8);  hi[0] = __funnelshift_r(w1, w2, 8);
    lo[1] = __funnelshift_r(w0, w1, 16); hi[1] = __funnelshift_r(w1, w2, 16);
    lo[2] = __funnelshift_r(w0, w1, 24); hi[2] = __funnelshift_r(w1, w2, 24);
    lo[3] = w1;                          hi[3] = w2;
    // fx = 0: (s << 6) - 8192, s = input column x+4 = byte 0 of w1 + ...
    {
      int t0 = (int)((w1)&0xff) * 64 - 8192, t1 = (int)((w1 >> 8) & 0xff) * 64 - 8192;
      int t2 = (int)((w1 >> 16) & 0xff) * 64 - 8192, t3 = (int)((w1 >> 24) & 0xff) * 64 - 8192;
      int2 o;
      o.x = (t0 & 0xffff) | (t1 << 16);
      o.y = (t2 & 0xffff) | (t3 << 16);
      *reinterpret_cast<int2*>(&s_t[0][r][q * 4]) = o;
    }
#pragma unroll
    for (int f = 1; f < 4; ++f) {
      int t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] = dp4a_u8s8(hi[k], c_lumaHi[f], dp4a_u8s8(lo[k], c_lumaLo[f], -8192));
      int2 o;
      o.x = (t[0] & 0xffff) | (t[1] << 16);
      o.y = (t[2] & 0xffff) | (t[3] << 16);
      *reinterpret_cast<int2*>(&s_t[f][r][q * 4]) = o;
    }
  }
  __syncthreads();

  // ---- stage C: vertical filters, 16 planes, u8x4 stores -------------------------------------
  // item = (quad of 4 columns, group of 4 output rows); the 11 intermediate rows a group needs are loaded once
  // and shared by its 4 rows x 3 vertical phases (register blocking), results are shifted, saturated and packed
  // with cvt.pack.sat (I2IP) and stored as one 32-bit word per plane row.
  // A 128 x 16 tile has 128 (quad, row group) items: the two halves of the CTA take different horizontal phases.
  static_assert((TH / 4) * (TW / 4) * 2 == K1_THREADS, "stage C mapping assumes 2 x 128 items");
  {
    const int i = tid & 127, fxHalf = tid >> 7;
    const int rg = i / (TW / 4), q = i - rg * (TW / 4);
    const int r0 = rg * 4;
    const int gy0 = y0 + r0, gx = x0 + q * 4;
    const bool inside = gx < Wp && gy0 < Hp;  // Wp is a multiple of 4
    uint8_t* outBase = planes + (size_t)gy0 * pitch + gx;
    const int planeBytes32 = (int)planeBytes;  // 16 planes of an 8K picture stay below 2^31 bytes
#pragma unroll 1
    for (int fx = fxHalf; fx < 4 && inside; fx += 2) {
      int2 v[11];  // rows r0 .. r0+10 of T_fx (output row r is centred on input row r+3)
#pragma unroll
      for (int k = 0; k < 11; ++k) v[k] = *reinterpret_cast<const int2*>(&s_t[fx][r0 + k][q * 4]);
      // fy = 0: (T + 8192 + 32) >> 6 on the centre rows; dp2a with taps (1,0) / (0,1) extracts a sign-extended half
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (gy0 + r < Hp) {
          const int2 c = v[r + 3];
          int a0 = dp2a_lo(c.x, 0x0001, 8224) >> 6, a1 = dp2a_lo(c.x, 0x0100, 8224) >> 6;
          int a2 = dp2a_lo(c.y, 0x0001, 8224) >> 6, a3 = dp2a_lo(c.y, 0x0100, 8224) >> 6;
          *reinterpret_cast<unsigned*>(outBase + (fx * planeBytes32 + r * pitch)) =
              pack_sat_u8x4(a0, a1, a2, a3);
        }
      }
      // columns two at a time (3,2 first, then 1,0 so that the second cvt.pack completes the 32-bit word):
      // vertical pairs (T(x,k), T(x,k+1)), k = 0..9
      unsigned hi16[3][4];
#pragma unroll
      for (int half = 1; half >= 0; --half) {
        int pa[10], pb[10];  // column 2*half and 2*half+1
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const int lo = half ? v[k].y : v[k].x, hi = half ? v[k + 1].y : v[k + 1].x;
          pa[k] = __byte_perm(lo, hi, 0x5410);
          pb[k] = __byte_perm(lo, hi, 0x7632);
        }
#pragma unroll
        for (int fy = 1; fy < 4; ++fy) {
          const int tl = c_lumaLo[fy], th = c_lumaHi[fy];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            int acc0 = 2048 + (8192 << 6), acc1 = acc0;
            acc0 = dp2a_lo(pa[r + 0], tl, acc0); acc1 = dp2a_lo(pb[r + 0], tl, acc1);
            acc0 = dp2a_hi(pa[r + 2], tl, acc0); acc1 = dp2a_hi(pb[r + 2], tl, acc1);
            acc0 = dp2a_lo(pa[r + 4], th, acc0); acc1 = dp2a_lo(pb[r + 4], th, acc1);
            acc0 = dp2a_hi(pa[r + 6], th, acc0); acc1 = dp2a_hi(pb[r + 6], th, acc1);
            if (half) {
              hi16[fy - 1][r] = pack_sat_u8x2(acc0 >> 12, acc1 >> 12, 0);
            } else if (gy0 + r < Hp) {
              *reinterpret_cast<unsigned*>(outBase + ((fy * 4 + fx) * planeBytes32 + r * pitch)) =
                  pack_sat_u8x2(acc0 >> 12, acc1 >> 12, hi16[fy - 1][r]);
            }
          }
        }
      }
    }
  }
    tile = nextTile;
  }  // tile loop
  if (tid == 0 && atomicAdd(&tileCounter[1], 1) == (int)gridDim.x - 1) {
    tileCounter[0] = 0;
    tileCounter[1] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K1 on the tensor pipe (FME_K1_PATH_MMA).  At 8 bit the first filter stage has shift 0 (IF.cpp:197-218), so a plane is
// the EXACT 2-D form  P[fy][fx] = clip255(floor((sum_ky sum_kx c_fy[ky] c_fx[kx] R(x+kx-3, y+ky-3) + 2048) / 4096))
// (the -8192 first-stage offset and the 8192<<6 second-stage offset cancel; fy = 0 is the tap row (0,0,0,64,0,0,0,0)),
// and both stages are small GEMMs against constant Toeplitz tap matrices:
//   H:  T[x][y]      = A1_fx[x][xin] * R[xin][y]      IMMA m16n8k32 u8 x s8 -> s32: 16 output columns x 8 rows from a
//                                                     32-column window; the B fragment is the two words a lane loaded
//   V:  P[(fy,r)][x] = A2[(fy,r)][yT] * T[yT][x]      HMMA m16n8k16 f16 -> f32: 2 phases x 8 rows, 8 columns, 15 T rows
// T (15 bits) enters V as two exact f16 operands scaled so that ONE tap matrix (the taps themselves) serves both: the
// IMMA accumulator starts from 0x54342000, so that one PRMT per pair builds the f16x2 words 0x3400|low byte =
// 0.25 + lo/4096 and 0x5400|high byte = 64 + hi/16 (the 64 is subtracted, the 0.25 folded into V's start value: the
// sum of the two is (T + 8192)/4096 + 0.25).  The D fragment of H *is* the B fragment of V (same lanes, same element
// order) when H puts columns on M and rows on N, so T never leaves the register file; a warp slides down a 16-column
// strip 8 rows at a time and keeps the previous T tile.  All f32 sums are multiples of 2^-12 below 2^10: exact.
// f32 -> u8 without the XU pipe (F2I runs at 16 lanes/clk/SM and was the limit of the first version):
// y = sat(D/256 + 2^-22), z = floor(y * (256 - 2^-15)) by FFMA.RM onto 2^23 -- the low mantissa byte is
// clip255(floor(D)): with u = D/256 a multiple of 2^-20, y*(256 - 2^-15) = 256u + e with 2^-15 < e <= 2^-14 < 2^-12 for
// 0 <= u < 1 (same floor as 256u), y = 0 for u < 0, and y = 1 gives 255.99997 for u >= 1.
// The warps of a CTA stage 64-byte plane rows in shared memory (16-byte chunks XOR-swizzled by row, conflict-free both
// ways) and write them out with 16-byte stores.
namespace k1m {

constexpr int kTap[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1},
                            {0, 1, -5, 17, 58, -10, 4, -1}};  // TComInterpolationFilter.cpp:57-63

// f16 bit pattern of c * 2^-s (|c| <= 64: exact, normal for s <= 14)
constexpr uint32_t f16_of(int c, int s) {
  if (c == 0) return 0;
  const uint32_t sign = c < 0 ? 0x8000u : 0u;
  const int a = c < 0 ? -c : c;
  int e = 0;
  while ((a >> (e + 1)) != 0) ++e;
  return sign | ((uint32_t)(e - s + 15) << 10) | (((uint32_t)a << (10 - e)) & 0x3ffu);
}
// H: A1 row m computes output column colOfRow(m) of the strip, chosen so that after V a lane owns 4 adjacent columns.
constexpr int colOfRow(int m) { return m < 8 ? 4 * (m / 2) + (m % 2) : 4 * ((m - 8) / 2) + 2 + (m % 2); }
// H: K index k is window column 8t + {0..7} of the lane t = (k % 16) / 4 that loaded it (word 0: k < 16, word 1: k >= 16)
constexpr int inColOfK(int k) { return 8 * ((k % 16) / 4) + 4 * (k / 16) + (k % 4); }
constexpr uint32_t a1_elem(int fx, int m, int k, bool natural = false) {
  const int idx = inColOfK(k) - (natural ? m : colOfRow(m)) - 5;  // window column 0 is strip column -8; taps reach x-3 .. x+4
  return idx >= 0 && idx < 8 ? (uint32_t)(kTap[fx][idx] & 0xff) : 0u;
}
constexpr uint32_t a2_elem(int mt, int m, int kk) {  // row m = (phase 2mt + m/8, output row m%8); T row kk is y-3+kk
  const int idx = kk - (m % 8);
  return idx >= 0 && idx < 8 ? f16_of(kTap[2 * mt + m / 8][idx], 0) : 0u;
}
struct Tables {
  uint32_t a1[4][32][4];     // [fx][lane] -> the lane's A fragment (m16n8k32 s8, row-major A)
  uint32_t a1n[4][32][4];    // the same with row m = column m (the UMMA path wants N in column order)
  uint32_t a2[2][2][32][4];  // [K halves exchanged?][m-tile][lane] (m16n8k16 f16, the taps themselves)
};
constexpr Tables make_tables() {
  Tables t{};
  for (int lane = 0; lane < 32; ++lane) {
    const int g = lane >> 2, q = lane & 3;
    for (int r = 0; r < 4; ++r) {
      const int m = g + (r & 1) * 8;
      for (int fx = 0; fx < 4; ++fx) {
        uint32_t v = 0, vn = 0;
        for (int j = 0; j < 4; ++j) {
          v |= a1_elem(fx, m, 4 * q + (r >> 1) * 16 + j) << (8 * j);
          vn |= a1_elem(fx, m, 4 * q + (r >> 1) * 16 + j, true) << (8 * j);
        }
        t.a1[fx][lane][r] = v;
        t.a1n[fx][lane][r] = vn;
      }
      const int k0 = 2 * q + (r >> 1) * 8;
      for (int mt = 0; mt < 2; ++mt) {
        t.a2[0][mt][lane][r] = a2_elem(mt, m, k0) | (a2_elem(mt, m, k0 + 1) << 16);
        t.a2[1][mt][lane][r] = a2_elem(mt, m, k0 ^ 8) | (a2_elem(mt, m, (k0 ^ 8) + 1) << 16);  // T rows 8..15 first
      }
    }
  }
  return t;
}
__device__ const Tables d_tables = make_tables();

__device__ __forceinline__ void hmma(float (&d)[4], const uint4& a, unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// first product of a chain: every accumulator element starts from the same constant (no per-chain register copies)
__device__ __forceinline__ void hmma_c(float (&d)[4], const uint4& a, unsigned b0, unsigned b1, float c) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(c));
}
__device__ __forceinline__ void imma_c(int (&d)[4], const uint4& a, unsigned b0, unsigned b1, int c) {
  asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "r"(c));
}
__device__ __forceinline__ unsigned hsub2u(unsigned a, unsigned b) {
  unsigned d;
  asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// clip255(floor(v)) in the low byte (upper bytes hold 0x4b0000), v a multiple of 2^-12 with |v| < 2^10
__device__ __forceinline__ unsigned floor_clip_u8(float v) {
  float y, z;
  asm("fma.rn.sat.f32 %0, %1, 0f3B800000, 0f34800000;" : "=f"(y) : "f"(v));           // sat(v / 256 + 2^-22)
  asm("fma.rm.f32 %0, %1, 0f437FFFFF, 0f4B000000;" : "=f"(z) : "f"(y));                // floor(y * (256 - 2^-15)) + 2^23
  return __float_as_uint(z);
}

#ifndef FME_K1M_WARPS
#define FME_K1M_WARPS 4
#endif
#ifndef FME_K1M_CTAS
#define FME_K1M_CTAS 4
#endif
constexpr int WARPS = FME_K1M_WARPS, THREADS = 32 * WARPS;
constexpr int BW = 16 * WARPS;            // CTA block: 128 (64) plane columns = one (half a) line per plane row
constexpr int SLAB = 16 * 8 * BW;         // one iteration's output: 16 planes x 8 rows x BW bytes
constexpr int CHUNKS = BW / 16;           // 16-byte chunks per slab row
constexpr int RING = 4;                   // slabs in flight (see the main loop)
// slab rows are XOR-swizzled by 16-byte chunk so that both the per-warp 4-byte writes (8 rows x one chunk) and the
// 16-byte read-back (whole rows) touch every bank once
__device__ __forceinline__ int swz(int chunk, int row) { return chunk ^ (WARPS == 8 ? row : WARPS == 4 ? (row >> 1) : (row >> 2)); }

struct TTile { unsigned lo[4][2], hi[4][2]; };  // per fx: B fragments of V for the strip's two 8-column halves

__global__ void __launch_bounds__(THREADS, FME_K1M_CTAS)
k1_interp_mma(const uint8_t* __restrict__ pic, int picPitch, int W, int H, int M, int Wp, int Hp, int pitch,
              size_t planeBytes, uint8_t* __restrict__ planes, int itBegin, int itersPerBlock, int totalUnits) {
  __shared__ __align__(128) uint8_t s_out[RING * SLAB];
  __shared__ __align__(8) unsigned long long s_full[RING];  // mbarrier per slab: all WARPS warps have written it
  __shared__ __align__(16) uint32_t s_tab[4][32][4];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;

  // V's tap fragments stay in registers; H's four (one per fx) are re-read from shared memory at each use
  for (int i = tid; i < (int)(sizeof(d_tables.a1) / 16); i += THREADS)
    reinterpret_cast<uint4*>(&s_tab)[i] = reinterpret_cast<const uint4*>(&d_tables.a1)[i];
  const unsigned barBase = (unsigned)__cvta_generic_to_shared(s_full);
  if (tid < RING) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barBase + tid * 8), "r"(WARPS));
  __syncthreads();
  const unsigned tabBase = (unsigned)__cvta_generic_to_shared(&s_tab) + lane * 16;
  auto A1 = [&](int fx) {  // volatile: not hoisted out of the loop into 16 registers
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tabBase + fx * 512));
    return v;
  };
  uint4 A2[2][2];
#pragma unroll
  for (int sw = 0; sw < 2; ++sw)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) A2[sw][mt] = *reinterpret_cast<const uint4*>(d_tables.a2[sw][mt][lane]);

  // H for all four fx: T + 8192 (15 bits, positive) as (0.25 + low byte / 4096, high byte / 16) f16 pairs in V's
  // B-fragment order.  The four products are independent: issued back to back, converted afterwards.
  auto h_stage = [&](unsigned w0, unsigned w1, TTile& T) {
    int d[4][4];
#pragma unroll
    for (int fx = 0; fx < 4; ++fx) imma_c(d[fx], A1(fx), w0, w1, 0x54342000);
#pragma unroll
    for (int fx = 0; fx < 4; ++fx) {
      T.lo[fx][0] = __byte_perm(d[fx][0], d[fx][1], 0x6420); T.lo[fx][1] = __byte_perm(d[fx][2], d[fx][3], 0x6420);
      T.hi[fx][0] = hsub2u(__byte_perm(d[fx][0], d[fx][1], 0x7531), 0x54005400u);
      T.hi[fx][1] = hsub2u(__byte_perm(d[fx][2], d[fx][3], 0x7531), 0x54005400u);
    }
  };
  // One 8-row iteration: a new T tile from (w0, w1), 16 planes x 8 rows x 16 columns into the slab.  The two tiles
  // alternate roles; the B operand of V is always the register pair (ta, tb), and when ta holds the NEWER rows the
  // K halves of the tap matrix are exchanged instead (a second fragment set), so no tile is ever copied.
  // V of fx+1 is issued before the f32 -> u8 epilogue of fx, so that the tensor pipe and the FMA pipe overlap in a warp.
  auto iteration = [&](auto swapTag, TTile& ta, TTile& tb, unsigned w0, unsigned w1, uint8_t* so) {
    constexpr bool SWAP = decltype(swapTag)::value;
    h_stage(w0, w1, SWAP ? ta : tb);
    float acc[2][2][2][4];  // [fx parity][8-column half][m-tile][fragment]
    auto v_stage = [&](int fx) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {  // start: + 2048 for the rounding, - 64 * (8192 + 1024) for the operand offsets, all / 4096
          hmma_c(acc[fx & 1][nt][mt], A2[SWAP ? 1 : 0][mt], ta.hi[fx][nt], tb.hi[fx][nt], 0.5f - 144.0f);
          hmma(acc[fx & 1][nt][mt], A2[SWAP ? 1 : 0][mt], ta.lo[fx][nt], tb.lo[fx][nt]);
        }
    };
    auto epilogue = [&](int fx) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // fragment rows g (phase 2mt) and g+8 (phase 2mt+1)
          const float(&a0)[4] = acc[fx & 1][0][mt], (&a1)[4] = acc[fx & 1][1][mt];
          const unsigned z0 = floor_clip_u8(a0[2 * half]), z1 = floor_clip_u8(a0[2 * half + 1]);
          const unsigned z2 = floor_clip_u8(a1[2 * half]), z3 = floor_clip_u8(a1[2 * half + 1]);
          // 0x4b000000 << 8 vanishes mod 2^32: the low halves of z1*256+z0 and z3*256+z2 are the four bytes
          const unsigned word = __byte_perm(z1 * 256u + z0, z3 * 256u + z2, 0x5410);
          const int p = (2 * mt + half) * 4 + fx;
          *reinterpret_cast<unsigned*>(so + p * (8 * BW)) = word;
        }
    };
    v_stage(0);
#pragma unroll
    for (int fx = 1; fx < 4; ++fx) {
      v_stage(fx);
      epilogue(fx - 1);
    }
    epilogue(3);
  };

  // this CTA's share of the (column block, 8-row iteration) sequence, contiguous so that T tiles are reused
  int u = (int)((long long)blockIdx.x * totalUnits / gridDim.x);
  const int uEnd = (int)((long long)(blockIdx.x + 1) * totalUnits / gridDim.x);
  // slab: this lane writes row g, chunk w; in the store phase it moves chunk sc of row sr, planes sp0 + k * THREADS/64
  const int sOff = g * BW + (swz(w, g) << 4) + 4 * t;
  const int sc = tid % CHUNKS, sr = (tid / CHUNKS) & 7, sp0 = tid / (8 * CHUNKS);
  constexpr int PSTEP = THREADS / (8 * CHUNKS), NST = 16 / PSTEP;  // planes per pass, passes
  const uint8_t* sread = s_out + (sp0 * 8 + sr) * BW + (swz(sc, sr) << 4);

  // The warps of a CTA meet only through the slabs.  Slab k of the ring is "full" when all warps have arrived on its
  // mbarrier; a warp stores slab i-1 to the planes AFTER filtering iteration i, so it practically never waits, and the
  // warps drift apart by up to an iteration (their tensor / FMA / store phases interleave instead of coinciding).
  // Ring of 4: before writing slab i+4 (same buffer as i) a warp has waited for "full(i+2)", and every warp stores
  // slab i before it filters i+2.
  unsigned seq = 0;  // slabs produced so far by this CTA: buffer seq % RING, phase parity (seq / RING) & 1
  auto arrive = [&](unsigned s) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(barBase + (s % RING) * 8) : "memory");
  };
  auto wait_full = [&](unsigned s) {
    const unsigned bar = barBase + (s % RING) * 8, parity = (s / RING) & 1;
    unsigned done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
  };
  const unsigned planeStep = (unsigned)planeBytes * PSTEP;  // a slot (16 planes) is below 2^31 bytes: 32-bit offsets
  auto store_slab = [&](unsigned s, unsigned goff, bool ok) {
    wait_full(s);
    if (ok) {
      const uint8_t* src = sread + (s % RING) * SLAB;
#pragma unroll
      for (int k = 0; k < NST; ++k)
        *reinterpret_cast<uint4*>(planes + (goff + k * planeStep)) = *reinterpret_cast<const uint4*>(src + k * (PSTEP * 8 * BW));
    }
  };

  while (u < uEnd) {
    const int cb = u / itersPerBlock, it0 = u - cb * itersPerBlock;
    const int n = min(itersPerBlock - it0, uEnd - u);
    u += n;
    const int x0 = cb * BW + w * 16;       // padded-plane column of the strip
    const bool active = x0 < Wp;
    // this lane's 8 input bytes: picture columns px0 .. px0+7.  W, M and px0 are multiples of 8, so the group lies
    // inside the picture or entirely in the replicated margin: there it is 8 copies of the row's first / last sample,
    // taken from the nearest inside group with a per-lane byte selector.
    const int px0 = x0 - 8 - M + 8 * t;
    const int pxl = min(max(px0, 0), W - 8);
    const unsigned selA = px0 < 0 ? 0x0000u : px0 >= W ? 0x7777u : 0x3210u, selB = px0 < 0 ? 0x0000u : px0 >= W ? 0x7777u : 0x7654u;
    const uint8_t* colp = pic + pxl;
    int yIn = 8 * (itBegin + it0) - 3 + g - M;  // picture row of this lane's input row in T tile k (advances by 8)
    auto load_rows = [&](unsigned& a, unsigned& b) {  // raw words; the selectors are applied at use (fix_cols)
      const uint8_t* row = colp + (size_t)min(max(yIn, 0), H - 1) * picPitch;
      a = *reinterpret_cast<const unsigned*>(row);
      b = *reinterpret_cast<const unsigned*>(row + 4);
      yIn += 8;
    };
    auto fix_cols = [&](unsigned a, unsigned b, unsigned& w0, unsigned& w1) {
      w0 = __byte_perm(a, b, selA);
      w1 = __byte_perm(a, b, selB);
    };

    TTile ta, tb;
    unsigned w0 = 0, w1 = 0, n0 = 0, n1 = 0;
    if (active) {
      load_rows(n0, n1);
      fix_cols(n0, n1, w0, w1);
      load_rows(n0, n1);
      h_stage(w0, w1, ta);
    }
    int y = 8 * (itBegin + it0) + sr;
    unsigned goff = sp0 * (unsigned)planeBytes + y * pitch + cb * BW + sc * 16;
    const bool colOk = cb * BW + sc * 16 < Wp;
    for (int i = 0; i < n; ++i) {
      if (active) {
        fix_cols(n0, n1, w0, w1);
        load_rows(n0, n1);  // in flight during this iteration's filtering
        uint8_t* so = s_out + (seq % RING) * SLAB + sOff;
        if (i & 1) iteration(std::true_type{}, ta, tb, w0, w1, so); else iteration(std::false_type{}, ta, tb, w0, w1, so);
      }
      arrive(seq);
      if (i > 0) {
        store_slab(seq - 1, goff, colOk && y < Hp);
        goff += 8 * pitch;
        y += 8;
      }
      ++seq;
    }
    store_slab(seq - 1, goff, colOk && y < Hp);  // the segment's last slab
  }
}

}  // namespace k1m

// ---------------------------------------------------------------------------------------------------------------
// K1 with the vertical stage on the asynchronous tensor path (FME_K1_PATH_UMMA): the horizontal stage is the IMMA of
// k1_interp_mma, but the T tiles go to shared memory in the canonical K-major UMMA layout (a warp's B fragment IS one
// 8 x 8 core matrix: 128 contiguous bytes per STS.32) and V is tcgen05.mma kind::f16 issued by one thread:
//   D[M = 4 phases x 32 rows][N = 4 fx x 64 columns] = A[M][K = 48 T rows] x B[K][N],   K-steps of 16, hi and lo parts
// with the accumulator in TMEM (256 columns per CTA, two CTAs per SM).  No warp waits on an mma.sync any more, and the
// 32 HMMA per 16 x 8 step leave the issue stream.  TMEM lane = (phase, row): tcgen05.ld hands a thread 64 adjacent
// pixels of one plane row, converted with the same FFMA.SAT / FFMA.RM pair and staged through the swizzled slab.
namespace k1u {

using k1m::kTap;
using k1m::f16_of;
constexpr int THREADS = 256, BWU = 64;       // 8 warps: 4 strips of 16 columns x 2
constexpr int BR = 32;                       // output rows per block
constexpr int KBS = 8;                       // T ring: k-blocks of 8 rows
constexpr int NCOL = 4 * BWU;                // UMMA N
constexpr int T_KB_BYTES = NCOL * 16;        // one k-block: 256 n x (8 k x 2 B)
constexpr int T_BYTES = KBS * T_KB_BYTES;    // per half (hi / lo)
constexpr int A_CHUNK_BYTES = 128 * 16;      // one 8-wide K chunk of the tap matrix: 128 m x 16 B
constexpr int A_BYTES = 6 * A_CHUNK_BYTES;
constexpr int SLAB_BYTES = 16 * BR * BWU;
constexpr int SMEM_BYTES = 2 * T_BYTES + A_BYTES + SLAB_BYTES;
constexpr int TMEM_COLS = NCOL;

struct TapA { uint16_t v[6][16][8][8]; };    // [k chunk][m group][m % 8][k % 8], m = phase * 32 + row, T row k = row + tap
constexpr TapA make_tapA() {
  TapA t{};
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 48; ++k) {
      const int idx = k - (m % 32);
      t.v[k / 8][m / 8][m % 8][k % 8] = idx >= 0 && idx < 8 ? (uint16_t)f16_of(kTap[m / 32][idx], 0) : (uint16_t)0;
    }
  return t;
}
__device__ const TapA d_tapA = make_tapA();
static_assert(sizeof(TapA) == A_BYTES, "tap matrix size");

__device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lboBytes, unsigned sboBytes) {
  // SWIZZLE_NONE, K-major: 8 x 16-byte core matrices; LBO = distance between the two K chunks of a K = 16 step,
  // SBO = distance between 8-row groups along M / N; descriptor version 1 (bits 46-47)
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lboBytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sboBytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16(unsigned tmemD, uint64_t descA, uint64_t descB, unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                 "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                 "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// clip255(floor(D - 143.5)) in the low byte: D carries + 128 + 16 of the operand offsets, the rounding needs + 0.5
__device__ __forceinline__ unsigned floor_clip_u8(unsigned dBits) {
  float y, z;
  asm("fma.rn.sat.f32 %0, %1, 0f3B800000, 0fBF0F7FFC;" : "=f"(y) : "f"(__uint_as_float(dBits)));  // sat((D - 143.5) / 256 + 2^-22)
  asm("fma.rm.f32 %0, %1, 0f437FFFFF, 0f4B000000;" : "=f"(z) : "f"(y));
  return __float_as_uint(z);
}

__global__ void __launch_bounds__(THREADS, 2)
k1_interp_umma(const uint8_t* __restrict__ pic, int picPitch, int W, int H, int M, int Wp, int Hp, int pitch,
               size_t planeBytes, uint8_t* __restrict__ planes, int blkBegin, int blocksPerCol, int totalUnits) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* const sThi = smem;
  uint8_t* const sTlo = smem + T_BYTES;
  uint8_t* const sA = smem + 2 * T_BYTES;
  uint8_t* const sSlab = sA + A_BYTES;
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ unsigned s_tmem;
  // eight warps: warp w and w + 4 share strip (w & 3) in the horizontal stage (even / odd k-blocks) and TMEM lane quarter
  // (w & 3) = vertical phase in the epilogue (fx 0-1 / fx 2-3)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, ws = w & 3, wh = w >> 2, g = lane >> 2, t = lane & 3;

  for (int i = tid; i < A_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(&d_tapA)[i];
  uint4 A1[4];
#pragma unroll
  for (int fx = 0; fx < 4; ++fx) A1[fx] = *reinterpret_cast<const uint4*>(k1m::d_tables.a1n[fx][lane]);
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar);
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // no barrier here: the first block's barrier publishes the tap matrix, the mbarrier and the TMEM address
  const unsigned aAddr = (unsigned)__cvta_generic_to_shared(sA), hiAddr = (unsigned)__cvta_generic_to_shared(sThi),
                 loAddr = (unsigned)__cvta_generic_to_shared(sTlo);
  constexpr unsigned IDESC = (1u << 4) | ((unsigned)(NCOL >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major, M 128, N 256
  unsigned parity = 0;

  int u = (int)((long long)blockIdx.x * totalUnits / gridDim.x);
  const int uEnd = (int)((long long)(blockIdx.x + 1) * totalUnits / gridDim.x);
  while (u < uEnd) {
    const int cb = u / blocksPerCol, b0 = u - cb * blocksPerCol;
    const int n = min(blocksPerCol - b0, uEnd - u);
    u += n;
    const int x0 = cb * BWU + ws * 16;
    const int px0 = x0 - 8 - M + 8 * t;
    const int pxl = min(max(px0, 0), W - 8);
    const unsigned selA = px0 < 0 ? 0x0000u : px0 >= W ? 0x7777u : 0x3210u, selB = px0 < 0 ? 0x0000u : px0 >= W ? 0x7777u : 0x7654u;
    const uint8_t* colp = pic + pxl;
    const unsigned tOff = (unsigned)((2 * ws) * 128 + g * 16 + 4 * t);   // + fx * 1024 + nt * 128 + slot * T_KB_BYTES
    int kbNext = 4 * (blkBegin + b0);                 // next T k-block to produce: plane rows 8 kb - 3 .. 8 kb + 4
    for (int bi = 0; bi < n; ++bi) {
      const int b = blkBegin + b0 + bi;
      // horizontal stage: k-blocks kbNext .. 4b+5 (six for the first block of a segment, four afterwards), this warp the
      // ones of its parity; all their rows are requested before the first is used
      const int kbEnd = 4 * b + 6;
      const int kb0 = kbNext + ((kbNext ^ wh) & 1);   // first k-block of this warp's parity
      unsigned ra[3], rb[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int kb = kb0 + 2 * j;
        if (kb < kbEnd) {
          const uint8_t* row = colp + (size_t)min(max(8 * kb - 3 + g - M, 0), H - 1) * picPitch;
          ra[j] = *reinterpret_cast<const unsigned*>(row);
          rb[j] = *reinterpret_cast<const unsigned*>(row + 4);
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int kb = kb0 + 2 * j;
        if (kb < kbEnd) {
          const unsigned w0 = __byte_perm(ra[j], rb[j], selA), w1 = __byte_perm(ra[j], rb[j], selB);
          const unsigned slotOff = (unsigned)(kb & (KBS - 1)) * T_KB_BYTES + tOff;
#pragma unroll
          for (int fx = 0; fx < 4; ++fx) {
            int d[4];
            k1m::imma_c(d, A1[fx], w0, w1, 0x54342000);
            const unsigned o = slotOff + fx * 1024;
            *reinterpret_cast<unsigned*>(sTlo + o) = __byte_perm(d[0], d[1], 0x6420);
            *reinterpret_cast<unsigned*>(sTlo + o + 128) = __byte_perm(d[2], d[3], 0x6420);
            *reinterpret_cast<unsigned*>(sThi + o) = k1m::hsub2u(__byte_perm(d[0], d[1], 0x7531), 0x54005400u);
            *reinterpret_cast<unsigned*>(sThi + o + 128) = k1m::hsub2u(__byte_perm(d[2], d[3], 0x7531), 0x54005400u);
          }
        }
      }
      kbNext = kbEnd;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();   // T tiles of this block are in shared memory; the previous block's slab has been stored
      const unsigned tmem = s_tmem;
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int hl = 0; hl < 2; ++hl)
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            const unsigned slot = (unsigned)((4 * b + 2 * ks) & (KBS - 1));
            umma_f16(tmem, smem_desc(aAddr + 2 * ks * A_CHUNK_BYTES, A_CHUNK_BYTES, 128),
                     smem_desc((hl ? loAddr : hiAddr) + slot * T_KB_BYTES, T_KB_BYTES, 128), IDESC, (hl | ks) != 0);
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
      }
      {
        unsigned done = 0;
        for (int spin = 0; !done; ++spin) {
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                       : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
          if (spin > (1 << 22)) __trap();   // a lost completion must not hang the device
        }
        parity ^= 1;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // epilogue: TMEM lane = (phase ws, row `lane`); columns fx * 64 + x; this warp takes fx = 2 wh, 2 wh + 1
#pragma unroll 1
      for (int qq = 0; qq < 4; ++qq) {   // q = fx * 2 + half: 32 adjacent pixels of plane (ws, fx), row lane
        const int q = 4 * wh + qq;
        unsigned v[32];
        tmem_ld32(tmem + ((unsigned)(32 * ws) << 16) + q * 32, v);
        const int p = ws * 4 + (q >> 1);
        uint8_t* row = sSlab + (p * BR + lane) * BWU;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 o;
          unsigned* ow = reinterpret_cast<unsigned*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const unsigned* s4 = &v[c * 16 + k * 4];
            const unsigned z0 = floor_clip_u8(s4[0]), z1 = floor_clip_u8(s4[1]), z2 = floor_clip_u8(s4[2]), z3 = floor_clip_u8(s4[3]);
            ow[k] = __byte_perm(z1 * 256u + z0, z3 * 256u + z2, 0x5410);
          }
          const int chunk = (q & 1) * 2 + c;
          *reinterpret_cast<uint4*>(row + ((chunk ^ ((lane >> 1) & 3)) << 4)) = o;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();   // slab complete, TMEM drained
      {
        const int c = tid & 3, r = (tid >> 2) & 31, ph = tid >> 7;   // 256 threads: planes ph, ph + 2, ...
        const int y = BR * b + r, x = cb * BWU + c * 16;
        if (y < Hp && x < Wp) {
          const uint8_t* src = sSlab + (ph * BR + r) * BWU + ((c ^ ((r >> 1) & 3)) << 4);
          uint8_t* dst = planes + ph * planeBytes + (size_t)y * pitch + x;
          uint4 vv[8];
#pragma unroll
          for (int pl = 0; pl < 8; ++pl) vv[pl] = *reinterpret_cast<const uint4*>(src + pl * (2 * BR * BWU));
#pragma unroll
          for (int pl = 0; pl < 8; ++pl) *reinterpret_cast<uint4*>(dst + (size_t)(2 * pl) * planeBytes) = vv[pl];
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(TMEM_COLS));
  }
}

}  // namespace k1u

// Edge-replicating copy of a chroma picture into its padded plane (used by MC only).
__global__ void k_pad_plane(const uint8_t* __restrict__ pic, int picPitch, int W, int H, int M, int Wp, int Hp,
                            int pitch, uint8_t* __restrict__ dst) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= Wp || y >= Hp) return;
  int px = min(max(x - M, 0), W - 1), py = min(max(y - M, 0), H - 1);
  dst[(size_t)y * pitch + x] = pic[(size_t)py * picPitch + px];
}

// Pel (int16) picture -> u8 picture (8-bit content).
__global__ void k_pel_to_u8(const int16_t* __restrict__ src, int srcStride, uint8_t* __restrict__ dst, int dstPitch,
                            int w, int h) {
  int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int16_t* s = src + (size_t)y * srcStride + x;
  uint8_t* d = dst + (size_t)y * dstPitch + x;
  if (x + 3 < w) {
    unsigned v = (unsigned)(s[0] & 0xff) | ((unsigned)(s[1] & 0xff) << 8) | ((unsigned)(s[2] & 0xff) << 16) |
                 ((unsigned)(s[3] & 0xff) << 24);
    *reinterpret_cast<unsigned*>(d) = v;
  } else {
    for (int k = 0; x + k < w; ++k) d[k] = (uint8_t)s[k];
  }
}

}  // namespace

// rowBegin / rowEnd: padded-plane rows [rowBegin, rowEnd) to produce (whole 16-row tiles covering them); 0 / Hp = all
cudaError_t fme_launch_k1(const FmeGeom& g, const uint8_t* d_pic, int picPitch, uint8_t* d_planes, int* d_tileCounter,
                          int numSMs, int rowBegin, int rowEnd, int path, cudaStream_t s, int64_t* launches) {
  // FME_K1_PATH_AUTO: measured on B200 (tools/k1_probe.py, us per launch, dp4a | mma): 720p 16.4 | 12.5, 1080p 21.5 | 16.7,
  // 1440p 33.2 | 26.8, 2160p 61.2 | 65.2 -- the tensor path up to ~6.5 M padded samples per launch, dp4a above.
  const long long samples = (long long)g.Wp * (min(rowEnd, g.Hp) - max(rowBegin, 0));
  if (path == 0) path = samples <= 6500000ll ? 2 : 1;
  if (path == 3 && g.W % 8 == 0 && g.M % 8 == 0) {  // FME_K1_PATH_UMMA: (64-column block, 32-row block) units over two CTAs per SM
    const int bBegin = max(rowBegin, 0) / k1u::BR, bEnd = (min(rowEnd, g.Hp) + k1u::BR - 1) / k1u::BR;
    if (bEnd <= bBegin) return cudaSuccess;
    const int nb = bEnd - bBegin, units = ((g.Wp + k1u::BWU - 1) / k1u::BWU) * nb;
    cudaError_t e = cudaFuncSetAttribute(k1u::k1_interp_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, k1u::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const int grid = max(1, min(numSMs * 2, units));
    k1u::k1_interp_umma<<<grid, k1u::THREADS, k1u::SMEM_BYTES, s>>>(d_pic, picPitch, g.W, g.H, g.M, g.Wp, g.Hp, g.pitch,
                                                                   g.planeBytes, d_planes, bBegin, nb, units);
    ++*launches;
    return cudaGetLastError();
  }
  if (path == 2 && g.W % 8 == 0 && g.M % 8 == 0) {  // FME_K1_PATH_MMA (its input groups of 8 columns must not straddle the picture edge)
    const int itBegin = max(rowBegin, 0) / 8, itEnd = (min(rowEnd, g.Hp) + 7) / 8;
    if (itEnd <= itBegin) return cudaSuccess;
    const int iters = itEnd - itBegin, units = ((g.Wp + k1m::BW - 1) / k1m::BW) * iters;
    const int grid = max(1, min(numSMs * FME_K1M_CTAS, units / 4));
    k1m::k1_interp_mma<<<grid, k1m::THREADS, 0, s>>>(d_pic, picPitch, g.W, g.H, g.M, g.Wp, g.Hp, g.pitch, g.planeBytes,
                                                     d_planes, itBegin, iters, units);
    ++*launches;
    return cudaGetLastError();
  }
  const int tilesX = (g.Wp + TW - 1) / TW;
  const int tyBegin = max(rowBegin, 0) / TH, tyEnd = (min(rowEnd, g.Hp) + TH - 1) / TH;
  if (tyEnd <= tyBegin) return cudaSuccess;
  const int nTiles = tilesX * (tyEnd - tyBegin);
  const int grid = nTiles < numSMs * FME_K1_CTAS ? nTiles : numSMs * FME_K1_CTAS;
  // ceil(2^32 / tilesX): __umulhi(tile, rcp) == tile / tilesX for every tile < 2^32 / tilesX (tiles are < 2^17 at 8K)
  const unsigned tilesXRcp = (unsigned)((0x100000000ull + tilesX - 1) / tilesX);
  k1_interp_planes<<<grid, K1_THREADS, 0, s>>>(d_pic, picPitch, g.W, g.H, g.M, g.Wp, g.Hp, g.pitch, g.planeBytes,
                                                d_planes, tilesX, tilesXRcp, nTiles, d_tileCounter, tyBegin);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_pad_chroma(const FmeGeom& g, const uint8_t* d_pic, int picPitch, uint8_t* d_plane,
                                  cudaStream_t s, int64_t* launches) {
  dim3 grid((g.Wcp + 127) / 128, g.Hcp);
  k_pad_plane<<<grid, 128, 0, s>>>(d_pic, picPitch, g.Wc, g.Hc, g.Mc, g.Wcp, g.Hcp, g.cPitch, d_plane);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_pel_to_u8(const int16_t* d_src, int srcStride, uint8_t* d_dst, int dstPitch, int w, int h,
                                 cudaStream_t s, int64_t* launches) {
  dim3 grid(((w + 3) / 4 + 127) / 128, h);
  k_pel_to_u8<<<grid, 128, 0, s>>>(d_src, srcStride, d_dst, dstPitch, w, h);
  ++*launches;
  return cudaGetLastError();
}
