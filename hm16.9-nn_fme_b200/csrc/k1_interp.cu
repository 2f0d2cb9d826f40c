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
    const int tr = (int)__umulhi((unsigned)tile, tilesXRcp), tx = tile - tr * tilesX;  // tilesXRcp = ceil(2^32 / tilesX)
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
    const int trI = (int)__umulhi((unsigned)tile, tilesXRcp), txI = tile - trI * tilesX;
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
                          int numSMs, int rowBegin, int rowEnd, cudaStream_t s, int64_t* launches) {
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
