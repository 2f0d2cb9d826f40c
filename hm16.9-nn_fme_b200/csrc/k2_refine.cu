// K2: batched half-/quarter-pel refinement with SATD or SAD + MV-bit cost
// (replaces TEncSearch::xPatternRefinement TEncSearch.cpp:1591-1645 and the distortion functions it
// calls, TComRdCost::xGetHADs / xCalcHADs8x8 / xCalcHADs4x4 TComRdCost.cpp:1234-1495, xGetSAD* :359-855,
// driven as in xPatternSearchFracDIF TEncSearch.cpp:5232-5269).
//
// Candidate addressing (derived from TEncSearch.cpp:1613-1623, 6343-6531): with
// (X,Y) = PU origin + integer MV and total quarter-pel offset (qx,qy) in [-3,3]^2 the candidate block is
// plane P[qy&3][qx&3] at integer origin (X + (qx>>2), Y + (qy>>2)).
//
// Work decomposition:
//   * a prepass groups PU indices by shape class (w,h) so that a warp always works on same-shape PUs and all warps
//     of an SM run the same k2_pack instantiation (class-major pack order: mixing classes thrashes the i-cache);
//   * one warp = one "pack" of P PUs whose tiles (8x8, or 4x4 when w or h is not a multiple of 8,
//     TComRdCost.cpp:1446-1492) fill the 32 lanes: lane = (PU in pack, tile in PU);
//   * the warp walks 12 staging steps per pack (4 half-pel planes, then one plane per quarter-pel candidate):
//     the candidate regions of all its PUs are copied from the HBM/L2-resident planes into shared memory
//     with 4/8/16-byte cp.async into one of two buffers while the previous step's candidates are evaluated
//     from the other (rows stay source-aligned, readers funnel-shift; each PU's region is copied by the lanes that
//     serve that PU, one pointer bump per row); every lane
//     computes the SATD of its own tile in registers: two 16-bit residuals per 32-bit register (SWAR),
//     butterflies as plain 32-bit adds, |a+b|+|a-b| = 2 max(|a|,|b|) for the intra-register stage;
//   * per-PU sums by xor-shuffles over the PU's power-of-two lane group, MV cost from an exact host-built LUT,
//     first-minimum argmin in the reference's table order (TEncSearch.cpp:212-236, strict < at :1634).
#include "k2_common.cuh"
#include "k3_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// prepass: counting sort of PU indices by shape class
// ------------------------------------------------------------------------------------------------
// Records are binned in two passes: uni-prediction records (wantBi = 0) and, when the ctx enables it, bi-predictive
// refinement records (FME_PU_BI, wantBi = 1) -- a pack never mixes the two kinds (different k2_pack instantiations).
__device__ __forceinline__ int k2_class_of(const fme_pu& p, int wantBi) {
  if (((p.flags & FME_PU_BI) != 0) != (wantBi != 0)) return -1;
  if (!fme_hevc_pu_shape(p.w, p.h)) return -1;
  return fme_dim_index(p.w) * 8 + fme_dim_index(p.h);
}
// binning key: (reference slot group, shape class); the schedule position of a key is key ^ 63 (slot-major, large
// shapes first inside a slot: a 64x64 PU is a whole double-length pack, the cheap 4x8 / 8x4 packs end the slot)
__device__ __forceinline__ int k2_key_of(const fme_pu& p, int wantBi) {
  const int c = k2_class_of(p, wantBi);
  return c < 0 ? -1 : ((p.refSlot & (FME_K2_SLOT_GROUPS - 1)) << 6) | c;
}

// res != nullptr (the uni-prediction pass): records no K2 pass will serve -- a shape HEVC cannot produce, or FME_PU_BI
// on a ctx without biPred -- get the sentinel result (zero vectors, cost 0xffffffff) instead of whatever the result
// buffer held (the unvalidated async / device entry points, include/fme_b200.h).
// errOnGpuCount != nullptr (the uni-prediction pass): += the number of records flagged FME_PU_ERR_ON_GPU, so that a K0 pass
// launched behind K2 on the unvalidated path can return at once when no record asks for it.
__global__ void k2_count(const fme_pu* __restrict__ pus, int n, int* __restrict__ classCount, int wantBi,
                         fme_result* __restrict__ res, int biServed, short* __restrict__ keys, int* __restrict__ errOnGpuCount) {
  __shared__ int s_cnt[FME_K2_KEYS];
  __shared__ int s_flagged;
  for (int i = threadIdx.x; i < FME_K2_KEYS; i += blockDim.x) s_cnt[i] = 0;
  if (threadIdx.x == 0) s_flagged = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const fme_pu p = pus[i];
    const int c = k2_key_of(p, wantBi);
    keys[i] = (short)c;
    if (errOnGpuCount && (p.flags & FME_PU_ERR_ON_GPU)) atomicAdd(&s_flagged, 1);
    // neighbouring records mostly share their key (raster CU order): one shared-memory atomic per distinct key per warp
    const unsigned peers = __match_any_sync(__activemask(), c);
    if (c >= 0) {
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[c], __popc(peers));
    } else if (res && (!fme_hevc_pu_shape(p.w, p.h) || ((p.flags & FME_PU_BI) && !biServed)))
      *reinterpret_cast<uint2*>(&res[i]) = make_uint2(0u, 0xffffffffu);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FME_K2_KEYS; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&classCount[i], s_cnt[i]);
  if (errOnGpuCount && threadIdx.x == 0 && s_flagged) atomicAdd(errOnGpuCount, s_flagged);
}

// Scatter PU indices into schedule order.  Every block derives the key and pack offsets itself from the key counts
// (one warp: 16 schedule positions per lane + a warp scan, cheaper than a separate launch); block 0 also publishes
// them for k2_refine.
__global__ void k2_scatter(const short* __restrict__ keys, int n, const int* __restrict__ classCount,
                           int* __restrict__ classOffset, int* __restrict__ packOffset,
                           int* __restrict__ classCursor, int* __restrict__ order, int mmaGroups) {
  __shared__ int s_cnt[FME_K2_KEYS];
  __shared__ int s_base[FME_K2_KEYS];
  __shared__ int s_classOff[FME_K2_KEYS];
  constexpr int PER = FME_K2_KEYS / 32;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int cnt[PER], pk[PER], sc = 0, sp = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int v = lane * PER + j;  // schedule position; key = v ^ 63, class = 63 - (v & 63)
      cnt[j] = classCount[v ^ 63];
      const int P = pack_pus(class_info(63 - (v & 63)), mmaGroups);
      pk[j] = cnt[j] ? (cnt[j] + P - 1) / P : 0;
      sc += cnt[j];
      sp += pk[j];
    }
    int ec = sc, ep = sp;  // inclusive warp scan of the lane totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int tc = __shfl_up_sync(0xffffffffu, ec, d), tp = __shfl_up_sync(0xffffffffu, ep, d);
      if (lane >= d) { ec += tc; ep += tp; }
    }
    int oc = ec - sc, op = ep - sp;  // exclusive offsets of this lane's first position
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int v = lane * PER + j;
      s_classOff[v] = oc;
      if (blockIdx.x == 0) { classOffset[v] = oc; packOffset[v] = op; }
      oc += cnt[j];
      op += pk[j];
    }
    if (blockIdx.x == 0 && lane == 31) { classOffset[FME_K2_KEYS] = oc; packOffset[FME_K2_KEYS] = op; }
  }
  // one contiguous chunk per block so that local ranks are well defined
  int chunk = (n + gridDim.x - 1) / gridDim.x;
  int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  for (int i = threadIdx.x; i < FME_K2_KEYS; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int c = keys[i];
    const unsigned peers = __match_any_sync(__activemask(), c);
    if (c >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_cnt[c], __popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FME_K2_KEYS; i += blockDim.x) {
    s_base[i] = s_cnt[i] ? s_classOff[i ^ 63] + atomicAdd(&classCursor[i], s_cnt[i]) : 0;
    s_cnt[i] = 0;
  }
  __syncthreads();
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int c = keys[i];
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, c);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (c >= 0 && lane == leader) base = atomicAdd(&s_cnt[c], __popc(peers));
    base = __shfl_sync(act, base, leader);
    if (c >= 0) order[s_base[c] + base + __popc(peers & ((1u << lane) - 1u))] = i;
  }
}

// ------------------------------------------------------------------------------------------------
// tile distortion in registers
// ------------------------------------------------------------------------------------------------
// x264-style packed |.| on two 16-bit lanes held carry-tolerantly in one 32-bit word
// (value = hi*65536 + lo with signed lo); returns true unsigned fields (|hi| , |lo|).
__device__ __forceinline__ unsigned abs2(unsigned a) {
  // s = 0xffff in every 16-bit field whose raw sign bit is set: one PRMT in sign-replicate mode (selector nibbles
  // 9 / b replicate the msb of byte 1 / byte 3) instead of shift + mask + multiply
  // (inline PTX: the __byte_perm intrinsic drops bit 3 of each selector nibble, i.e. the sign-replicate mode)
  unsigned s;
  asm("prmt.b32 %0, %1, %2, 0xbb99;" : "=r"(s) : "r"(a), "r"(0u));
  return (a + s) ^ s;
}
// acc + max(A.lo, A.hi) + max(B.lo, B.hi) for two words of unsigned 16-bit fields: regroup (A.hi, B.hi) and
// (A.lo, B.lo) with two PRMTs, one packed max, one dp2a with multipliers (1, 1) sums both fields into 32 bits
__device__ __forceinline__ unsigned acc_hmax2x2(unsigned A, unsigned B, unsigned acc) {
  unsigned mx = __vmaxu2(__byte_perm(A, B, 0x7632), __byte_perm(A, B, 0x5410));
  unsigned d;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(mx), "r"(0x0101u), "r"(acc));
  return d;
}

// In-place 4-point Hadamard of (a, b, c, d) on SWAR words: two 2-input and four 3-input adds.
__device__ __forceinline__ void had4(unsigned& a, unsigned& b, unsigned& c, unsigned& d) {
  const unsigned t = a + b, u = a - b;
  const unsigned y0 = t + c + d, y2 = t - c - d, y1 = u + c - d, y3 = u - c + d;
  a = y0; b = y1; c = y2; d = y3;
}


// Transform + sum of an 8x8 residual held as d[4r + j] = (res(r,2j), res(r,2j+1)) packed lo/hi.
__device__ __forceinline__ unsigned satd8x8_tail(unsigned (&d)[32]) {
  // horizontal: column-index bits 1 and 2 (bit 0 lives inside a word and is folded into the final max)
  // (radix-4: 6 three-input adds per 4 points instead of 8 two-input ones; IADD3 takes negated operands)
#pragma unroll
  for (int r = 0; r < 8; ++r) had4(d[4 * r], d[4 * r + 1], d[4 * r + 2], d[4 * r + 3]);
  // vertical: three stages over the row index
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unsigned v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = d[4 * r + j];
    had4(v[0], v[1], v[2], v[3]);
    had4(v[4], v[5], v[6], v[7]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned a = v[k], b = v[k + 4];
      v[k] = a + b;
      v[k + 4] = a - b;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) d[4 * r + j] = v[r];
  }
  // last horizontal stage + abs: |lo+hi| + |lo-hi| = 2 max(|lo|,|hi|)
  unsigned sum = 0;
#pragma unroll
  for (int i = 0; i < 32; i += 2) sum = acc_hmax2x2(abs2(d[i]), abs2(d[i + 1]), sum);
  return (2 * sum + 2) >> 2;  // TComRdCost.cpp:1421
}

// SATD of one 8x8 tile: xCalcHADs8x8 (TComRdCost.cpp:1330-1425).  o[] holds the source tile as
// 16 words of u8x4 (row r -> o[2r], o[2r+1]); cand points at the candidate tile's row 0 in smem.
template <typename OrgRow>
__device__ __forceinline__ unsigned satd8x8(OrgRow orgRow, const uint8_t* cand, int candPitch) {
  unsigned d[32];  // d[4r + j] = (res(r,2j), res(r,2j+1)) packed lo/hi
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned c0, c1, o0, o1;
    cand_row8(ct, r, c0, c1);
    orgRow(r, o0, o1);
    d[4 * r + 0] = __byte_perm(o0, 0, 0x4140) - __byte_perm(c0, 0, 0x4140);
    d[4 * r + 1] = __byte_perm(o0, 0, 0x4342) - __byte_perm(c0, 0, 0x4342);
    d[4 * r + 2] = __byte_perm(o1, 0, 0x4140) - __byte_perm(c1, 0, 0x4140);
    d[4 * r + 3] = __byte_perm(o1, 0, 0x4342) - __byte_perm(c1, 0, 0x4342);
  }
  return satd8x8_tail(d);
}

__device__ __forceinline__ unsigned satd4x4_tail(unsigned (&d)[8]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    unsigned a = d[2 * r], b = d[2 * r + 1];
    d[2 * r] = a + b;
    d[2 * r + 1] = a - b;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) had4(d[j], d[2 + j], d[4 + j], d[6 + j]);
  unsigned sum = 0;
#pragma unroll
  for (int i = 0; i < 8; i += 2) sum = acc_hmax2x2(abs2(d[i]), abs2(d[i + 1]), sum);
  return (2 * sum + 1) >> 1;  // TComRdCost.cpp:1325
}

// SATD of one 4x4 tile: xCalcHADs4x4 (TComRdCost.cpp:1234-1328).  o[r] = source row r (u8x4).
__device__ __forceinline__ unsigned satd4x4(const unsigned (&o)[4], const uint8_t* cand, int candPitch) {
  unsigned d[8];
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    unsigned c = cand_row4(ct, r);
    d[2 * r + 0] = __byte_perm(o[r], 0, 0x4140) - __byte_perm(c, 0, 0x4140);
    d[2 * r + 1] = __byte_perm(o[r], 0, 0x4342) - __byte_perm(c, 0, 0x4342);
  }
  return satd4x4_tail(d);
}



// ------------------------------------------------------------------------------------------------
// 8x8 SATD on the tensor pipe (fp16 in, fp32 accumulate -- the variant BASELINE.json's north_star names)
// ------------------------------------------------------------------------------------------------
// xCalcHADs8x8 (TComRdCost.cpp:1330-1425) is sum |H8 D H8| = sum |(H8 (x) H8) vec(D)|, a 64x64 +-1 matrix applied to the
// 64 residuals of a tile.  One warp evaluates eight tiles at a time ("column block"): lane (g, t), g = lane >> 2,
// t = lane & 3, feeds rows 2t and 2t+1 of tile g.  With pixel (y, x), y = 2t + ks1, x = 4 ks0 + 2 c1 + c0 and
// coefficient (u, v), u = 2 uh + u0, v = 4 v2 + 2 v1 + v0, the sign (-1)^(popc(u&y) + popc(v&x)) splits into
//   (-1)^(u0 ks1 + v2 ks0)  *  (-1)^(popc(uh & t) + v1 c1 + v0 c0):
// the first factor is a 2x2 butterfly over the four 4-pixel groups (ks1, ks0) a lane holds -- done on packed f16x2
// (values <= 4 * 255, exact) -- and the second is the natural 16x16 Hadamard matrix over k = 8 c1 + 2 t + c0, which is
// exactly how mma.m16n8k16 spreads its K index over the lanes.  So 16 HADD2 + 4 HMMA (one constant A fragment)
// replace the 64x64 product; all values are integers below 2^24, so fp32 accumulation is exact.
// u8 -> f16 is one PRMT per pixel pair: byte b under the high byte 0x64 is the f16 number 1024 + b.
__device__ __forceinline__ unsigned f16pair_lo(unsigned w) {
  unsigned r;
  asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(r) : "r"(w), "r"(0x64646464u));
  return r;
}
__device__ __forceinline__ unsigned f16pair_hi(unsigned w) {
  unsigned r;
  asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(r) : "r"(w), "r"(0x64646464u));
  return r;
}
__device__ __forceinline__ unsigned hadd2u(unsigned a, unsigned b) {
  unsigned r;
  asm("add.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ unsigned hsub2u(unsigned a, unsigned b) {
  unsigned r;
  asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ unsigned lds_u32(unsigned sa) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa));
  return v;
}
__device__ __forceinline__ void hmma16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// A fragment of the natural 16x16 Hadamard matrix, A[row][k] = (-1)^popc(row & k) as f16 +-1
// (m16n8k16 .row layout: a0,a1 = (g; 2t, 2t+1)  a2,a3 = (g+8; 2t, 2t+1)  a4,a5 = (g; 2t+8, 2t+9)  a6,a7 = (g+8; 2t+8, 2t+9))
__device__ __forceinline__ void had16_afrag(int lane, unsigned (&a)[4]) {
  const int g = lane >> 2, t = lane & 3;
  auto one = [](int row, int k) { return (__popc(row & k) & 1) ? 0xbc00u : 0x3c00u; };
  auto pair = [&](int row, int k) { return one(row, k) | (one(row, k + 1) << 16); };
  a[0] = pair(g, 2 * t);
  a[1] = pair(g + 8, 2 * t);
  a[2] = pair(g, 2 * t + 8);
  a[3] = pair(g + 8, 2 * t + 8);
}
// The eight f16x2 words of a lane's two tile rows: f[2 ks + r], ks = 2 ks1 + ks0 (row 2t + ks1, columns 4 ks0 ..),
// r = 0: columns +0,+1 (k = 2t, 2t+1), r = 1: columns +2,+3 (k = 2t+8, 2t+9).
__device__ __forceinline__ void f16_rows(unsigned lo0, unsigned hi0, unsigned lo1, unsigned hi1, unsigned (&f)[8]) {
  f[0] = f16pair_lo(lo0); f[1] = f16pair_hi(lo0);
  f[2] = f16pair_lo(hi0); f[3] = f16pair_hi(hi0);
  f[4] = f16pair_lo(lo1); f[5] = f16pair_hi(lo1);
  f[6] = f16pair_lo(hi1); f[7] = f16pair_hi(hi1);
}
// Residual rows of one column block -> the four B fragments of its 2x2-butterflied pixel groups (f[2 mt], f[2 mt + 1]).
__device__ __forceinline__ void cand_bfrags(unsigned cAddr, int RW, const unsigned (&o)[8], int lane, unsigned (&f)[8]) {
  const unsigned a0 = cAddr + 2 * (lane & 3) * RW;
  const unsigned base = a0 & ~3u, sh = (a0 & 3u) * 8u;
  const unsigned w0 = lds_u32(base), w1 = lds_u32(base + 4), w2 = lds_u32(base + 8);
  const unsigned x0 = lds_u32(base + RW), x1 = lds_u32(base + RW + 4), x2 = lds_u32(base + RW + 8);
  f16_rows(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(x0, x1, sh),
           __funnelshift_r(x1, x2, sh), f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = hsub2u(f[i], o[i]);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const unsigned s0 = hadd2u(f[0 + r], f[2 + r]), d0 = hsub2u(f[0 + r], f[2 + r]);  // over ks0, ks1 = 0
    const unsigned s1 = hadd2u(f[4 + r], f[6 + r]), d1 = hsub2u(f[4 + r], f[6 + r]);  // over ks0, ks1 = 1
    f[0 + r] = hadd2u(s0, s1); f[2 + r] = hadd2u(d0, d1);
    f[4 + r] = hsub2u(s0, s1); f[6 + r] = hsub2u(d0, d1);
  }
}
// sum |coefficients| of one column block: acc0 = this lane's share of tile 2t, acc1 of tile 2t+1
__device__ __forceinline__ void had_abs_sums(const unsigned (&f)[8], const unsigned (&afrag)[4], float& acc0, float& acc1) {
  float d[4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) hmma16816(d[mt], afrag, f[2 * mt], f[2 * mt + 1]);
  acc0 = (fabsf(d[0][0]) + fabsf(d[0][2])) + (fabsf(d[1][0]) + fabsf(d[1][2])) +
         ((fabsf(d[2][0]) + fabsf(d[2][2])) + (fabsf(d[3][0]) + fabsf(d[3][2])));
  acc1 = (fabsf(d[0][1]) + fabsf(d[0][3])) + (fabsf(d[1][1]) + fabsf(d[1][3])) +
         ((fabsf(d[2][1]) + fabsf(d[2][3])) + (fabsf(d[3][1]) + fabsf(d[3][3])));
}
// SATD of the 32 tiles of a pack (four column blocks).  cOwn: shared-window address of row 0 of the candidate tile
// this lane OWNS; orow(cb, o): source rows 2t, 2t+1 (f16 pairs) of the tile this lane FEEDS in block cb (tile
// 8 cb + g).  The eight per-lane partial sums (4 blocks x 2 columns) are reduced over the eight lanes g with a
// transposing butterfly, which leaves lane (g, t) with tile 8 (g >> 1) + 2 t + (g & 1); one more shuffle hands every
// lane the SATD of its own tile, rounded as TComRdCost.cpp:1421.
template <typename OrgRows>
__device__ __forceinline__ unsigned satd8x8_pack_mma(unsigned cOwn, int RW, OrgRows orow, const unsigned (&afrag)[4],
                                                     int lane, int gatherLane) {
  float a0[4], a1[4];
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    const unsigned ca = __shfl_sync(0xffffffffu, cOwn, 8 * cb + (lane >> 2));
    unsigned f[8];
    cand_bfrags(ca, RW, orow(cb), lane, f);
    had_abs_sums(f, afrag, a0[cb], a1[cb]);
  }
  const bool g0 = (lane & 4) != 0, g1 = (lane & 8) != 0, g2 = (lane & 16) != 0;
  float w[4];
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {  // lanes with g0 = j keep column j
    const float keep = g0 ? a1[cb] : a0[cb], send = g0 ? a0[cb] : a1[cb];
    w[cb] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float x[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {     // lanes with g1 = i keep block 2k + i
    const float keep = g1 ? w[2 * k + 1] : w[2 * k], send = g1 ? w[2 * k] : w[2 * k + 1];
    x[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float keep = g2 ? x[1] : x[0], send = g2 ? x[0] : x[1];
  const float z = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  // exact integer < 2^23: read it from the mantissa of z + 2^23
  const unsigned v = (unsigned)__float_as_int(z + 8388608.0f) - 0x4b000000u;
  return __shfl_sync(0xffffffffu, (v + 2) >> 2, gatherLane);
}


// The same transforms for a source block that is already held as carry-tolerant 16-bit pairs (the pattern
// 2*org - pred of bi-predictive refinement, samples in [-255, 510]): ou[4r + j] = (pat(r,2j), pat(r,2j+1)) for an 8x8
// tile, ou[2r + j] for a 4x4 tile.  Residuals stay within +-510, the 8x8 coefficients within +-32640: int16 fields.
__device__ __forceinline__ unsigned satd8x8_p(const unsigned* ou, const uint8_t* cand, int candPitch) {
  unsigned d[32];
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned c0, c1;
    cand_row8(ct, r, c0, c1);
    d[4 * r + 0] = ou[4 * r + 0] - __byte_perm(c0, 0, 0x4140);
    d[4 * r + 1] = ou[4 * r + 1] - __byte_perm(c0, 0, 0x4342);
    d[4 * r + 2] = ou[4 * r + 2] - __byte_perm(c1, 0, 0x4140);
    d[4 * r + 3] = ou[4 * r + 3] - __byte_perm(c1, 0, 0x4342);
  }
  return satd8x8_tail(d);
}
__device__ __forceinline__ unsigned satd4x4_p(const unsigned* ou, const uint8_t* cand, int candPitch) {
  unsigned d[8];
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const unsigned c = cand_row4(ct, r);
    d[2 * r + 0] = ou[2 * r + 0] - __byte_perm(c, 0, 0x4140);
    d[2 * r + 1] = ou[2 * r + 1] - __byte_perm(c, 0, 0x4342);
  }
  return satd4x4_tail(d);
}
// SAD on 16-bit pairs: packed |.| then a dp2a with multipliers (1, 1)
__device__ __forceinline__ unsigned acc_abs2(unsigned d, unsigned acc) {
  unsigned r;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(abs2(d)), "r"(0x0101u), "r"(acc));
  return r;
}
__device__ __forceinline__ unsigned sad8x8_p(const unsigned* ou, const uint8_t* cand, int candPitch) {
  unsigned s = 0;
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned c0, c1;
    cand_row8(ct, r, c0, c1);
    s = acc_abs2(ou[4 * r + 0] - __byte_perm(c0, 0, 0x4140), s);
    s = acc_abs2(ou[4 * r + 1] - __byte_perm(c0, 0, 0x4342), s);
    s = acc_abs2(ou[4 * r + 2] - __byte_perm(c1, 0, 0x4140), s);
    s = acc_abs2(ou[4 * r + 3] - __byte_perm(c1, 0, 0x4342), s);
  }
  return s;
}
__device__ __forceinline__ unsigned sad4x4_p(const unsigned* ou, const uint8_t* cand, int candPitch) {
  unsigned s = 0;
  const CandTile ct = cand_tile(cand, candPitch);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const unsigned c = cand_row4(ct, r);
    s = acc_abs2(ou[2 * r + 0] - __byte_perm(c, 0, 0x4140), s);
    s = acc_abs2(ou[2 * r + 1] - __byte_perm(c, 0, 0x4342), s);
  }
  return s;
}

// Distortion of one lane unit against the candidate at `cand` (row 0 of the unit's first tile).
template <int TS, bool BI = false, bool SWAR8 = true>
__device__ __forceinline__ unsigned unit_dist(const unsigned* o, const uint8_t* cand, int candPitch, int tile1Off,
                                              bool had) {
  if constexpr (BI && TS == 8) {
    return had ? satd8x8_p(o, cand, candPitch) : sad8x8_p(o, cand, candPitch);
  } else if constexpr (BI) {
    return had ? satd4x4_p(o, cand, candPitch) + satd4x4_p(o + 8, cand + tile1Off, candPitch)
               : sad4x4_p(o, cand, candPitch) + sad4x4_p(o + 8, cand + tile1Off, candPitch);
  } else if constexpr (TS == 8) {
    auto row = [&](int r, unsigned& lo, unsigned& hi) { lo = o[2 * r]; hi = o[2 * r + 1]; };
    if constexpr (SWAR8) return had ? satd8x8(row, cand, candPitch) : sad8x8(row, cand, candPitch);
    else return sad8x8(row, cand, candPitch);  // the 8x8 SATD of these packs runs on the tensor pipe (k2_pack)
  } else {
    const unsigned(&oa)[4] = *reinterpret_cast<const unsigned(*)[4]>(o);
    const unsigned(&ob)[4] = *reinterpret_cast<const unsigned(*)[4]>(o + 4);
    return had ? satd4x4(oa, cand, candPitch) + satd4x4(ob, cand + tile1Off, candPitch)
               : sad4x4(oa, cand, candPitch) + sad4x4(ob, cand + tile1Off, candPitch);
  }
}

// BI = bi-predictive refinement records (FME_PU_BI): the source block is the pattern 2*org - P_other, P_other being
// the other list's uni-prediction = plane P[mvY&3][mvX&3] of slot err[0] at the integer part of the MV packed in err[1]
// (TEncSearch.cpp:4462-4472, TComYuv::removeHighFreq without clipping).  Everything else is the same search.
// PATH selects how the 8x8 SATD of uni-prediction Hadamard packs is computed: 0 = SWAR integer (registers),
// 1 / 2 = tensor pipe (satd8x8_pack_mma; with PATH 2 only the groups k2_group_mma hands back land here).
template <int TS, int A, bool BI, int PATH>
__device__ __noinline__ void k2_pack(const ClassInfo ci, const int* __restrict__ order, int first, int count,
                                     const fme_pu* __restrict__ pus, fme_result* __restrict__ res,
                                     const uint8_t* __restrict__ planes, const uint8_t* __restrict__ org,
                                     const FmeGeom& g, const uint32_t* __restrict__ costLut, int useHad,
                                     uint8_t* smem) {
  const int lane = threadIdx.x & 31;
  const int w = ci.w, h = ci.h;
  const int U = ci.units;                      // lane units per PU
  // PUs with more than 32 units (64x64, 64x48, 48x64: one PU per pack, 8x8 tiles, 16-byte granules): every lane
  // also serves unit lane + 32, whose source tile is parked in shared memory
  const bool twoUnits = TS == 8 && A == 16 && U > 32;
  [[maybe_unused]] long long otherOff = 0;     // BI: byte offset of the other list's prediction block in the plane set
  const int gPitch = g.pitch, gOrgPitch = g.orgPitch;
  const int gPlaneBytes = (int)g.planeBytes;   // < 2^31 (padded 8K plane: 4480 * 7936)
  const int lanesPerPu = ci.lanes;             // power of two
  const int myPu = lane / lanesPerPu;          // PU slot in the pack served by this lane
  const bool laneActive = myPu < count;
  const int unit0 = lane - myPu * lanesPerPu;  // unit index in round 0 (may be >= U: idle lane of a padded group)

  // ---- this lane's PU ----
  int predX = 0, predY = 0, mvIntX = 0, mvIntY = 0, lossless = 0, puIdx = -1, alignX = 0;
  int ox = 0, oy = 0;
  long long slotOff = 0;  // byte offset of sample (X, Y) of the slot's plane 0; steps add plane * planeBytes and (dx, dy)
  if (laneActive) {
    puIdx = order[first + myPu];
    const fme_pu p = pus[puIdx];
    predX = p.mvPredX; predY = p.mvPredY; mvIntX = p.mvIntX; mvIntY = p.mvIntY;
    lossless = p.flags & FME_PU_LOSSLESS;
    // The reference clips MVs so that reads stay inside its padded planes (TComDataCU.cpp:2773-2786:
    // X in [-71, W+7]); the clamp below only guards device memory against out-of-contract records.
    const int X = min(max(p.x + p.mvIntX, -(g.M - 8)), g.W + g.M - 8 - w);
    const int Y = min(max(p.y + p.mvIntY, -(g.M - 8)), g.H + g.M - 8 - h);
    alignX = X + g.M;
    slotOff = (long long)min((int)p.refSlot, g.numSlots - 1) * (long long)g.slotBytes +
              (long long)((Y + g.M) * gPitch + alignX);
    ox = min(max((int)p.x, 0), g.W - w);
    oy = min(max((int)p.y, 0), g.H - h);
    if constexpr (BI) {
      const int oslot = min((int)(p.err[0] & 0xff), g.numSlots - 1);
      const int omx = (int)(short)(p.err[1] & 0xffff), omy = (int)(short)(p.err[1] >> 16);
      const int X2 = min(max(p.x + (omx >> 2), -(g.M - 8)), g.W + g.M - 8 - w);
      const int Y2 = min(max(p.y + (omy >> 2), -(g.M - 8)), g.H + g.M - 8 - h);
      otherOff = (long long)oslot * (long long)g.slotBytes +
                 (long long)(((omy & 3) * 4 + (omx & 3)) * gPlaneBytes + (Y2 + g.M) * gPitch + X2 + g.M);
    }
  }

  // ---- staging geometry ----
  StageGeom sg;
  sg.RW = ((w + A + A - 1) / A) * A;  // covers [x0 & ~(A-1), x0 + w + 1)
  sg.G = sg.RW / A;
  sg.RB = (h + 1) * sg.RW;
  if (((sg.RB / A) & 1) == 0) sg.RB += A;  // odd region stride in granules: lanes of different PUs spread over the banks
  const int bufBytes = (ci.P * sg.RB + 16 + 15) & ~15;
  uint8_t* const bufA = smem;
  uint8_t* const bufB = smem + bufBytes;
  // lane -> (granule column, first row, row step) inside its PU's region
  int gpShift = 0;
  while ((1 << gpShift) < sg.G && (1 << gpShift) < lanesPerPu) ++gpShift;
  const int Gp = 1 << gpShift;
  const int stGi = unit0 & (Gp - 1), stRowSub = unit0 >> gpShift, stRowStep = lanesPerPu >> gpShift;
  const bool stSecond = Gp < sg.G;                   // only single-lane groups: G == 2
  const bool stOn = laneActive && stGi < sg.G;
  const unsigned stDst = (unsigned)__cvta_generic_to_shared(smem) + myPu * sg.RB + stRowSub * sg.RW + stGi * A;
  const int stSrcOff = stRowSub * gPitch + stGi * A;

  // half-pel winner of this lane's PU (identical in all lanes of the PU after the per-PU sums)
  int bhx = 0, bhy = 0;
  // Region of step s.  Steps 0..3: half-pel planes (0,0) (0,2) (2,0) (2,2), origin (X-1, Y-1), h+1 rows.
  // Steps 4..11: quarter-pel candidate s-3 of s_acMvRefineQ around the PU's half-pel winner, h rows.
  auto stage = [&](int s) {
    if (stOn) {
      int plane, dx, dy;
      if (s < 4) {
        plane = (s & 1) * 2 + (s >> 1) * 8;
        dx = -1;
        dy = -1;
      } else {
        int qx = 2 * bhx + c_refineQ[s - 3][0], qy = 2 * bhy + c_refineQ[s - 3][1];
        plane = (qy & 3) * 4 + (qx & 3);
        dx = qx >> 2;
        dy = qy >> 2;
      }
      // planes is 128-byte aligned and pitch / planeBytes are multiples of 128: aligning the offset aligns the address
      const long long off = slotOff + (long long)(plane * gPlaneBytes + dy * gPitch + dx);  // 16 planes < 2^31 bytes
      stage_rows<A>(stDst + ((s & 1) ? bufBytes : 0), planes + (off & ~(long long)(A - 1)) + stSrcOff, stRowSub, stRowStep,
                    s < 4 ? h + 1 : h, sg.RW, gPitch, stSecond);
    }
    cp_async_commit();
  };
  stage(0);  // in flight while the source tiles are fetched

  // ---- source tile(s) of this lane into registers ----
  // packed u8: 16 words (one 8x8 tile) or 8 words (two 4x4 tiles); BI: 16-bit pattern pairs, 32 or 16 words
  constexpr int OW = BI ? (TS == 8 ? 32 : 16) : (TS == 8 ? 16 : 8);
  unsigned o[OW];
  const bool had = useHad && !lossless;
  // 8x8 SATD of uni-prediction packs: tensor pipe (satd8x8_block_mma).  Warp-uniform; lossless PUs inside such a pack
  // get their SAD from the SWAR path afterwards.
  constexpr bool kMma = TS == 8 && !BI && PATH != 0;
  const bool useMma = kMma && useHad;
  [[maybe_unused]] int orgOffOwn = 0;  // byte offset of this lane's (first) source tile in the source picture
  // unit -> byte offset of its (first) tile inside a staged region, offset of the second 4x4 tile, source load
  int uOff = 0, u1Off = 0;
  bool uOn = false;
  // BI: pattern pair word = 2 * (source pair) - (other prediction pair); a plain 32-bit subtraction leaves the
  // carry-tolerant form (value = hi * 65536 + lo with signed lo) that the transforms expect
  auto patPair = [](unsigned srcWord, unsigned predWord, unsigned sel) {
    return 2u * __byte_perm(srcWord, 0, sel) - __byte_perm(predWord, 0, sel);
  };
  auto loadUnit = [&](int u) {
    uOn = laneActive && u < U;
    if (!uOn) return;
    if constexpr (BI && TS == 8) {
      int ty = u / ci.tilesX, tx = u - ty * ci.tilesX;
      uOff = ty * 8 * sg.RW + tx * 8;
      const uint8_t* src = org + (size_t)(oy + ty * 8) * gOrgPitch + ox + tx * 8;
      const uint8_t* oth = planes + otherOff + (long long)(ty * 8 * gPitch + tx * 8);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        unsigned s0, s1, p0, p1;
        ldg_row8(src + (size_t)r * gOrgPitch, s0, s1);
        ldg_row8(oth + (size_t)r * gPitch, p0, p1);
        o[4 * r + 0] = patPair(s0, p0, 0x4140); o[4 * r + 1] = patPair(s0, p0, 0x4342);
        o[4 * r + 2] = patPair(s1, p1, 0x4140); o[4 * r + 3] = patPair(s1, p1, 0x4342);
      }
    } else if constexpr (BI) {
      int t0 = 2 * u, t1 = 2 * u + 1;
      int ty0 = t0 / ci.tilesX, tx0 = t0 - ty0 * ci.tilesX;
      int ty1 = t1 / ci.tilesX, tx1 = t1 - ty1 * ci.tilesX;
      uOff = ty0 * 4 * sg.RW + tx0 * 4;
      u1Off = (ty1 * 4 * sg.RW + tx1 * 4) - uOff;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int ty = k ? ty1 : ty0, tx = k ? tx1 : tx0;
        const uint8_t* src = org + (size_t)(oy + ty * 4) * gOrgPitch + ox + tx * 4;
        const uint8_t* oth = planes + otherOff + (long long)(ty * 4 * gPitch + tx * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const unsigned sw = ldg_row4(src + (size_t)r * gOrgPitch), pw = ldg_row4(oth + (size_t)r * gPitch);
          o[8 * k + 2 * r + 0] = patPair(sw, pw, 0x4140);
          o[8 * k + 2 * r + 1] = patPair(sw, pw, 0x4342);
        }
      }
    } else if constexpr (TS == 8) {
      int ty = u / ci.tilesX, tx = u - ty * ci.tilesX;
      uOff = ty * 8 * sg.RW + tx * 8;
      orgOffOwn = (oy + ty * 8) * gOrgPitch + ox + tx * 8;  // < 2^31 (8K picture)
      if (!useMma) {
        const uint8_t* src = org + orgOffOwn;
#pragma unroll
        for (int r = 0; r < 8; ++r) ldg_row8(src + (size_t)r * gOrgPitch, o[2 * r], o[2 * r + 1]);
      }
    } else {
      int t0 = 2 * u, t1 = 2 * u + 1;
      int ty0 = t0 / ci.tilesX, tx0 = t0 - ty0 * ci.tilesX;
      int ty1 = t1 / ci.tilesX, tx1 = t1 - ty1 * ci.tilesX;
      uOff = ty0 * 4 * sg.RW + tx0 * 4;
      u1Off = (ty1 * 4 * sg.RW + tx1 * 4) - uOff;
      const uint8_t* s0 = org + (size_t)(oy + ty0 * 4) * gOrgPitch + ox + tx0 * 4;
      const uint8_t* s1 = org + (size_t)(oy + ty1 * 4) * gOrgPitch + ox + tx1 * 4;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        o[r] = ldg_row4(s0 + (size_t)r * gOrgPitch);
        o[4 + r] = ldg_row4(s1 + (size_t)r * gOrgPitch);
      }
    }
  };
  loadUnit(unit0);
  [[maybe_unused]] uint2* const so2 = reinterpret_cast<uint2*>(smem + K2_STAGE_BYTES) + lane;  // [row][lane]
  [[maybe_unused]] int uOff2 = 0;
  [[maybe_unused]] bool uOn2 = false;
  if constexpr (TS == 8 && A == 16 && !BI) {
    if (twoUnits) {
      const int u = unit0 + 32;
      uOn2 = laneActive && u < U;
      if (uOn2) {
        const int ty = u / ci.tilesX, tx = u - ty * ci.tilesX;
        uOff2 = ty * 8 * sg.RW + tx * 8;
        const uint8_t* src = org + (size_t)(oy + ty * 8) * gOrgPitch + ox + tx * 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          unsigned lo, hi;
          ldg_row8(src + (size_t)r * gOrgPitch, lo, hi);
          so2[r * 32] = make_uint2(lo, hi);
        }
      }
    }
  }

  // ---- tensor-pipe state: Hadamard A fragment, source rows of the tiles this lane feeds, as f16 pairs ----
  [[maybe_unused]] unsigned afrag[4];
  [[maybe_unused]] unsigned o16[4][8];
  // owner lane L = 8 cb + 2 t' + b reads its tile's SATD from lane (g = 2 cb + b, t = t')
  [[maybe_unused]] const int gatherLane = (lane & 24) + 4 * (lane & 1) + ((lane & 7) >> 1);
  if constexpr (kMma) {
    if (useMma) {
      had16_afrag(lane, afrag);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const int off = __shfl_sync(0xffffffffu, orgOffOwn, 8 * cb + (lane >> 2));
        const uint8_t* src = org + off + (size_t)(2 * (lane & 3)) * gOrgPitch;
        unsigned lo0, hi0, lo1, hi1;
        ldg_row8(src, lo0, hi0);
        ldg_row8(src + gOrgPitch, lo1, hi1);
        f16_rows(lo0, hi0, lo1, hi1, o16[cb]);
      }
    }
  }
  [[maybe_unused]] const unsigned bufSA = (unsigned)__cvta_generic_to_shared(smem);

  // ---- 12 steps: prefetch step s+1 while evaluating the candidates served by step s ----
  // running first-minimum: candidates of the half-pel stage are evaluated plane by plane, i.e. out of table
  // order, so ties are broken by the table index explicitly (strict < in table order, TEncSearch.cpp:1634)
  unsigned hBest = 0xffffffffu, qBest = 0xffffffffu;
  int hBestI = 9, qBestI = 0;
  // MV bits per axis for offsets -1, 0, +1.  Half stage, cost scale 1 (TEncSearch.cpp:4531): ((int<<1)+h)<<1 - pred
  // (bit counts are < 64: the three of an axis are packed into one register, byte t for offset t - 1)
  unsigned bitsX = 0, bitsY = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    bitsX |= (unsigned)golomb_bits((((mvIntX << 1) + (t - 1)) << 1) - predX) << (8 * t);
    bitsY |= (unsigned)golomb_bits((((mvIntY << 1) + (t - 1)) << 1) - predY) << (8 * t);
  }
  const int halfOff = sg.RW + ((alignX - 1) & (A - 1)) + 1;   // half-pel candidate (0, 0) inside a region
  int cq = 0;                                                 // sequence number of the next candidate
#pragma unroll 1
  for (int s = 0; s < 12; ++s) {
    const bool prefetch = (s != 3) && (s != 11);  // step 4 depends on the half-pel winner found after step 3
    if (prefetch) {
      stage(s + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const uint8_t* region = ((s & 1) ? bufB : bufA) + myPu * sg.RB;
    // candidates served by this step, in table order (TEncSearch.cpp:212-236):
    // s=0: H0 | s=1 (fx=2): H3 (-1,0), H4 (1,0) | s=2 (fy=2): H1 (0,-1), H2 (0,1) | s=3: H5..H8 | s>=4: Q(s-3)
    const int iCount = c_stepCount[s];
#pragma unroll 1
    for (int c = 0; c < iCount; ++c, ++cq) {
      // candidate cq of the pack's sequence: table index i, offset (ox, oy) in {-1, 0, 1}^2 in units of the current stage,
      // held as the PRMT selectors of byte ox + 1 / oy + 1 (k2_common.cuh)
      const int i = c_seqI[cq];
      const unsigned sel = c_seqSel[cq];
      int candOff;
      if (s < 4) {
        // inside the region (origin X-1, Y-1): x offset 1 + ((2hx)>>2), y offset 1 + ((2hy)>>2), i.e. one sample / one row
        // back for the offsets -1
        candOff = halfOff - ((sel & 3u) == 0 ? 1 : 0) - ((sel & 0x30000u) == 0 ? sg.RW : 0);
      } else {
        const int qx = 2 * bhx + (int)(sel & 3u) - 1;
        candOff = (alignX + (qx >> 2)) & (A - 1);
      }
      // exp-Golomb bit counts of the three possible vector components per axis were computed once per stage
      const int bits = (int)(__byte_perm(bitsX, 0, sel) + __byte_perm(bitsY, 0, sel >> 16));
      unsigned dist = 0;
      bool swar = true;
      if constexpr (kMma) {
        if (useMma) {
          swar = false;
          // every lane publishes the address of its own tile's candidate; the lanes of a column block fetch the
          // address of the tile they feed
          const unsigned cOwn = bufSA + ((s & 1) ? bufBytes : 0) + myPu * sg.RB + uOff + candOff;
          struct OrgRegs {
            const unsigned (*o)[8];
            __device__ __forceinline__ const unsigned (&operator()(int cb) const)[8] { return o[cb]; }
          };
          dist = satd8x8_pack_mma(cOwn, sg.RW, OrgRegs{o16}, afrag, lane, gatherLane);
          if (!uOn) dist = 0;
          if constexpr (A == 16) {
            if (twoUnits) {  // units 32.. of a 64x64 / 64x48 / 48x64 PU: source rows come from the parked tiles
              struct OrgSmem {
                const uint2* sp;
                int lane;
                mutable unsigned o2[8];
                __device__ __forceinline__ const unsigned (&operator()(int cb) const)[8] {
                  const uint2* q = sp + 8 * cb + (lane >> 2);
                  const uint2 r0 = q[(2 * (lane & 3)) * 32], r1 = q[(2 * (lane & 3) + 1) * 32];
                  f16_rows(r0.x, r0.y, r1.x, r1.y, o2);
                  return o2;
                }
              };
              const unsigned d2 = satd8x8_pack_mma(cOwn - uOff + uOff2, sg.RW,
                                                   OrgSmem{reinterpret_cast<const uint2*>(smem + K2_STAGE_BYTES), lane, {}},
                                                   afrag, lane, gatherLane);
              if (uOn2) dist += d2;
            }
          }
          if (__any_sync(0xffffffffu, laneActive && !had)) {  // lossless PUs in a Hadamard pack: SAD (TEncSearch.cpp:5258)
            if (uOn && !had) {
              const uint8_t* src = org + orgOffOwn;
              auto rowG = [&](int r, unsigned& lo, unsigned& hi) { ldg_row8(src + (size_t)r * gOrgPitch, lo, hi); };
              dist = sad8x8(rowG, region + uOff + candOff, sg.RW);
              if constexpr (A == 16) {
                if (uOn2) {
                  auto row2 = [&](int r, unsigned& lo, unsigned& hi) { const uint2 v = so2[r * 32]; lo = v.x; hi = v.y; };
                  dist += sad8x8(row2, region + uOff2 + candOff, sg.RW);
                }
              }
            }
          }
        }
      }
      if (swar) {
      if (uOn) dist = unit_dist<TS, BI, !kMma>(o, region + uOff + candOff, sg.RW, u1Off, had);
      if constexpr (TS == 8 && A == 16 && BI) {
        if (twoUnits) {  // rare (64x64, 64x48, 48x64 bi PUs): the second unit's pattern is rebuilt per candidate
          loadUnit(unit0 + 32);
          if (uOn) dist += unit_dist<TS, BI, !kMma>(o, region + uOff + candOff, sg.RW, u1Off, had);
          loadUnit(unit0);
        }
      }
      if constexpr (TS == 8 && A == 16 && !BI) {
        if (uOn2) {
          auto row2 = [&](int r, unsigned& lo, unsigned& hi) { const uint2 v = so2[r * 32]; lo = v.x; hi = v.y; };
          if constexpr (!kMma) dist += had ? satd8x8(row2, region + uOff2 + candOff, sg.RW) : sad8x8(row2, region + uOff2 + candOff, sg.RW);
          else dist += sad8x8(row2, region + uOff2 + candOff, sg.RW);
        }
      }
      }  // swar
      // per-PU sum: one REDUX when the PU owns the whole warp, xor-shuffles over the lane group otherwise
      if (lanesPerPu == 32) dist = __reduce_add_sync(0xffffffffu, dist);
      else for (int d = lanesPerPu >> 1; d > 0; d >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, d);
      if (laneActive) {
        dist += costLut[bits];
        if (s < 4) {
          if (dist < hBest || (dist == hBest && i < hBestI)) { hBest = dist; hBestI = i; }
        } else if (dist < qBest) {  // ascending i == table order, strict <
          qBest = dist;
          qBestI = i;
        }
      }
    }
    __syncwarp();  // all lanes are done with this buffer before step s+2 overwrites it
    if (s == 3) {
      const int bestI = hBestI < 9 ? hBestI : 0;
      bhx = c_refineH[bestI][0];
      bhy = c_refineH[bestI][1];
      qBest = hBest;  // candidate 0 of the quarter stage is the half-pel winner itself (same block, same bits)
      qBestI = 0;
      stage(4);
      // quarter stage, cost scale 0 (TEncSearch.cpp:5260): (((int<<1)+half)<<1) + q - pred
      bitsX = bitsY = 0;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        bitsX |= (unsigned)golomb_bits((((mvIntX << 1) + bhx) << 1) + (t - 1) - predX) << (8 * t);
        bitsY |= (unsigned)golomb_bits((((mvIntY << 1) + bhy) << 1) + (t - 1) - predY) << (8 * t);
      }
    }
  }

  if (laneActive && unit0 == 0) {
    fme_result* r = &res[puIdx];
    r->halfX = (int8_t)bhx; r->halfY = (int8_t)bhy;
    r->qterX = c_refineQ[qBestI][0]; r->qterY = c_refineQ[qBestI][1];
    r->cost = qBest;
  }
  __syncwarp();
}


// ------------------------------------------------------------------------------------------------
// MMA-native decomposition for 8x8-tiled uni-prediction PUs with Hadamard distortion (the default path)
// ------------------------------------------------------------------------------------------------
// A warp takes a "group": the eight tiles an m16n8k16 column block holds -- 8 / 4 / 2 PUs of 1 / 2 / 4 tiles, or one
// strip (R tile rows, R * tilesX <= 8 tiles) of a larger PU, strips being walked one after the other.  The four column
// blocks of one evaluation are four CANDIDATES of the same eight tiles, so a lane feeds rows 2t, 2t+1 of one tile
// (column g = lane >> 2) for all of them: the source rows are 8 registers instead of 32, the four fetch -> butterfly
// -> HMMA chains are independent (ILP), and the transposing reduction leaves lane (g, t) with tile column 2t + (g & 1)
// of candidate g >> 1.  Half-pel phase: the 4 planes of the strip are staged at once and the 9 candidates evaluated
// as 4 + 4 + 1; quarter-pel phase: the 8 candidate regions are staged as two batches of 4 into the two halves of the
// buffer (the second batch lands while the first is evaluated; with several strips the copies roll on).
// Lane roles: FEEDER of column g (PU pf, tile (tyF, txF)) and holder of the RESULT of column 2t + (g & 1) (PU pr);
// staging uses lane = [pf][plane k][sub], so the staging PU is the feeder PU.
constexpr int K2_GROUP_BYTES = 2 * 4608 + 16;  // two halves of 4 regions x P PUs (max: 8x8 and 16x8, 4 * 8 * 9 * 16) + read slack
static_assert(K2_GROUP_BYTES <= K2_SMEM_PER_WARP, "a group's staging buffer must fit the warp's shared-memory slice");

// row-2t address variant of cand_bfrags
__device__ __forceinline__ void cand_bfrags_at(unsigned a0, int RW, const unsigned (&o)[8], unsigned (&f)[8]) {
  const unsigned base = a0 & ~3u, sh = (a0 & 3u) * 8u;
  const unsigned w0 = lds_u32(base), w1 = lds_u32(base + 4), w2 = lds_u32(base + 8);
  const unsigned x0 = lds_u32(base + RW), x1 = lds_u32(base + RW + 4), x2 = lds_u32(base + RW + 8);
  f16_rows(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(x0, x1, sh),
           __funnelshift_r(x1, x2, sh), f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = hsub2u(f[i], o[i]);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const unsigned s0 = hadd2u(f[0 + r], f[2 + r]), d0 = hsub2u(f[0 + r], f[2 + r]);
    const unsigned s1 = hadd2u(f[4 + r], f[6 + r]), d1 = hsub2u(f[4 + r], f[6 + r]);
    f[0 + r] = hadd2u(s0, s1); f[2 + r] = hadd2u(d0, d1);
    f[4 + r] = hsub2u(s0, s1); f[6 + r] = hsub2u(d0, d1);
  }
}
__device__ __forceinline__ unsigned round_satd8(float z) {  // exact integer < 2^23 -> (z + 2) >> 2, TComRdCost.cpp:1421
  const unsigned v = (unsigned)__float_as_int(z + 8388608.0f) - 0x4b000000u;
  return (v + 2) >> 2;
}
// NC candidates of the group's eight tiles.  addr[j]: shared-window address of row 2t of the tile this lane feeds in
// candidate j.  NC == 4: lane (g, t) gets the SATD of tile column 2t + (g & 1) for candidate g >> 1;
// NC == 1: for the one candidate, replicated over g >> 1.
template <int NC>
__device__ __forceinline__ unsigned satd8x8_cands_mma(const unsigned (&addr)[NC], int RW, const unsigned (&o)[8],
                                                      const unsigned (&afrag)[4], int lane) {
  float a0[NC], a1[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    unsigned f[8];
    cand_bfrags_at(addr[j], RW, o, f);
    had_abs_sums(f, afrag, a0[j], a1[j]);
  }
  const bool g0 = (lane & 4) != 0, g1 = (lane & 8) != 0, g2 = (lane & 16) != 0;
  if constexpr (NC == 4) {
    float w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float keep = g0 ? a1[j] : a0[j], send = g0 ? a0[j] : a1[j];
      w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    float x[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float keep = g1 ? w[2 * k + 1] : w[2 * k], send = g1 ? w[2 * k] : w[2 * k + 1];
      x[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float keep = g2 ? x[1] : x[0], send = g2 ? x[0] : x[1];
    return round_satd8(keep + __shfl_xor_sync(0xffffffffu, send, 16));
  } else {
    float keep = g0 ? a1[0] : a0[0];
    const float send = g0 ? a0[0] : a1[0];
    keep += __shfl_xor_sync(0xffffffffu, send, 4);
    keep += __shfl_xor_sync(0xffffffffu, keep, 8);
    keep += __shfl_xor_sync(0xffffffffu, keep, 16);
    return round_satd8(keep);
  }
}

// One work item = one pack of k2_scatter (ci.P PUs of one shape: 32 tiles, or one PU of 48 / 64 tiles), walked as
// SUB "sub-items" of eight tiles: C groups of P PUs (tiles per PU <= 4), C whole PUs of eight tiles, or the nStrips
// strips of a larger PU.  Every sub-item contributes three JOBS -- H (its 4 half-pel planes, 9 candidates as 4 + 4 + 1),
// Qa and Qb (quarter-pel candidates 1..4 and 5..8 around its PU's half-pel winner) -- and the jobs of the work item
// form one list, H jobs first.  Job k is staged into buffer half k & 1, job k + 1 is in flight while job k is
// evaluated and job k + 2 is issued as soon as job k has freed its half, so the global -> shared copies of a job
// overlap the evaluation of the one before it (a Q job can only be issued once its PU's half-pel decision exists:
// with several PUs per work item that is long before it is needed).  The 16-byte record heads of the pack are fetched
// once, coalesced, into shared memory; per-PU half-pel decisions are parked there between the phases.
// Returns false (nothing done) when the pack holds a lossless PU: the caller then runs it through k2_pack.
constexpr int K2_GROUP_HEADS = K2_GROUP_BYTES;            // 32 x 4 words: record heads of the pack
constexpr int K2_GROUP_IDX = K2_GROUP_HEADS + 32 * 16;    // 32 PU indices
constexpr int K2_GROUP_DEC = K2_GROUP_IDX + 32 * 4;       // 32 x (half-pel winner, its cost)
static_assert(K2_GROUP_DEC + 32 * 8 <= K2_SMEM_PER_WARP, "group scratch must fit the warp's shared-memory slice");

template <int A>
__device__ __noinline__ bool k2_group_mma(int w, int h, const int* __restrict__ order, int first, int count,
                                          const fme_pu* __restrict__ pus, fme_result* __restrict__ res,
                                          const uint8_t* __restrict__ planes, const uint8_t* __restrict__ org,
                                          const FmeGeom& g, const uint32_t* __restrict__ costLut, uint8_t* smem) {
  const int lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
  // ---- the pack's record heads -> shared memory (one coalesced pass) ----
  unsigned* const sHead = reinterpret_cast<unsigned*>(smem + K2_GROUP_HEADS);
  int* const sIdx = reinterpret_cast<int*>(smem + K2_GROUP_IDX);
  uint2* const sDec = reinterpret_cast<uint2*>(smem + K2_GROUP_DEC);
  unsigned myFlags = 0;
  if (lane < count) {
    const int idx = order[first + lane];
    const unsigned* hp = reinterpret_cast<const unsigned*>(pus + idx);
    const unsigned h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2), h3 = __ldg(hp + 3);
    sHead[4 * lane + 0] = h0; sHead[4 * lane + 1] = h1; sHead[4 * lane + 2] = h2; sHead[4 * lane + 3] = h3;
    sIdx[lane] = idx;
    myFlags = h1 >> 24;
  }
  if (__any_sync(0xffffffffu, (myFlags & FME_PU_LOSSLESS) != 0)) return false;
  __syncwarp();

  // ---- class geometry (uniform) ----
  const int tilesX = w >> 3, tilesY = h >> 3, tpp = tilesX * tilesY;
  const bool small = tpp <= 4;                        // several whole PUs per sub-item
  const int lt = small ? (tpp >> 1) : 0;              // log2(tpp) for 1, 2, 4
  const int P = small ? (8 >> lt) : 1;                // PUs per sub-item
  const int R = small ? tilesY : 8 / tilesX;          // tile rows per sub-item
  const int nStrips = small ? 1 : (tilesY + R - 1) / R;
  const int nSub = small ? (count + P - 1) / P : count * nStrips;
  const int tps = small ? tpp : 8;                    // result columns per PU
  auto tile_row = [&](int j) { return small ? j >> (tilesX >> 1) : (j >= tilesX) + (j >= 2 * tilesX) + (j >= 3 * tilesX); };
  // lane roles inside a sub-item: feeder of column gq, holder of the result of column 2t + (gq & 1)
  const int pf = small ? gq >> lt : 0;
  const int jF = small ? gq & (tpp - 1) : gq;
  const int tyF = tile_row(jF), txF = jF - tyF * tilesX;
  const int nR = 2 * t + (gq & 1);
  const int pr = small ? nR >> lt : 0;
  const int tyR = tile_row(small ? nR & (tpp - 1) : nR);
  const int cq = gq >> 1;                             // this lane's candidate inside a batch of four
  const int gPitch = g.pitch, gOrgPitch = g.orgPitch, gPlaneBytes = (int)g.planeBytes;

  // ---- staging geometry (granule A, as in k2_pack); staging lane = [pf][slot k][sub] ----
  const int RW = ((w + A + A - 1) / A) * A;
  const int G = RW / A;
  const int hsMax = small ? h : R * 8;
  const int RB = (hsMax + 1) * RW;
  const int PRB = P * RB;                  // one plane / candidate slot: the regions of the sub-item's PUs
  const int HALF = (4 * PRB + 15) & ~15;
  const int LPI = small ? tpp : 8;         // lanes per (slot, PU) item
  const int lpiShift = small ? lt : 3;
  const int kS = (lane >> lpiShift) & 3, subS = lane & (LPI - 1);
  int gpShift = 0;
  while ((1 << gpShift) < G && (1 << gpShift) < LPI) ++gpShift;
  const int stGi = subS & ((1 << gpShift) - 1), stRow0 = subS >> gpShift, stRowStep = LPI >> gpShift;
  const bool stSecond = (1 << gpShift) < G;  // single-lane items (8x8 PUs): both granules of a row
  const unsigned bufSA = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned stDst = bufSA + (kS * P + pf) * RB + stRow0 * RW + stGi * A;
  const int stSrcOff = stRow0 * gPitch + stGi * A;
  const int feedTile = pf * RB + (tyF * 8 + 2 * t) * RW + txF * 8;

  unsigned afrag[4];
  had16_afrag(lane, afrag);

  // A job is addressed by a cursor (no divisions): H or Q phase, group-or-PU index `u` (small shapes: sub-item = group u of
  // P PUs; large shapes: PU u, strip `strip`), batch b of the Q phase.  PU slot of the pack in a lane role p: pu_of(u, p).
  struct Cursor {
    int k, u, strip, b;
    bool isH;
  };
  auto pu_of = [&](int u, int p) { return small ? u * P + p : u; };
  auto advance = [&](Cursor& c) {
    ++c.k;
    if (c.isH || c.b == 1) {
      c.b = 0;
      if (++c.strip == nStrips) { c.strip = 0; ++c.u; }
      if (c.isH && c.k == nSub) { c.isH = false; c.u = 0; c.strip = 0; }
    } else {
      c.b = 1;
    }
  };
  struct Feeder {  // what the feeder / staging role needs of its PU
    int X, Y, alignX, ox, oy;
    const uint8_t* slotBase;
    bool on;
  };
  auto feeder = [&](int u) {
    Feeder f;
    const int pc = pu_of(u, pf);
    f.on = pc < count;
    const unsigned* hd = sHead + 4 * min(pc, count - 1);
    const unsigned h0 = hd[0], h1 = hd[1], h2 = hd[2];
    const int px = (short)(h0 & 0xffff), py = (short)(h0 >> 16);
    f.X = min(max(px + (short)(h2 & 0xffff), -(g.M - 8)), g.W + g.M - 8 - w);  // as k2_pack: guards memory only
    f.Y = min(max(py + (short)(h2 >> 16), -(g.M - 8)), g.H + g.M - 8 - h);
    f.alignX = f.X + g.M;
    f.slotBase = planes + (size_t)min((int)((h1 >> 16) & 0xff), g.numSlots - 1) * g.slotBytes;
    f.ox = min(max(px, 0), g.W - w);
    f.oy = min(max(py, 0), g.H - h);
    return f;
  };
  // source rows 2t, 2t+1 of the fed tile: fetched one job ahead as aligned words, turned into f16 pairs at use
  auto org_fetch = [&](const Feeder& f, int strip, unsigned (&r)[4]) {
    const uint8_t* src = org + (size_t)(f.oy + min(strip * R * 8 + tyF * 8, h - 8) + 2 * t) * gOrgPitch + f.ox + txF * 8;
    ldg_row8(src, r[0], r[1]);
    ldg_row8(src + gOrgPitch, r[2], r[3]);
  };
  auto pu_sum = [&](unsigned v, bool on) {
    v = on ? v : 0u;
    if (tps >= 2) v += __shfl_xor_sync(0xffffffffu, v, 4);   // column bit 0 = g & 1
    if (tps >= 4) v += __shfl_xor_sync(0xffffffffu, v, 1);   // column bits 1, 2 = t
    if (tps >= 8) v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  };

  // ---- the job list: H jobs of all sub-items, then (Qa, Qb) per sub-item ----
  const int nJobs = 3 * nSub;
  auto issue = [&](const Cursor& c) {
    const int half = c.k & 1;
    const Feeder f = feeder(c.u);
#ifdef FME_K2_EXPERIMENT_NO_STAGING
    // measurement only (profiles/r2_k2_paths.txt): every instruction of the kernel except the global -> shared copies;
    // the candidates read are whatever the buffer holds, so the results are meaningless
    if (false) {
#else
    if (f.on && stGi < G) {
#endif
      int rows = small ? h : min(R, tilesY - c.strip * R) * 8, plane, dy = 0, ax;
      if (c.isH) {  // slot kS = plane (fx = 2 (kS & 1), fy = 2 (kS >> 1)), rows Y - 1 .. Y + hs of the strip
        rows += 1;
        plane = (kS & 1) * 2 + (kS >> 1) * 8;
        dy = -1;
        ax = f.alignX - 1;
      } else {      // slot kS = candidate q = 4 b + 1 + kS of s_acMvRefineQ around the PU's half-pel winner
        const unsigned dec = sDec[min(pu_of(c.u, pf), count - 1)].x;
        const int bhx = (int)(signed char)(dec & 0xff), bhy = (int)(signed char)((dec >> 8) & 0xff);
        const int q = 4 * c.b + 1 + kS;
        const int qx = 2 * bhx + c_refineQ[q][0], qy = 2 * bhy + c_refineQ[q][1];
        plane = (qy & 3) * 4 + (qx & 3);
        dy = qy >> 2;
        ax = f.alignX + (qx >> 2);
      }
      const int off = plane * gPlaneBytes + (f.Y + g.M + c.strip * R * 8 + dy) * gPitch + (ax & ~(A - 1));
      stage_rows<A>(stDst + half * HALF, f.slotBase + off + stSrcOff, stRow0, stRowStep, rows, RW, gPitch, stSecond);
    }
    cp_async_commit();
  };
  Cursor ic = {0, 0, 0, 0, true};  // next job to issue
  // issue every job below `lim` whose dependency is satisfied: a Q job needs the half-pel decision of its PU(s), made
  // by the H job of its last strip, i.e. job (u + 1) * nStrips - 1 (small shapes: nStrips == 1), which must be <= done
  auto issue_ready = [&](int lim, int done) {
    while (ic.k < lim && ic.k < nJobs && (ic.isH || (ic.u + 1) * nStrips - 1 <= done)) {
      issue(ic);
      advance(ic);
    }
  };

  unsigned hAcc[3] = {0, 0, 0}, qAcc[2] = {0, 0};
  unsigned oraw[4];
  {
    const Feeder f0 = feeder(0);
    org_fetch(f0, 0, oraw);
  }
  issue_ready(2, -1);
  Cursor ec = {0, 0, 0, 0, true};  // the job being evaluated
#pragma unroll 1
  for (int k = 0; k < nJobs; ++k) {
    const bool isH = ec.isH;
    const int u = ec.u, b = ec.b, strip = ec.strip;
    const int rowsIn = small ? tilesY : min(R, tilesY - strip * R);
    const Feeder f = feeder(u);
    if (ic.k <= k) issue_ready(k + 1, k - 1);          // (a Q job whose decision only just became available)
    if (ic.k > k + 1) cp_async_wait<1>(); else cp_async_wait<0>();
    unsigned o[8];
    f16_rows(oraw[0], oraw[1], oraw[2], oraw[3], o);
    // the next job's source rows (another sub-item unless this is Qa -> Qb of the same one)
    advance(ec);
    if (ec.k < nJobs && (ec.u != u || ec.strip != strip)) {
      const Feeder fn = feeder(ec.u);
      org_fetch(fn, ec.strip, oraw);
    }
    __syncwarp();
    const int pcR = pu_of(u, pr);
    const bool onR = pcR < count && tyR < rowsIn;
    const unsigned hb = bufSA + (k & 1) * HALF;
    if (isH) {
      // candidate c of s_acMvRefineH: slot k_c, (dx, dy) = ((2hx)>>2, (2hy)>>2); regions start at (X-1, Y-1)
      //   c:   0      1       2      3       4      5        6       7       8
      //  hx,hy (0,0) (0,-1) (0,1) (-1,0)  (1,0) (-1,-1)  (1,-1)  (-1,1)  (1,1)
      const unsigned base = hb + feedTile + RW + ((f.alignX - 1) & (A - 1)) + 1;
      if (strip == 0) hAcc[0] = hAcc[1] = hAcc[2] = 0;
      {
        const unsigned a[4] = {base, base + 2 * PRB - RW, base + 2 * PRB, base + PRB - 1};
        hAcc[0] += pu_sum(satd8x8_cands_mma<4>(a, RW, o, afrag, lane), onR);
      }
      {
        const unsigned a[4] = {base + PRB, base + 3 * PRB - RW - 1, base + 3 * PRB - RW, base + 3 * PRB - 1};
        hAcc[1] += pu_sum(satd8x8_cands_mma<4>(a, RW, o, afrag, lane), onR);
      }
      {
        const unsigned a[1] = {base + 3 * PRB};
        hAcc[2] += pu_sum(satd8x8_cands_mma<1>(a, RW, o, afrag, lane), onR);
      }
      if (strip == nStrips - 1) {
        // ---- half-pel decision (TEncSearch.cpp:1634 strict <, first minimum in table order) ----
        // MV bits, cost scale 1 (TEncSearch.cpp:4531): ((int << 1) + h) << 1 against the predictor
        const unsigned* hd = sHead + 4 * min(pcR, count - 1);
        const unsigned r2 = hd[2], r3 = hd[3];
        const int mvIntX = (short)(r2 & 0xffff), mvIntY = (short)(r2 >> 16);
        const int predX = (short)(r3 & 0xffff), predY = (short)(r3 >> 16);
        auto half_cost = [&](int c, unsigned dist) {
          const int hx = c_refineH[c][0], hy = c_refineH[c][1];
          return dist + costLut[golomb_bits((((mvIntX << 1) + hx) << 1) - predX) + golomb_bits((((mvIntY << 1) + hy) << 1) - predY)];
        };
        unsigned bestC = half_cost(cq, hAcc[0]);
        int bestI = cq;
        const unsigned c1 = half_cost(4 + cq, hAcc[1]);
        if (c1 < bestC) { bestC = c1; bestI = 4 + cq; }
        const unsigned c2 = half_cost(8, hAcc[2]);
        if (c2 < bestC) { bestC = c2; bestI = 8; }
#pragma unroll
        for (int d = 8; d <= 16; d <<= 1) {  // across the four candidate lanes of a batch
          const unsigned oc = __shfl_xor_sync(0xffffffffu, bestC, d);
          const int oi = __shfl_xor_sync(0xffffffffu, bestI, d);
          if (oc < bestC || (oc == bestC && oi < bestI)) { bestC = oc; bestI = oi; }
        }
        if (pcR < count && nR == pr * tps && cq == 0)
          sDec[pcR] = make_uint2((unsigned)(c_refineH[bestI][0] & 0xff) | ((unsigned)(c_refineH[bestI][1] & 0xff) << 8), bestC);
      }
    } else {
      // quarter-pel: the staged regions start at the candidate's integer origin; x offsets (2 bhx + ox) >> 2
      //   q:    1      2       3       4       5      6      7      8
      //   ox:   0      0      -1       1      -1      1     -1      1
      const int bhxF = (int)(signed char)(sDec[min(pu_of(u, pf), count - 1)].x & 0xff);
      const unsigned fq = hb + feedTile;
      const unsigned qm = fq + ((f.alignX + ((2 * bhxF - 1) >> 2)) & (A - 1));
      const unsigned q0 = fq + ((f.alignX + ((2 * bhxF) >> 2)) & (A - 1));
      const unsigned qp = fq + ((f.alignX + ((2 * bhxF + 1) >> 2)) & (A - 1));
      if (strip == 0 && b == 0) qAcc[0] = qAcc[1] = 0;
      if (b == 0) {
        const unsigned a[4] = {q0, q0 + PRB, qm + 2 * PRB, qp + 3 * PRB};
        qAcc[0] += pu_sum(satd8x8_cands_mma<4>(a, RW, o, afrag, lane), onR);
      } else {
        const unsigned a[4] = {qm, qp + PRB, qm + 2 * PRB, qp + 3 * PRB};
        qAcc[1] += pu_sum(satd8x8_cands_mma<4>(a, RW, o, afrag, lane), onR);
      }
      if (strip == nStrips - 1 && b == 1) {
        // ---- quarter-pel decision: candidate 0 is the half-pel winner itself (same block, same bits) ----
        // cost scale 0 (TEncSearch.cpp:5260): (((int << 1) + half) << 1) + q against the predictor
        const unsigned* hd = sHead + 4 * min(pcR, count - 1);
        const unsigned r2 = hd[2], r3 = hd[3];
        const int mvIntX = (short)(r2 & 0xffff), mvIntY = (short)(r2 >> 16);
        const int predX = (short)(r3 & 0xffff), predY = (short)(r3 >> 16);
        const uint2 dec = sDec[min(pcR, count - 1)];
        const int bhx = (int)(signed char)(dec.x & 0xff), bhy = (int)(signed char)((dec.x >> 8) & 0xff);
        auto qter_cost = [&](int q, unsigned dist) {
          const int bx = (((mvIntX << 1) + bhx) << 1) + c_refineQ[q][0], by = (((mvIntY << 1) + bhy) << 1) + c_refineQ[q][1];
          return dist + costLut[golomb_bits(bx - predX) + golomb_bits(by - predY)];
        };
        unsigned qC = qter_cost(1 + cq, qAcc[0]);
        int qI = 1 + cq;
        const unsigned c1 = qter_cost(5 + cq, qAcc[1]);
        if (c1 < qC) { qC = c1; qI = 5 + cq; }
#pragma unroll
        for (int d = 8; d <= 16; d <<= 1) {
          const unsigned oc = __shfl_xor_sync(0xffffffffu, qC, d);
          const int oi = __shfl_xor_sync(0xffffffffu, qI, d);
          if (oc < qC || (oc == qC && oi < qI)) { qC = oc; qI = oi; }
        }
        if (dec.y <= qC) { qC = dec.y; qI = 0; }  // strict < against the running best that starts at candidate 0
        if (pcR < count && nR == pr * tps && cq == 0) {
          fme_result* r = &res[sIdx[pcR]];
          r->halfX = (int8_t)bhx; r->halfY = (int8_t)bhy;
          r->qterX = c_refineQ[qI][0]; r->qterY = c_refineQ[qI][1];
          r->cost = qC;
        }
      }
    }
    __syncwarp();  // this job's buffer half is free, and any decision it made is visible
    issue_ready(k + 3, k);
  }
  __syncwarp();
  return true;
}

// warps per CTA (one CTA per SM): SWAR k2_pack needs 168 registers (12 warps); k2_group_mma runs at 128 (16 warps)
#ifndef FME_K2_WARPS_GROUP
#define FME_K2_WARPS_GROUP 16
#endif
template <int PATH>
constexpr int k2_warps() { return PATH == 2 ? FME_K2_WARPS_GROUP : K2_WARPS; }

// NNF != 0 (uni-prediction SWAR kernel only): K3's work rides along as extra work items -- NN_pred for 64 list-order PUs per
// item (k3_common.cuh NnWarpNet, the shipped 17-22-20-49 shape; NNF 1 exact, 2 fused multiply-add), one after every R SATD
// packs.  The idea: an NN item is independent fp32 work that needs nothing but the records, so it could fill the issue
// slots the integer packs leave idle (K2 alone issues on 67 % of the cycles).  Measured (fme_config.k3Fuse = 1, off by
// default): the results are identical, but a pack body (23 KB) and the NN item's code do not fit the SM's 32 KB instruction
// cache together -- 1.365 ms per 1080p frame with the stand-alone kernel's unrolling (48 KB of NN code,
// stall_no_instruction 3.0 per issue), 0.961 ms with the compact item below (25 KB), against 0.757 + 0.143 ms apart.
using K2Nn = NnWarpNet<2, 22, 20, 49, false>;
using K2NnFma = NnWarpNet<2, 22, 20, 49, true>;
static_assert(K2Nn::ACT_WORDS * 4 <= K2_SMEM_PER_WARP, "the activation transpose of an NN item lives in the warp's staging slice");

template <bool BI, int PATH, int NNF = 0>
__global__ void __launch_bounds__(k2_warps<PATH>() * 32, 1)
k2_refine(const fme_pu* __restrict__ pus, fme_result* __restrict__ res, const uint8_t* __restrict__ planes,
          const uint8_t* __restrict__ org, const FmeGeom g, const FmeCostLut costLutG, int useHad,
          const int* __restrict__ order, const int* __restrict__ classOffset, const int* __restrict__ packOffset,
          int* __restrict__ workCounter, int nPus, const float* __restrict__ nnBlob, float nnOutClamp) {
  constexpr int mmaGroups = PATH == 2 ? 1 : 0;
  extern __shared__ __align__(16) uint8_t dynSmem[];
  constexpr int kWarps = PATH == 2 ? FME_K2_WARPS_GROUP : K2_WARPS;
  [[maybe_unused]] float* const s_net = reinterpret_cast<float*>(dynSmem + kWarps * K2_SMEM_PER_WARP);
  if constexpr (NNF != 0) K2Nn::load(s_net, nnBlob);   // (same image for both arithmetic modes)
  __shared__ uint32_t s_lut[FME_COST_LUT_SIZE];
  __shared__ int s_packOff[FME_K2_KEYS + 1];
  __shared__ int s_classOff[FME_K2_KEYS + 1];

  for (int i = threadIdx.x; i < FME_COST_LUT_SIZE; i += blockDim.x) s_lut[i] = costLutG.v[i];
  for (int i = threadIdx.x; i <= FME_K2_KEYS; i += blockDim.x) {
    s_packOff[i] = packOffset[i];
    s_classOff[i] = classOffset[i];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = dynSmem + warp * K2_SMEM_PER_WARP;
  const int totalPacks = s_packOff[FME_K2_KEYS];
  // work items: the packs, and with NNF groups of R packs followed by one NN item
  [[maybe_unused]] const int nnItems = (nPus + K2Nn::PUS - 1) / K2Nn::PUS;
  [[maybe_unused]] const int nnR = max(1, totalPacks / max(nnItems, 1));
  const int totalItems = NNF != 0 ? max((totalPacks + nnR - 1) / nnR, nnItems) * (nnR + 1) : totalPacks;

  // dynamic pack scheduler; the next pack index is requested while the current pack is processed
  int nextPack = 0;
  if (lane == 0) nextPack = atomicAdd(workCounter, 1);
  while (true) {
    const int item = __shfl_sync(0xffffffffu, nextPack, 0);
    if (item >= totalItems) break;
    if (lane == 0) nextPack = atomicAdd(workCounter, 1);
    int pack = item;
    if constexpr (NNF != 0) {
      const int q = item / (nnR + 1), r = item - q * (nnR + 1);
      if (r == nnR) {
        if (q < nnItems) {
          if constexpr (NNF == 2) K2NnFma::warp_item(s_net, reinterpret_cast<float*>(smem), pus, res, q * K2Nn::PUS, nPus, lane, nnOutClamp);
          else K2Nn::warp_item(s_net, reinterpret_cast<float*>(smem), pus, res, q * K2Nn::PUS, nPus, lane, nnOutClamp);
          __syncwarp();
        }
        continue;
      }
      pack = q * nnR + r;
      if (pack >= totalPacks) continue;
    }
    // schedule position of this pack: the last v with packOff[v] <= pack (binary search, uniform across the warp);
    // position v holds shape class 63 - (v & 63) of slot group v >> 6 (k2_scatter)
    int v = 0;
    {
      int hi = FME_K2_KEYS;
      while (hi - v > 1) {
        const int mid = (v + hi) >> 1;
        if (s_packOff[mid] <= pack) v = mid; else hi = mid;
      }
    }
    ClassInfo ci = class_info(63 - (v & 63));
    int inClass = s_classOff[v + 1] - s_classOff[v];
    const int packP = pack_pus(ci, mmaGroups);
    int first = (pack - s_packOff[v]) * packP;
    int count = min(packP, inClass - first);
    first += s_classOff[v];
#define K2_ARGS ci, order, first, count, pus, res, planes, org, g, s_lut, useHad, smem
    if constexpr (PATH == 2) {
      if (ci.ts == 8) {
        const bool done = ci.w >= 16 ? k2_group_mma<16>(ci.w, ci.h, order, first, count, pus, res, planes, org, g, s_lut, smem)
                                     : k2_group_mma<8>(ci.w, ci.h, order, first, count, pus, res, planes, org, g, s_lut, smem);
        if (done) continue;  // otherwise (a lossless PU in the group) the generic path below serves it
      }
    }
    if (ci.ts == 8) {
      if (ci.w >= 16) k2_pack<8, 16, BI, PATH>(K2_ARGS);
      else k2_pack<8, 8, BI, PATH>(K2_ARGS);
    } else {
      if (ci.w >= 16) k2_pack<4, 16, BI, PATH>(K2_ARGS);
      else if (ci.w >= 8) k2_pack<4, 8, BI, PATH>(K2_ARGS);
      else k2_pack<4, 4, BI, PATH>(K2_ARGS);
    }
#undef K2_ARGS
  }
}

// ------------------------------------------------------------------------------------------------
// Prediction error of a uni-predicted PU at a given quarter-pel MV: luma motion compensation + HADs/SAD against
// the source block, i.e. TEncSearch::xGetInterPredictionError (TEncSearch.cpp:3576-3596, used by merge estimation)
// and the distortion part of xGetTemplateCost (TEncSearch.cpp:4397-4436, AMVP candidate cost, SAD).
// The MC block at MV (mx,my) is plane P[my&3][mx&3] at integer offset (mx>>2, my>>2)  (SURVEY.md A.1).
// One warp per item: the candidate block is staged into shared memory, lanes take tiles.
// ------------------------------------------------------------------------------------------------
constexpr int PE_WARPS = 4;
constexpr int PE_SMEM_PER_WARP = 64 * 68 + 16;

__global__ void __launch_bounds__(PE_WARPS * 32) k_pred_error(const fme_mc_pu* __restrict__ pus, int n,
                                                              const uint8_t* __restrict__ planes,
                                                              const uint8_t* __restrict__ org, const FmeGeom g,
                                                              int useHad, uint32_t* __restrict__ out) {
  __shared__ __align__(16) uint8_t s_buf[PE_WARPS][PE_SMEM_PER_WARP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* buf = s_buf[warp];
  for (int i = blockIdx.x * PE_WARPS + warp; i < n; i += gridDim.x * PE_WARPS) {
    const fme_mc_pu p = pus[i];
    const int w = p.w, h = p.h;
    if (!fme_hevc_pu_shape(w, h)) {
      if (lane == 0) out[i] = 0xffffffffu;
      continue;
    }
    const int fx = p.mvX & 3, fy = p.mvY & 3;
    const int X = min(max(p.x + (p.mvX >> 2), -(g.M - 8)), g.W + g.M - 8 - w);
    const int Y = min(max(p.y + (p.mvY >> 2), -(g.M - 8)), g.H + g.M - 8 - h);
    const int ox = min(max((int)p.x, 0), g.W - w), oy = min(max((int)p.y, 0), g.H - h);
    const uint8_t* src = planes + (size_t)min((int)p.refSlot, g.numSlots - 1) * g.slotBytes +
                         (size_t)(fy * 4 + fx) * g.planeBytes + (size_t)(Y + g.M) * g.pitch + ((X + g.M) & ~3);
    const int al = (X + g.M) & 3;
    const int RW = w + 4, wpr = RW >> 2;
    __syncwarp();
    for (int t = lane; t < h * wpr; t += 32) {
      int r = t / wpr, c = t - r * wpr;
      reinterpret_cast<unsigned*>(buf + r * RW)[c] = __ldg(reinterpret_cast<const unsigned*>(src + (size_t)r * g.pitch) + c);
    }
    __syncwarp();
    const bool had = useHad && !(p.flags & FME_PU_LOSSLESS);
    const int ts = ((w & 7) == 0 && (h & 7) == 0) ? 8 : 4;
    const int tilesX = w / ts, tiles = tilesX * (h / ts);
    unsigned dist = 0;
    for (int t = lane; t < tiles; t += 32) {
      const int ty = t / tilesX, tx = t - ty * tilesX;
      const uint8_t* o = org + (size_t)(oy + ty * ts) * g.orgPitch + ox + tx * ts;
      const uint8_t* c = buf + ty * ts * RW + al + tx * ts;
      if (ts == 8) {
        auto row = [&](int r, unsigned& lo, unsigned& hi) { ldg_row8(o + (size_t)r * g.orgPitch, lo, hi); };
        dist += had ? satd8x8(row, c, RW) : sad8x8(row, c, RW);
      } else {
        unsigned oo[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) oo[r] = ldg_row4(o + (size_t)r * g.orgPitch);
        dist += had ? satd4x4(oo, c, RW) : sad4x4(oo, c, RW);
      }
    }
    dist = __reduce_add_sync(0xffffffffu, dist);
    if (lane == 0) out[i] = dist;
  }
}

// Batched candidate costs (fme_cand_cost): k_pred_error's distortion + the bit cost of the candidate's side information,
// one warp per candidate; a second pass picks the first minimum of every group.
__global__ void __launch_bounds__(PE_WARPS * 32) k_cand_cost(const fme_cand_pu* __restrict__ cands, int n,
                                                             const uint8_t* __restrict__ planes,
                                                             const uint8_t* __restrict__ org, const FmeGeom g,
                                                             const FmeCostLut lut, int useHad, uint32_t* __restrict__ out) {
  __shared__ __align__(16) uint8_t s_buf[PE_WARPS][PE_SMEM_PER_WARP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* buf = s_buf[warp];
  for (int i = blockIdx.x * PE_WARPS + warp; i < n; i += gridDim.x * PE_WARPS) {
    const fme_cand_pu p = cands[i];
    const int w = p.w, h = p.h;
    if (!fme_hevc_pu_shape(w, h)) {
      if (lane == 0) out[i] = 0xffffffffu;
      continue;
    }
    const int fx = p.mvX & 3, fy = p.mvY & 3;
    const int X = min(max(p.x + (p.mvX >> 2), -(g.M - 8)), g.W + g.M - 8 - w);
    const int Y = min(max(p.y + (p.mvY >> 2), -(g.M - 8)), g.H + g.M - 8 - h);
    const int ox = min(max((int)p.x, 0), g.W - w), oy = min(max((int)p.y, 0), g.H - h);
    const uint8_t* src = planes + (size_t)min((int)p.refSlot, g.numSlots - 1) * g.slotBytes +
                         (size_t)(fy * 4 + fx) * g.planeBytes + (size_t)(Y + g.M) * g.pitch + ((X + g.M) & ~3);
    const int al = (X + g.M) & 3;
    const int RW = w + 4, wpr = RW >> 2;
    __syncwarp();
    for (int t = lane; t < h * wpr; t += 32) {
      const int r = t / wpr, c = t - r * wpr;
      reinterpret_cast<unsigned*>(buf + r * RW)[c] = __ldg(reinterpret_cast<const unsigned*>(src + (size_t)r * g.pitch) + c);
    }
    __syncwarp();
    const bool had = useHad && !(p.flags & (FME_PU_LOSSLESS | FME_CAND_SAD));
    const int ts = ((w & 7) == 0 && (h & 7) == 0) ? 8 : 4;
    const int tilesX = w / ts, tiles = tilesX * (h / ts);
    unsigned dist = 0;
    for (int t = lane; t < tiles; t += 32) {
      const int ty = t / tilesX, tx = t - ty * tilesX;
      const uint8_t* o = org + (size_t)(oy + ty * ts) * g.orgPitch + ox + tx * ts;
      const uint8_t* c = buf + ty * ts * RW + al + tx * ts;
      if (ts == 8) {
        auto row = [&](int r, unsigned& lo, unsigned& hi) { ldg_row8(o + (size_t)r * g.orgPitch, lo, hi); };
        dist += had ? satd8x8(row, c, RW) : sad8x8(row, c, RW);
      } else {
        unsigned oo[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) oo[r] = ldg_row4(o + (size_t)r * g.orgPitch);
        dist += had ? satd4x4(oo, c, RW) : sad4x4(oo, c, RW);
      }
    }
    dist = __reduce_add_sync(0xffffffffu, dist);
    if (lane == 0) out[i] = dist + lut.v[min((int)p.bits, FME_COST_LUT_SIZE - 1)];
  }
}
__global__ void k_cand_best(const fme_cand_pu* __restrict__ cands, int n, const uint32_t* __restrict__ cost,
                            int32_t* __restrict__ best) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!cands[i].groupStart && i != 0) { best[i] = -1; return; }
  unsigned bc = cost[i];
  int bi = i;
  for (int j = i + 1; j < n && !cands[j].groupStart; ++j)  // groups are a handful of candidates (<= 5 merge, <= 3 AMVP)
    if (cost[j] < bc) { bc = cost[j]; bi = j; }              // strict <: the first minimum wins
  best[i] = bi;
}

// Compact luma motion compensation (fme_mc_luma_compact): one warp per PU; the block is plane P[my & 3][mx & 3] at the
// integer part of the MV, copied 4 samples per lane step through an aligned-load funnel.
__global__ void __launch_bounds__(256) k_mc_luma_compact(const fme_mc_pu* __restrict__ pus, int n,
                                                         const uint32_t* __restrict__ offsets,
                                                         const uint8_t* __restrict__ planes, const FmeGeom g,
                                                         uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warpsPerGrid) {
    const fme_mc_pu p = pus[i];
    const int w = p.w, h = p.h;
    if (w < 4 || h < 4 || w > 64 || h > 64 || (w & 3)) continue;
    const int X = min(max(p.x + (p.mvX >> 2), -(g.M - 8)), g.W + g.M - 8 - w);
    const int Y = min(max(p.y + (p.mvY >> 2), -(g.M - 8)), g.H + g.M - 8 - h);
    const uint8_t* src = planes + (size_t)min((int)p.refSlot, g.numSlots - 1) * g.slotBytes +
                         (size_t)((p.mvY & 3) * 4 + (p.mvX & 3)) * g.planeBytes + (size_t)(Y + g.M) * g.pitch + (X + g.M);
    unsigned* dst = reinterpret_cast<unsigned*>(out + (offsets[i] & ~3u));
    const int wq = w >> 2, items = wq * h;
    for (int t = lane; t < items; t += 32) {
      const int r = t / wq, c = t - r * wq;
      dst[t] = ldg_row4(src + (size_t)r * g.pitch + 4 * c);
    }
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

cudaError_t fme_k2_bin(const fme_pu* d_pus, int n, fme_result* d_res, int wantBi, int biServed, const FmeK2Scratch& sc,
                       int packMode, int numSMs, cudaStream_t s, int64_t* launches) {
  // classCount[64], classCursor[64], the work counter and (behind it) the FME_PU_ERR_ON_GPU count are adjacent (fme_create);
  // the bi-predictive pass leaves the count of the uni-prediction pass alone
  cudaError_t e = cudaMemsetAsync(sc.classCount, 0, sizeof(int) * (2 * FME_K2_KEYS + (wantBi ? 1 : 2)), s);
  if (e != cudaSuccess) return e;
  int blocks = min(numSMs * 4, (n + 255) / 256);
  k2_count<<<blocks, 256, 0, s>>>(d_pus, n, sc.classCount, wantBi, d_res, biServed, sc.keys, wantBi ? nullptr : sc.workCounter + 1);
  k2_scatter<<<blocks, 256, 0, s>>>(sc.keys, n, sc.classCount, sc.classOffset, sc.packOffset, sc.classCursor, sc.order, packMode);
  *launches += 2;
  return cudaGetLastError();
}

template <bool BI, int PATH, int NNF = 0>
static cudaError_t launch_k2_pass(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_pu* d_pus, int n,
                                  fme_result* d_res, const FmeCostLut& costLut, int useHad, int biServed, const FmeK2Scratch& sc,
                                  int numSMs, cudaStream_t s, int64_t* launches, const FmeK2NnFuse* nn = nullptr) {
  cudaError_t e = fme_k2_bin(d_pus, n, BI ? nullptr : d_res, BI ? 1 : 0, biServed, sc, PATH == 2 ? 1 : 0, numSMs, s, launches);
  if (e != cudaSuccess) return e;
  const int smemBytes = k2_warps<PATH>() * K2_SMEM_PER_WARP + (NNF != 0 ? K2Nn::WORDS * 4 : 0);
  // the opt-in to > 48 KB dynamic shared memory is a per-device, per-function attribute; setting it is cheap and
  // idempotent, so it is simply set before every launch (no shared state between host threads / contexts)
  e = cudaFuncSetAttribute(k2_refine<BI, PATH, NNF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
  if (e != cudaSuccess) return e;
  k2_refine<BI, PATH, NNF><<<numSMs, k2_warps<PATH>() * 32, smemBytes, s>>>(
      d_pus, d_res, d_planes, d_org, g, costLut, useHad, sc.order, sc.classOffset, sc.packOffset, sc.workCounter, n,
      nn ? nn->d_blob : nullptr, nn ? nn->outClamp : 0.f);
  ++*launches;
  return cudaGetLastError();
}

// k2Path: how the 8x8 SATD of uni-prediction Hadamard PUs is computed (fme_config.k2Path): FME_K2_PATH_SWAR (integer
// SWAR in registers), FME_K2_PATH_MMA_PACK (tensor pipe inside the 32-lane packs) or FME_K2_PATH_MMA_GROUP (tensor pipe,
// eight-tile groups x four candidates).  All three are bit-identical; SAD mode and bi-predictive records always take
// the SWAR kernels.
// biPred != 0: a second binning + refinement pass serves the bi-predictive refinement records (FME_PU_BI); with
// biPred == 0 such records are left untouched (the synchronous entry points reject them).
// nn != nullptr: the caller also wants NN_pred for the same records and the net has the shipped shape; *nnDone is set when
// the pass that ran carried K3's work along (the SWAR kernel), otherwise the caller launches K3 itself.
cudaError_t fme_launch_k2(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_pu* d_pus, int n,
                          fme_result* d_res, const FmeCostLut& costLut, int useHad, int biPred, int k2Path,
                          const FmeK2Scratch& sc, int numSMs, cudaStream_t s, int64_t* launches, const FmeK2NnFuse* nn,
                          bool* nnDone) {
  if (nnDone) *nnDone = false;
  if (n <= 0) return cudaSuccess;
  cudaError_t e;
  if (nn && nnDone && (!useHad || k2Path == FME_K2_PATH_SWAR)) {
    e = nn->fma ? launch_k2_pass<false, 0, 2>(g, d_planes, d_org, d_pus, n, d_res, costLut, useHad, biPred, sc, numSMs, s, launches, nn)
                : launch_k2_pass<false, 0, 1>(g, d_planes, d_org, d_pus, n, d_res, costLut, useHad, biPred, sc, numSMs, s, launches, nn);
    *nnDone = e == cudaSuccess;
    if (e != cudaSuccess || !biPred) return e;
    return launch_k2_pass<true, 0>(g, d_planes, d_org, d_pus, n, d_res, costLut, useHad, biPred, sc, numSMs, s, launches);
  }
#define K2_UNI(PATH) launch_k2_pass<false, PATH>(g, d_planes, d_org, d_pus, n, d_res, costLut, useHad, biPred, sc, numSMs, s, launches)
  if (!useHad || k2Path == FME_K2_PATH_SWAR) e = K2_UNI(0);
  else if (k2Path == FME_K2_PATH_MMA_PACK) e = K2_UNI(1);
  else if (k2Path == FME_K2_PATH_UMMA) e = fme_launch_k2_umma(g, d_planes, d_org, d_pus, n, d_res, costLut, biPred, sc, numSMs, s, launches);
  else e = K2_UNI(2);
#undef K2_UNI
  if (e != cudaSuccess || !biPred) return e;
  return launch_k2_pass<true, 0>(g, d_planes, d_org, d_pus, n, d_res, costLut, useHad, biPred, sc, numSMs, s, launches);
}

cudaError_t fme_launch_cand_cost(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_cand_pu* d_cands,
                                 int n, const FmeCostLut& lut, int useHad, uint32_t* d_cost, int32_t* d_best, cudaStream_t s,
                                 int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  const int blocks = fme_one_wave(k_cand_cost, PE_WARPS * 32, 0, (n + PE_WARPS - 1) / PE_WARPS);
  k_cand_cost<<<blocks, PE_WARPS * 32, 0, s>>>(d_cands, n, d_planes, d_org, g, lut, useHad, d_cost);
  ++*launches;
  if (d_best) {
    k_cand_best<<<(n + 255) / 256, 256, 0, s>>>(d_cands, n, d_cost, d_best);
    ++*launches;
  }
  return cudaGetLastError();
}

cudaError_t fme_launch_mc_luma_compact(const FmeGeom& g, const uint8_t* d_planes, const fme_mc_pu* d_pus, int n,
                                       const uint32_t* d_offsets, uint8_t* d_out, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  const int blocks = fme_one_wave(k_mc_luma_compact, 256, 0, (n + 7) / 8);  // 8 warps per CTA, one PU per warp step
  k_mc_luma_compact<<<blocks, 256, 0, s>>>(d_pus, n, d_offsets, d_planes, g, d_out);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t fme_launch_pred_error(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_mc_pu* d_pus,
                                  int n, int useHad, uint32_t* d_out, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  const int blocks = fme_one_wave(k_pred_error, PE_WARPS * 32, 0, (n + PE_WARPS - 1) / PE_WARPS);
  k_pred_error<<<blocks, PE_WARPS * 32, 0, s>>>(d_pus, n, d_planes, d_org, g, useHad, d_out);
  ++*launches;
  return cudaGetLastError();
}
