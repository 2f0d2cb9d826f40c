// K2 on the asynchronous tensor path (FME_K2_PATH_UMMA): the same search as k2_refine.cu (xPatternRefinement
// TEncSearch.cpp:1591-1645 over the candidates of xPatternSearchFracDIF :5232-5269, distortion xCalcHADs8x8 / xCalcHADs4x4
// TComRdCost.cpp:1234-1425, MV-bit cost :172-185), same packs, same staging, same argmin -- but the Hadamard transform of
// every tile-candidate is one row of a tcgen05.mma:
//
//   D[M = 128 lane units][N coefficients] (s32, TMEM) = A[M][K pixels] (u8, shared memory) x Hk[N][K] (s8, +-1)
//
// 8x8 tiles: K = N = 64, Hk = H8 (x) H8 (two K = 32 steps); 4x4-tiled shapes: a lane unit is a pair of 4x4 tiles, K = N = 32,
// Hk = diag(H4 (x) H4, H4 (x) H4) (one step).  kind::i8 takes the pixels as they are (u8) and accumulates in s32: exact.
// The residual needs no subtraction pass: the source tile is parked once per pack as its one's complement and accumulated
// into the same D with the same Hk,  Hk (c + 255 - o) = Hk (c - o) + 255 * K * e0,  and the constant leaves coefficient 0
// in the epilogue.  A CTA is four worker warps -- warp w owns rows 32 w .. 32 w + 31 of A and the same TMEM lanes of D (a
// warp can only read its own 32 TMEM lanes, which is what ties four packs into one CTA), one pack each, lane = unit exactly
// as in k2_pack -- plus an issuer warp whose lane 0 waits for the four warps' rows (mbarrier), issues the MMAs and commits
// them onto a second mbarrier.  Per candidate a worker lane copies its tile's 64 bytes from the staged region into the
// canonical K-major operand layout (row m, 16-byte chunk kc at 16 m + 2048 kc: SWIZZLE_NONE core matrices of 8 rows x 16
// bytes, SBO 128, LBO 2048), and one round later -- while the next candidate's MMAs run -- reads its row of D with
// tcgen05.ld and sums the 64 |coefficients|.  UM_DEPTH candidate buffers / accumulators / mbarrier pairs per CTA, three
// CTAs per SM.  Layout, descriptors and arithmetic were pinned first in tools/proto_umma_satd.cu.
//
// Measured (profiles/r2_k2_umma.txt): bit-exact over the whole GPU suite, 1.20 ms per 1080p frame against 0.74 ms for the
// SWAR kernel.  The transform does leave the issue stream, but an SM runs 3 independent instruction streams here instead
// of 12, and the scalar search code around the transform (staging, bookkeeping, reductions) issues at one instruction per
// ~10 clk; the path is selectable (fme_config.k2Path), FME_K2_PATH_AUTO stays on SWAR.
//
// Served here: uni-prediction records with Hadamard distortion (lossless PUs inside such packs get their SAD from the
// staged region, TEncSearch.cpp:5258).  SAD mode and bi-predictive records stay on k2_refine.cu.
#include "k2_common.cuh"

namespace {

constexpr int UM_WORKERS = 4;                       // worker warps = rows of one M = 128 UMMA / 32
constexpr int UM_THREADS = (UM_WORKERS + 1) * 32;   // + the issuer warp
constexpr int UM_A_BYTES = 128 * 64;                // one operand buffer: 128 rows x 64 bytes
constexpr int UM_OFF_B8 = 0;                        // H8 (x) H8, 64 x 64 s8
constexpr int UM_OFF_B4 = 4096;                     // diag(H4 (x) H4, H4 (x) H4), 32 x 32 s8
#ifndef FME_UM_DEPTH
#define FME_UM_DEPTH 2
#endif
constexpr int UM_DEPTH = FME_UM_DEPTH;              // rounds in flight per CTA: candidate buffers, accumulators, mbarrier pairs
static_assert(UM_DEPTH == 2 || UM_DEPTH == 4, "UM_DEPTH must be 2 or 4 (TMEM allocations are powers of two)");
constexpr int UM_OFF_A = 5120;                      // UM_DEPTH candidate buffers (round r uses buffer r % UM_DEPTH)
constexpr int UM_OFF_ORG = UM_OFF_A + UM_DEPTH * UM_A_BYTES;  // complemented source tiles (units 0..31 of every PU)
// The second source tile of the PUs with more than 32 units (64x64, 64x48, 48x64) is parked in candidate buffer 1; packs of
// those shapes run one round at a time through candidate buffer 0 (the other CTAs of the SM cover the latency).
constexpr int UM_OFF_STAGE = UM_OFF_ORG + UM_A_BYTES;      // per worker warp: the two staging buffers of k2_pack
constexpr int UM_SMEM = UM_OFF_STAGE + UM_WORKERS * K2_STAGE_BYTES;
constexpr int UM_CTAS_PER_SM = UM_DEPTH == 2 ? 3 : 2;   // by shared memory (3 x 77 KB) and TMEM columns
constexpr int UM_TMEM_COLS = 64 * UM_DEPTH;         // UM_DEPTH accumulators of 64 columns
// round command: what the issuing thread multiplies (candidate buffer = round % UM_DEPTH unless _BUF0; source rows from
// the source area unless _ORG_IN_BUF1)
constexpr unsigned UM_CMD_QUIT = 1u, UM_CMD_TS4 = 2u, UM_CMD_BUF0 = 4u, UM_CMD_ORG_IN_BUF1 = 8u;

__device__ __forceinline__ uint64_t um_desc(unsigned addr, unsigned lboBytes, unsigned sboBytes) {
  // SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes; LBO = distance between the
  // two 16-byte K chunks of one K = 32 step, SBO = distance between 8-row groups; descriptor version 1 (bits 46-47)
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lboBytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sboBytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(unsigned tmemD, uint64_t descA, uint64_t descB, unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, int (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                 "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                 "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void um_mbar_wait(unsigned mbar, unsigned parity) {
  unsigned done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    if (spin > (1 << 24)) __trap();   // a lost arrival must not hang the device
  }
}
__device__ __forceinline__ void sts128(unsigned sa, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ unsigned lds32(unsigned sa) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa));
  return v;
}

// Optional phase timers (-DFME_UM_PROF, tools/um_prof.py): clock64 deltas of lane 0 of every worker warp, summed per phase.
#ifdef FME_UM_PROF
__device__ unsigned long long g_umProf[16];
#define UM_PROF_DECL long long umT = clock64();
#define UM_PROF(k) { const long long umN = clock64(); um.prof[k] += umN - umT; umT = umN; }
#else
#define UM_PROF_DECL
#define UM_PROF(k)
#endif
enum { UMP_SETUP, UMP_STAGE_ISSUE, UMP_STAGE_WAIT, UMP_DONE_WAIT, UMP_COLLECT, UMP_BOOK, UMP_REPACK, UMP_SUBMIT, UMP_SCHED, UMP_TOTAL, UMP_N };

// What a worker warp needs of its CTA's tensor-path state.
struct UmCtx {
  unsigned aAddr;     // shared address of candidate buffer 0 (+ UM_A_BYTES: buffer 1)
  unsigned orgAddr;   // shared address of the source rows
  unsigned full0;     // mbarrier "rows of round r written" (one arrival per worker warp), + 8 * (r % UM_DEPTH)
  unsigned done0;     // mbarrier "MMAs of round r complete" (tcgen05.commit), + 8 * (r % UM_DEPTH)
  unsigned tmem;      // TMEM base address of the CTA's 128 columns
  volatile unsigned* cmd;  // [UM_DEPTH] what the issuer runs for round r % UM_DEPTH (written by warp 0)
  unsigned round;     // rounds submitted so far (identical in all worker warps and in the issuer)
  unsigned* sadFifo;  // [UM_DEPTH][128] per-thread SAD of the rounds in flight (lossless PUs)
#ifdef FME_UM_PROF
  long long prof[UMP_N];
#endif
};

// This thread's operand rows of the current round are in shared memory: make them visible to the async proxy and count
// the warp in (the issuer warp's lane 0 waits for the four arrivals).  A dedicated issuer beats lane 0 of warp 0 doing
// the issue itself (1.21 against 1.31 ms per 1080p frame): the worker goes straight on to the next candidate.
__device__ __forceinline__ void um_submit(UmCtx& um, unsigned cmd, int warp, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const unsigned r = um.round++;
  if (lane == 0) {
    if (warp == 0) um.cmd[r % UM_DEPTH] = cmd;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(um.full0 + 8 * (r % UM_DEPTH)) : "memory");
  }
}

// Rounded SATD of this thread's row of round r: (sum |coef| + 2) >> 2 for an 8x8 tile (TComRdCost.cpp:1421),
// (sum + 1) >> 1 per 4x4 tile (:1325), summed over the pair.
template <int TS>
__device__ __forceinline__ unsigned um_collect(UmCtx& um, unsigned r, int warp) {
  UM_PROF_DECL
  um_mbar_wait(um.done0 + 8 * (r % UM_DEPTH), (r / UM_DEPTH) & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  UM_PROF(UMP_DONE_WAIT)
  const unsigned ta = um.tmem + ((unsigned)(32 * warp) << 16) + (r % UM_DEPTH) * 64;
  unsigned s = 0;
  if constexpr (TS == 8) {
    int v0[32], v1[32];
    tmem_ld32(ta, v0);
    tmem_ld32(ta + 32, v1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v0[0] -= 255 * 64;
    unsigned s1 = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) { s += (unsigned)abs(v0[j]); s1 += (unsigned)abs(v1[j]); }
    s = (s + s1 + 2) >> 2;
  } else {
    int v[32];
    tmem_ld32(ta, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] -= 255 * 16;
    v[16] -= 255 * 16;
    unsigned s1 = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s += (unsigned)abs(v[j]); s1 += (unsigned)abs(v[16 + j]); }
    s = ((s + 1) >> 1) + ((s1 + 1) >> 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // this accumulator is free once every warp has counted in again
  UM_PROF(UMP_COLLECT)
  return s;
}

// One 8x8 candidate tile from a staged region (shared address of its first byte, row pitch a multiple of 4) into this
// thread's operand row: chunk kc = tile rows 2 kc, 2 kc + 1.  Rows are read as aligned words and funnel-shifted.
__device__ __forceinline__ void um_repack8(unsigned cand, int pitch, unsigned dst) {
  const unsigned sh = (cand & 3u) * 8u;
  unsigned p = cand & ~3u;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    const unsigned w0 = lds32(p), w1 = lds32(p + 4), w2 = lds32(p + 8);
    p += pitch;
    const unsigned x0 = lds32(p), x1 = lds32(p + 4), x2 = lds32(p + 8);
    p += pitch;
    sts128(dst + kc * 2048, __funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(x0, x1, sh),
           __funnelshift_r(x1, x2, sh));
  }
}
// A 4x4 candidate tile -> one 16-byte chunk.
__device__ __forceinline__ void um_repack4(unsigned cand, int pitch, unsigned dst) {
  const unsigned sh = (cand & 3u) * 8u;
  unsigned p = cand & ~3u;
  unsigned v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    v[r] = __funnelshift_r(lds32(p), lds32(p + 4), sh);
    p += pitch;
  }
  sts128(dst, v[0], v[1], v[2], v[3]);
}

// One pack on worker warp `warp` (count <= 0: the warp has no PUs in this CTA pack and only keeps the rounds in step).
template <int TS, int A>
__device__ __noinline__ void k2_pack_umma(const ClassInfo ci, const int* __restrict__ order, int first, int count,
                                          const fme_pu* __restrict__ pus, fme_result* __restrict__ res,
                                          const uint8_t* __restrict__ planes, const uint8_t* __restrict__ org,
                                          const FmeGeom& g, const uint32_t* __restrict__ costLut, uint8_t* smem, UmCtx& um,
                                          int warp) {
  UM_PROF_DECL
  const int lane = threadIdx.x & 31;
  const int w = ci.w, h = ci.h;
  const int U = ci.units;
  const bool twoUnits = TS == 8 && A == 16 && U > 32;   // 64x64, 64x48, 48x64: every lane also serves unit lane + 32
  const int gPitch = g.pitch, gOrgPitch = g.orgPitch;
  const int gPlaneBytes = (int)g.planeBytes;
  const int lanesPerPu = ci.lanes;
  const int myPu = lane / lanesPerPu;
  const bool laneActive = myPu < count;
  const int unit0 = lane - myPu * lanesPerPu;

  // ---- this lane's PU (as in k2_pack) ----
  int predX = 0, predY = 0, mvIntX = 0, mvIntY = 0, lossless = 0, puIdx = -1, alignX = 0;
  int ox = 0, oy = 0;
  long long slotOff = 0;
  if (laneActive) {
    puIdx = order[first + myPu];
    const fme_pu p = pus[puIdx];
    predX = p.mvPredX; predY = p.mvPredY; mvIntX = p.mvIntX; mvIntY = p.mvIntY;
    lossless = p.flags & FME_PU_LOSSLESS;
    const int X = min(max(p.x + p.mvIntX, -(g.M - 8)), g.W + g.M - 8 - w);
    const int Y = min(max(p.y + p.mvIntY, -(g.M - 8)), g.H + g.M - 8 - h);
    alignX = X + g.M;
    slotOff = (long long)min((int)p.refSlot, g.numSlots - 1) * (long long)g.slotBytes +
              (long long)((Y + g.M) * gPitch + alignX);
    ox = min(max((int)p.x, 0), g.W - w);
    oy = min(max((int)p.y, 0), g.H - h);
  }
  const bool had = !lossless;
  const bool anyLossless = __any_sync(0xffffffffu, laneActive && lossless);

  // ---- staging geometry (as in k2_pack) ----
  StageGeom sg;
  sg.RW = ((w + A + A - 1) / A) * A;
  sg.G = sg.RW / A;
  sg.RB = (h + 1) * sg.RW;
  if (((sg.RB / A) & 1) == 0) sg.RB += A;
  const int bufBytes = (ci.P * sg.RB + 16 + 15) & ~15;
  uint8_t* const bufA = smem;
  uint8_t* const bufB = smem + bufBytes;
  int gpShift = 0;
  while ((1 << gpShift) < sg.G && (1 << gpShift) < lanesPerPu) ++gpShift;
  const int Gp = 1 << gpShift;
  const int stGi = unit0 & (Gp - 1), stRowSub = unit0 >> gpShift, stRowStep = lanesPerPu >> gpShift;
  const bool stSecond = Gp < sg.G;
  const bool stOn = laneActive && stGi < sg.G;
  const unsigned stDst = (unsigned)__cvta_generic_to_shared(smem) + myPu * sg.RB + stRowSub * sg.RW + stGi * A;
  const int stSrcOff = stRowSub * gPitch + stGi * A;

  int bhx = 0, bhy = 0;
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
      const long long off = slotOff + (long long)(plane * gPlaneBytes + dy * gPitch + dx);
      stage_rows<A>(stDst + ((s & 1) ? bufBytes : 0), planes + (off & ~(long long)(A - 1)) + stSrcOff, stRowSub, stRowStep,
                    s < 4 ? h + 1 : h, sg.RW, gPitch, stSecond);
    }
    cp_async_commit();
  };
  stage(0);

  // ---- source tile(s) of this lane: one's complement into the source slot(s), operand layout ----
  const unsigned rowOff = 16u * (unsigned)(warp * 32 + lane);
  int uOff = 0, u1Off = 0, uOff2 = 0;
  bool uOn = laneActive && unit0 < U, uOn2 = false;
  if (uOn) {
    const unsigned dst = um.orgAddr + rowOff;
    if constexpr (TS == 8) {
      const int ty = unit0 / ci.tilesX, tx = unit0 - ty * ci.tilesX;
      uOff = ty * 8 * sg.RW + tx * 8;
      const uint8_t* src = org + (size_t)(oy + ty * 8) * gOrgPitch + ox + tx * 8;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        unsigned a0, a1, b0, b1;
        ldg_row8(src + (size_t)(2 * kc) * gOrgPitch, a0, a1);
        ldg_row8(src + (size_t)(2 * kc + 1) * gOrgPitch, b0, b1);
        sts128(dst + kc * 2048, ~a0, ~a1, ~b0, ~b1);
      }
    } else {
      const int t0 = 2 * unit0, t1 = 2 * unit0 + 1;
      const int ty0 = t0 / ci.tilesX, tx0 = t0 - ty0 * ci.tilesX;
      const int ty1 = t1 / ci.tilesX, tx1 = t1 - ty1 * ci.tilesX;
      uOff = ty0 * 4 * sg.RW + tx0 * 4;
      u1Off = (ty1 * 4 * sg.RW + tx1 * 4) - uOff;
      const uint8_t* s0 = org + (size_t)(oy + ty0 * 4) * gOrgPitch + ox + tx0 * 4;
      const uint8_t* s1 = org + (size_t)(oy + ty1 * 4) * gOrgPitch + ox + tx1 * 4;
      sts128(dst, ~ldg_row4(s0), ~ldg_row4(s0 + gOrgPitch), ~ldg_row4(s0 + 2 * (size_t)gOrgPitch), ~ldg_row4(s0 + 3 * (size_t)gOrgPitch));
      sts128(dst + 2048, ~ldg_row4(s1), ~ldg_row4(s1 + gOrgPitch), ~ldg_row4(s1 + 2 * (size_t)gOrgPitch), ~ldg_row4(s1 + 3 * (size_t)gOrgPitch));
    }
  }
  if constexpr (TS == 8 && A == 16) {
    if (twoUnits) {
      const int u = unit0 + 32;
      uOn2 = laneActive && u < U;
      if (uOn2) {
        const int ty = u / ci.tilesX, tx = u - ty * ci.tilesX;
        uOff2 = ty * 8 * sg.RW + tx * 8;
        const uint8_t* src = org + (size_t)(oy + ty * 8) * gOrgPitch + ox + tx * 8;
        const unsigned dst = um.aAddr + UM_A_BYTES + rowOff;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          unsigned a0, a1, b0, b1;
          ldg_row8(src + (size_t)(2 * kc) * gOrgPitch, a0, a1);
          ldg_row8(src + (size_t)(2 * kc + 1) * gOrgPitch, b0, b1);
          sts128(dst + kc * 2048, ~a0, ~a1, ~b0, ~b1);
        }
      }
    }
  }
  // SAD of this lane's unit(s) against a candidate, for lossless PUs (source rows read back from the slots)
  auto laneSad = [&](const uint8_t* regionCand) -> unsigned {
    unsigned d = 0;
    if constexpr (TS == 8) {
      if (uOn) {
        const unsigned so = um.orgAddr + rowOff;
        auto row = [&](int r, unsigned& lo, unsigned& hi) {
          lo = ~lds32(so + (r >> 1) * 2048 + (r & 1) * 8);
          hi = ~lds32(so + (r >> 1) * 2048 + (r & 1) * 8 + 4);
        };
        d = sad8x8(row, regionCand + uOff, sg.RW);
      }
      if (uOn2) {
        const unsigned so = um.aAddr + UM_A_BYTES + rowOff;
        auto row = [&](int r, unsigned& lo, unsigned& hi) {
          lo = ~lds32(so + (r >> 1) * 2048 + (r & 1) * 8);
          hi = ~lds32(so + (r >> 1) * 2048 + (r & 1) * 8 + 4);
        };
        d += sad8x8(row, regionCand + uOff2, sg.RW);
      }
    } else if (uOn) {
      const unsigned so = um.orgAddr + rowOff;
      unsigned oa[4], ob[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { oa[r] = ~lds32(so + 4 * r); ob[r] = ~lds32(so + 2048 + 4 * r); }
      d = sad4x4(oa, regionCand + uOff, sg.RW) + sad4x4(ob, regionCand + uOff + u1Off, sg.RW);
    }
    return d;
  };

  // ---- 12 staging steps; every candidate is one round (two for the big PUs), retired one round later ----
  unsigned hBest = 0xffffffffu, qBest = 0xffffffffu;
  int hBestI = 9, qBestI = 0;
  unsigned bitsX = 0, bitsY = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    bitsX |= (unsigned)golomb_bits((((mvIntX << 1) + (t - 1)) << 1) - predX) << (8 * t);
    bitsY |= (unsigned)golomb_bits((((mvIntY << 1) + (t - 1)) << 1) - predY) << (8 * t);
  }
  // Rounds are retired in submission order, up to UM_DEPTH - 1 rounds after their submission.  What a retiring round was
  // follows from its sequence number in the pack: candidate q (half-pel candidates in staging order, then Q1..Q8), and for
  // the big PUs the unit (even rounds: units 0..31, odd rounds: units 32..).  The MV bits are recomputed at retirement
  // (the per-axis bit counts of the half-pel stage are replaced only after the stage has drained).
  int sub = 0, ret = 0;                 // rounds of this pack submitted / retired
  const unsigned round0 = um.round;     // CTA round index of the pack's first round
  unsigned pPartial = 0;
  unsigned* const sadFifo = um.sadFifo + (warp * 32 + lane);   // [UM_DEPTH][128]: SAD of lossless PUs' units per round
  auto retire = [&]() {
    const int q = twoUnits ? ret >> 1 : ret;
    const bool second = twoUnits && (ret & 1);
    const unsigned d = um_collect<TS>(um, round0 + ret, warp);
#ifdef FME_UM_PROF
    umT = clock64();
#endif
    const int slot = ret % UM_DEPTH;
    ++ret;
    if (twoUnits && !second) {
      pPartial = uOn ? d : 0u;
      return;
    }
    const bool half = q < 9;
    const int i = c_seqI[q];
    const unsigned sel = c_seqSel[q];
    const int bits = (int)(__byte_perm(bitsX, 0, sel) + __byte_perm(bitsY, 0, sel >> 16));
    unsigned dist = second ? pPartial + (uOn2 ? d : 0u) : (uOn ? d : 0u);
    if (anyLossless && !had) dist = sadFifo[slot * 128];   // lossless PUs: SAD (TEncSearch.cpp:5258)
    if (lanesPerPu == 32) dist = __reduce_add_sync(0xffffffffu, dist);
    else for (int dd = lanesPerPu >> 1; dd > 0; dd >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, dd);
    if (laneActive) {
      dist += costLut[bits];
      if (half) {
        if (dist < hBest || (dist == hBest && i < hBestI)) { hBest = dist; hBestI = i; }
      } else if (dist < qBest) {
        qBest = dist;
        qBestI = i;
      }
    }
    UM_PROF(UMP_BOOK)
  };
  // submit the rows this thread has just written as the pack's next round
  auto submit = [&](unsigned cmd, unsigned sadNow) {
    if (anyLossless) sadFifo[(sub % UM_DEPTH) * 128] = sadNow;   // read back by this thread only
    UM_PROF(UMP_REPACK)
    um_submit(um, cmd, warp, lane);
    ++sub;
    UM_PROF(UMP_SUBMIT)
  };
  // before writing the rows of a new round: its buffer's previous round (UM_DEPTH rounds ago) must have been retired
  const int depth = twoUnits ? 1 : UM_DEPTH;
  auto makeRoom = [&]() {
    UM_PROF(UMP_BOOK)
    if (sub - ret == depth) retire();
  };

  // shared address of this lane's unit inside its PU's region of staging buffer A; half-pel candidates sit at (1, 1) of
  // the region (origin X - 1, Y - 1), minus one sample / one row for the offsets -1
  const unsigned regionSA0 = (unsigned)__cvta_generic_to_shared(smem) + myPu * sg.RB + uOff;
  const int halfOff = sg.RW + ((alignX - 1) & (A - 1)) + 1;
  int cq = 0;   // sequence number of the next candidate
  UM_PROF(UMP_SETUP)
#pragma unroll 1
  for (int s = 0; s < 12; ++s) {
    const bool prefetch = (s != 3) && (s != 11);
    if (prefetch) {
      stage(s + 1);
      UM_PROF(UMP_STAGE_ISSUE)
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    UM_PROF(UMP_STAGE_WAIT)
    const unsigned regionSA = regionSA0 + ((s & 1) ? bufBytes : 0);   // shared address of this lane's unit in the region
    const int iCount = c_stepCount[s];
#pragma unroll 1
    for (int c = 0; c < iCount; ++c, ++cq) {
      int candOff;   // candidate's first byte relative to the region's
      if (s < 4) {
        const unsigned sel = c_seqSel[cq];   // ox + 1 in bits 0-1, oy + 1 in bits 16-17; offsets -1 reach one sample back
        candOff = halfOff - ((sel & 3u) == 0 ? 1 : 0) - ((sel & 0x30000u) == 0 ? sg.RW : 0);
      } else {
        const int qx = 2 * bhx + (int)(c_seqSel[cq] & 3u) - 1;
        candOff = (alignX + (qx >> 2)) & (A - 1);
      }
      makeRoom();
      if (uOn) {
        const unsigned dst = um.aAddr + (twoUnits ? 0u : (um.round % UM_DEPTH) * UM_A_BYTES) + rowOff;
        if constexpr (TS == 8) um_repack8(regionSA + candOff, sg.RW, dst);
        else {
          um_repack4(regionSA + candOff, sg.RW, dst);
          um_repack4(regionSA + candOff + u1Off, sg.RW, dst + 2048);
        }
      }
      unsigned sadNow = 0;
      if (anyLossless && laneActive && !had) sadNow = laneSad(((s & 1) ? bufB : bufA) + myPu * sg.RB + candOff);
      submit(TS == 4 ? UM_CMD_TS4 : twoUnits ? UM_CMD_BUF0 : 0u, sadNow);
      if constexpr (TS == 8 && A == 16) {
        if (twoUnits) {
          makeRoom();
          if (uOn2) um_repack8(regionSA + candOff + (uOff2 - uOff), sg.RW, um.aAddr + rowOff);
          submit(UM_CMD_BUF0 | UM_CMD_ORG_IN_BUF1, sadNow);
        }
      }
    }
    __syncwarp();  // all lanes are done with this buffer before step s+2 overwrites it
    UM_PROF(UMP_BOOK)
    if (s == 3) {
      while (ret < sub) retire();   // the half-pel winner needs every half-pel candidate
      const int bestI = hBestI < 9 ? hBestI : 0;
      bhx = c_refineH[bestI][0];
      bhy = c_refineH[bestI][1];
      qBest = hBest;
      qBestI = 0;
      stage(4);
      UM_PROF(UMP_STAGE_ISSUE)
      bitsX = bitsY = 0;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        bitsX |= (unsigned)golomb_bits((((mvIntX << 1) + bhx) << 1) + (t - 1) - predX) << (8 * t);
        bitsY |= (unsigned)golomb_bits((((mvIntY << 1) + bhy) << 1) + (t - 1) - predY) << (8 * t);
      }
    }
  }
  while (ret < sub) retire();

  if (laneActive && unit0 == 0) {
    fme_result* r = &res[puIdx];
    r->halfX = (int8_t)bhx; r->halfY = (int8_t)bhy;
    r->qterX = c_refineQ[qBestI][0]; r->qterY = c_refineQ[qBestI][1];
    r->cost = qBest;
  }
  __syncwarp();
  UM_PROF(UMP_BOOK)
}

__global__ void __launch_bounds__(UM_THREADS, UM_CTAS_PER_SM)
k2_refine_umma(const fme_pu* __restrict__ pus, fme_result* __restrict__ res, const uint8_t* __restrict__ planes,
               const uint8_t* __restrict__ org, const FmeGeom g, const FmeCostLut costLutG, const int* __restrict__ order,
               const int* __restrict__ classOffset, const int* __restrict__ packOffset, int* __restrict__ workCounter) {
  extern __shared__ __align__(128) uint8_t dynSmem[];
  __shared__ uint32_t s_lut[FME_COST_LUT_SIZE];
  __shared__ int s_packOff[FME_K2_KEYS + 1];
  __shared__ __align__(8) unsigned long long s_full[UM_DEPTH], s_done[UM_DEPTH];
  __shared__ unsigned s_cmd[UM_DEPTH];
  __shared__ unsigned s_sad[UM_DEPTH * 128];
  __shared__ unsigned s_tmem;
  __shared__ int s_next[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < FME_COST_LUT_SIZE; i += UM_THREADS) s_lut[i] = costLutG.v[i];
  for (int i = tid; i <= FME_K2_KEYS; i += UM_THREADS) s_packOff[i] = packOffset[i];
  // transform matrices in the canonical layout: byte i -> (n, k)
  for (int i = tid; i < 4096; i += UM_THREADS) {
    const int kc = i >> 10, n = ((i & 1023) >> 7) * 8 + ((i & 127) >> 4), k = kc * 16 + (i & 15);
    dynSmem[UM_OFF_B8 + i] = (uint8_t)(int8_t)(((__popc((n >> 3) & (k >> 3)) + __popc(n & k & 7)) & 1) ? -1 : 1);
  }
  for (int i = tid; i < 1024; i += UM_THREADS) {
    const int kc = i >> 9, n = ((i & 511) >> 7) * 8 + ((i & 127) >> 4), k = kc * 16 + (i & 15);
    int v = 0;
    if ((n >> 4) == (k >> 4)) v = ((__popc((n >> 2) & (k >> 2) & 3) + __popc(n & k & 3)) & 1) ? -1 : 1;
    dynSmem[UM_OFF_B4 + i] = (uint8_t)(int8_t)v;
  }
  const unsigned full0 = (unsigned)__cvta_generic_to_shared(&s_full[0]), done0 = (unsigned)__cvta_generic_to_shared(&s_done[0]);
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < UM_DEPTH; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * b), "r"(UM_WORKERS));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done0 + 8 * b));
    }
    s_next[0] = atomicAdd(workCounter, 1);
  }
  if (warp == UM_WORKERS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "r"(UM_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;
  const unsigned base = (unsigned)__cvta_generic_to_shared(dynSmem);

  if (warp == UM_WORKERS) {
    // ---- issuer: one thread turns every completed round of rows into MMAs ----
    if (lane == 0) {
      constexpr unsigned IDESC8 = (2u << 4) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // u8 x s8 -> s32, K-major, M 128, N 64
      constexpr unsigned IDESC4 = (2u << 4) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);  // N 32
      const uint64_t b8lo = um_desc(base + UM_OFF_B8, 1024, 128), b8hi = um_desc(base + UM_OFF_B8 + 2048, 1024, 128);
      const uint64_t b4 = um_desc(base + UM_OFF_B4, 512, 128);
      for (unsigned r = 0;; ++r) {
        const unsigned b = r % UM_DEPTH;
        um_mbar_wait(full0 + 8 * b, (r / UM_DEPTH) & 1);
        const unsigned cmd = *reinterpret_cast<volatile unsigned*>(&s_cmd[b]);
        if (cmd & UM_CMD_QUIT) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned d = tmem + b * 64;
        const unsigned a = base + UM_OFF_A + ((cmd & UM_CMD_BUF0) ? 0u : b * UM_A_BYTES);
        const unsigned o = (cmd & UM_CMD_ORG_IN_BUF1) ? base + UM_OFF_A + UM_A_BYTES : base + UM_OFF_ORG;
        if (!(cmd & UM_CMD_TS4)) {
          umma_i8(d, um_desc(a, 2048, 128), b8lo, IDESC8, 0);
          umma_i8(d, um_desc(a + 4096, 2048, 128), b8hi, IDESC8, 1);
          umma_i8(d, um_desc(o, 2048, 128), b8lo, IDESC8, 1);
          umma_i8(d, um_desc(o + 4096, 2048, 128), b8hi, IDESC8, 1);
        } else {
          umma_i8(d, um_desc(a, 2048, 128), b4, IDESC4, 0);
          umma_i8(d, um_desc(o, 2048, 128), b4, IDESC4, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done0 + 8 * b) : "memory");
      }
    }
  } else {
    // ---- workers: a CTA pack = four warp packs of one (slot group, shape class) key ----
    UmCtx um;
    um.aAddr = base + UM_OFF_A;
    um.orgAddr = base + UM_OFF_ORG;
    um.full0 = full0;
    um.done0 = done0;
    um.tmem = tmem;
    um.cmd = s_cmd;
    um.round = 0;
    um.sadFifo = s_sad;
#ifdef FME_UM_PROF
    for (int k = 0; k < UMP_N; ++k) um.prof[k] = 0;
    const long long umStart = clock64();
    long long umT = umStart;
#endif
    uint8_t* const smem = dynSmem + UM_OFF_STAGE + warp * K2_STAGE_BYTES;
    const int totalPacks = s_packOff[FME_K2_KEYS];
    for (int it = 0;; ++it) {
      asm volatile("bar.sync 1, %0;" ::"n"(UM_WORKERS * 32) : "memory");   // s_next[it & 1] is published; the previous pack is done
      const int pack = s_next[it & 1];
      if (pack >= totalPacks) break;
      if (tid == 0) s_next[(it + 1) & 1] = atomicAdd(workCounter, 1);
      int v = 0;
      {
        int hi = FME_K2_KEYS;
        while (hi - v > 1) {
          const int mid = (v + hi) >> 1;
          if (s_packOff[mid] <= pack) v = mid; else hi = mid;
        }
      }
      const ClassInfo ci = class_info(63 - (v & 63));
      const int classOff = __ldg(classOffset + v);   // (two loads per pack: the table stays out of shared memory)
      const int inClass = __ldg(classOffset + v + 1) - classOff;
      const int ctaFirst = (pack - s_packOff[v]) * 4 * ci.P;
      const int first = classOff + ctaFirst + warp * ci.P;
      const int count = min(ci.P, inClass - ctaFirst - warp * ci.P);   // <= 0: nothing for this warp
      UM_PROF(UMP_SCHED)
#define K2U_ARGS ci, order, first, count, pus, res, planes, org, g, s_lut, smem, um, warp
      if (ci.ts == 8) {
        if (ci.w >= 16) k2_pack_umma<8, 16>(K2U_ARGS);
        else k2_pack_umma<8, 8>(K2U_ARGS);
      } else {
        if (ci.w >= 16) k2_pack_umma<4, 16>(K2U_ARGS);
        else if (ci.w >= 8) k2_pack_umma<4, 8>(K2U_ARGS);
        else k2_pack_umma<4, 4>(K2U_ARGS);
      }
#undef K2U_ARGS
#ifdef FME_UM_PROF
      umT = clock64();
#endif
    }
    um_submit(um, UM_CMD_QUIT, warp, lane);
#ifdef FME_UM_PROF
    um.prof[UMP_TOTAL] = clock64() - umStart;
    if (lane == 0) for (int k = 0; k < UMP_N; ++k) atomicAdd(&g_umProf[k], (unsigned long long)um.prof[k]);
#endif
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == UM_WORKERS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(UM_TMEM_COLS));
  }
}

}  // namespace

cudaError_t fme_launch_k2_umma(const FmeGeom& g, const uint8_t* d_planes, const uint8_t* d_org, const fme_pu* d_pus, int n,
                               fme_result* d_res, const FmeCostLut& costLut, int biServed, const FmeK2Scratch& sc, int numSMs,
                               cudaStream_t s, int64_t* launches) {
  cudaError_t e = fme_k2_bin(d_pus, n, d_res, 0, biServed, sc, 2, numSMs, s, launches);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k2_refine_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k2_refine_umma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  // UM_CTAS_PER_SM CTAs per SM; the pack hand-out is dynamic, so a CTA that starts late just finds less work
  const int perSM = UM_CTAS_PER_SM;
  k2_refine_umma<<<numSMs * perSM, UM_THREADS, UM_SMEM, s>>>(d_pus, d_res, d_planes, d_org, g, costLut, sc.order, sc.classOffset,
                                                             sc.packOffset, sc.workCounter);
  ++*launches;
  return cudaGetLastError();
}

#ifdef FME_UM_PROF
// Debug build only: phase timers of k2_refine_umma summed over the worker warps since the last call (then reset).
extern "C" int fme_debug_um_prof(unsigned long long* out, int n) {
  unsigned long long h[16] = {0}, z[16] = {0};
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(h, g_umProf, sizeof(h)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_umProf, z, sizeof(z));
  for (int i = 0; i < n && i < 16; ++i) out[i] = h[i];
  return UMP_N;
}
#endif
