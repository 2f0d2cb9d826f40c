// Pieces of the NN_pred pass shared by k3_nn.cu (the stand-alone K3 kernels) and k2_refine.cu (K3's work items fused into
// the persistent K2 kernel): embedding index map, class -> sub-pel vector map, the register-blocked dense layer and the
// shared-memory image of a layer.  Internal linkage.
#pragma once
#include "fme_common.cuh"

namespace {

__device__ __forceinline__ int emb_index(int v, bool isHeight) {
  switch (v) {
    case 4: return 1;
    case 8: return 2;
    case 16: return isHeight ? 3 : 4;  // TEncSearch.cpp:96 vs :108
    case 12: return isHeight ? 4 : 3;  // TEncSearch.cpp:97 vs :107
    case 24: return 5;
    case 32: return 6;
    case 64: return 7;
    default: return 0;
  }
}

// TEncSearch.cpp:136-193: class -> per-axis (half, quarter)
__device__ __forceinline__ void store_class(fme_result* r, int best) {
  int qx = best % 7, qy = best / 7;
  const int kHalf = (0 << 0) | (0 << 2) | (1 << 4) | (1 << 6) | (1 << 8) | (2 << 10) | (2 << 12);  // +1 biased
  const int kQter = (0 << 0) | (1 << 2) | (0 << 4) | (1 << 6) | (2 << 8) | (1 << 10) | (2 << 12);  // +1 biased
  int8_t hx = 0, hy = 0, tx = 0, ty = 0;
  if (best >= 0 && best <= 48) {
    hx = (int8_t)(((kHalf >> (2 * qx)) & 3) - 1); tx = (int8_t)(((kQter >> (2 * qx)) & 3) - 1);
    hy = (int8_t)(((kHalf >> (2 * qy)) & 3) - 1); ty = (int8_t)(((kQter >> (2 * qy)) & 3) - 1);
  }
  r->nnHalfX = hx; r->nnHalfY = hy; r->nnQterX = tx; r->nnQterY = ty;
  r->nnClass = (uint8_t)best;
}

constexpr int K3F_THREADS = 128;
constexpr int K3F_NPU = 2;  // PUs per thread

__host__ __device__ constexpr int pad4(int v) { return (v + 3) & ~3; }

// One dense layer for NPU input vectors held in registers.  sW: rows padded to pad4(IN) floats, 16-byte aligned.
// emit(o, acc[NPU]) receives the pre-activation W x + b of unit o (ascending-k, mul then add).
// UO = output units in flight (unroll factor of the unit loop): 4 in the stand-alone kernel (measured best of 1 / 2 / 4 / 7 /
// 10); 1 where code size matters more than the last bit of instruction-level parallelism (the items fused into K2).
template <bool FMA, int IN, int OUT, int UO = 4, typename Emit>
__device__ __forceinline__ void dense_rows(const float* __restrict__ sW, const float* __restrict__ sb,
                                           const float (&x)[K3F_NPU][IN], Emit emit) {
  constexpr int IN4 = pad4(IN);
#pragma unroll UO
  for (int o = 0; o < OUT; ++o) {
    const float4* wr = reinterpret_cast<const float4*>(sW + o * IN4);
    float acc[K3F_NPU];
#pragma unroll
    for (int q = 0; q < IN4 / 4; ++q) {
      const float4 wv = wr[q];  // broadcast LDS.128
      const float wk[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = 4 * q + t;
        if (k < IN) {
#pragma unroll
          for (int u = 0; u < K3F_NPU; ++u) {
            if (FMA) {  // opt-in relaxed mode (fme_config.nnFma): one rounding per tap
              acc[u] = (k == 0) ? __fmul_rn(wk[t], x[u][k]) : __fmaf_rn(wk[t], x[u][k], acc[u]);
            } else {
              const float prod = __fmul_rn(wk[t], x[u][k]);
              acc[u] = (k == 0) ? prod : __fadd_rn(acc[u], prod);
            }
          }
        }
      }
    }
    const float bo = sb[o];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u) acc[u] = __fadd_rn(acc[u], bo);
    emit(o, acc);
  }
}

// shared-memory image of one layer
template <int IN, int OUT>
struct LayerSmem {
  static constexpr int IN4 = pad4(IN);
  static constexpr int WORDS = OUT * IN4 + 3 * pad4(OUT);
  __device__ static const float* W(const float* s) { return s; }
  __device__ static const float* b(const float* s) { return s + OUT * IN4; }
  __device__ static const float* g(const float* s) { return s + OUT * IN4 + pad4(OUT); }
  __device__ static const float* be(const float* s) { return s + OUT * IN4 + 2 * pad4(OUT); }
  // blob layout: W[OUT][IN] b[OUT] gamma[OUT] beta[OUT] (gamma/beta absent for the output layer)
  __device__ static void load(float* s, const float* __restrict__ blob, bool hasBn) {
    for (int i = threadIdx.x; i < OUT * IN4; i += blockDim.x) {
      int o = i / IN4, k = i - o * IN4;
      s[i] = k < IN ? blob[o * IN + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < OUT; i += blockDim.x) {
      s[OUT * IN4 + i] = blob[OUT * IN + i];
      if (hasBn) {
        s[OUT * IN4 + pad4(OUT) + i] = blob[OUT * IN + OUT + i];
        s[OUT * IN4 + 2 * pad4(OUT) + i] = blob[OUT * IN + 2 * OUT + i];
      }
    }
  }
  __host__ __device__ static constexpr int blob_words_bn() { return OUT * IN + 3 * OUT; }
};

// The shipped 2-hidden-layer shape as a unit that ONE WARP can run for 64 PUs (two per lane): used by k2_refine.cu to fold
// K3's work into the persistent K2 kernel as extra work items (fme_config.k3Fuse, experimental).  Same arithmetic, same
// order as k3_nn_fixed.
template <int NEMB, int H1, int H2, int NOUT, bool FMA>
struct NnWarpNet {
  static constexpr int IN0 = 9 + 4 * NEMB;
  using L1 = LayerSmem<IN0, H1>;
  using L2 = LayerSmem<H1, H2>;
  using LO = LayerSmem<H2, NOUT>;
  static constexpr int IN_WORDS = pad4(27 + NEMB * 32);
  static constexpr int WORDS = IN_WORDS + L1::WORDS + L2::WORDS + LO::WORDS;   // CTA-shared image of the net
  static constexpr int ACT_WORDS = (H1 > H2 ? H1 : H2) * 64;                   // per-warp activation transpose
  static constexpr int PUS = 64;                                               // PUs per work item

  // all threads of the CTA; the caller synchronises afterwards
  __device__ static void load(float* s, const float* __restrict__ blob) {
    const float* p = blob + 16;
    for (int i = threadIdx.x; i < 27 + NEMB * 32; i += blockDim.x) s[i] = p[i];
    p += 27 + NEMB * 32;
    L1::load(s + IN_WORDS, p, true); p += L1::blob_words_bn();
    L2::load(s + IN_WORDS + L1::WORDS, p, true); p += L2::blob_words_bn();
    LO::load(s + IN_WORDS + L1::WORDS + L2::WORDS, p, false);
  }

  __device__ static void warp_item(const float* __restrict__ s, float* __restrict__ sAct, const fme_pu* __restrict__ pus,
                                   fme_result* __restrict__ res, int base, int n, int lane, float outClamp) {
    const float* s_in = s;
    const float* s_l1 = s + IN_WORDS;
    const float* s_l2 = s_l1 + L1::WORDS;
    const float* s_lo = s_l2 + L2::WORDS;
    // Code size matters here (the item's code shares the instruction cache with the SATD pack that the other warps of
    // the SM run): the inputs are produced by rolled loops into the activation buffer and read back with static indices,
    // and the layers keep ONE output unit in flight.
#pragma unroll 1
    for (int u = 0; u < K3F_NPU; ++u) {
      const int ii = min(base + u * 32 + lane, n - 1);  // out-of-range lanes compute a valid PU and drop the result
      const unsigned wh = __ldg(reinterpret_cast<const unsigned*>(&pus[ii]) + 1);   // w, h, refSlot, flags
      float* col = sAct + u * 32 + lane;
      if (NEMB == 2) {
        const float* e0 = s_in + 27 + emb_index((int)((wh >> 8) & 0xff), true) * 4;
        const float* e1 = s_in + 27 + 32 + emb_index((int)(wh & 0xff), false) * 4;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) { col[k * 64] = e0[k]; col[(4 + k) * 64] = e1[k]; }
      }
#pragma unroll 1
      for (int k = 0; k < 9; ++k) {
        float e = __uint2float_rn(__ldg(&pus[ii].err[k]));
        e = __fdiv_rn(__fsub_rn(e, s_in[k]), s_in[9 + k]);
        col[(4 * NEMB + k) * 64] = __fmul_rn(e, s_in[18 + k]);
      }
    }
    float x0[K3F_NPU][IN0];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
#pragma unroll
      for (int k = 0; k < IN0; ++k) x0[u][k] = sAct[k * 64 + u * 32 + lane];
    dense_rows<FMA, IN0, H1, 1>(L1::W(s_l1), L1::b(s_l1), x0, [&](int o, const float (&acc)[K3F_NPU]) {
      const float g = L1::g(s_l1)[o], be = L1::be(s_l1)[o];
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u) {
        float a = acc[u] < 0.0f ? 0.0f : acc[u];
        sAct[o * 64 + u * 32 + lane] = __fadd_rn(__fmul_rn(a, g), be);
      }
    });
    float x1[K3F_NPU][H1];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
#pragma unroll
      for (int k = 0; k < H1; ++k) x1[u][k] = sAct[k * 64 + u * 32 + lane];
    dense_rows<FMA, H1, H2, 1>(L2::W(s_l2), L2::b(s_l2), x1, [&](int o, const float (&acc)[K3F_NPU]) {
      const float g = L2::g(s_l2)[o], be = L2::be(s_l2)[o];
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u) {
        float a = acc[u] < 0.0f ? 0.0f : acc[u];
        sAct[o * 64 + u * 32 + lane] = __fadd_rn(__fmul_rn(a, g), be);
      }
    });
    float x2[K3F_NPU][H2];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
#pragma unroll
      for (int k = 0; k < H2; ++k) x2[u][k] = sAct[k * 64 + u * 32 + lane];
    int best[K3F_NPU];
    float bestV[K3F_NPU];
    dense_rows<FMA, H2, NOUT, 1>(LO::W(s_lo), LO::b(s_lo), x2, [&](int o, const float (&acc)[K3F_NPU]) {
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u) {   // first maximum (TEncSearch.cpp:134); outClamp as in k3_nn_fixed
        const float v = fminf(acc[u], outClamp);
        if (o == 0 || v > bestV[u]) { bestV[u] = v; best[u] = o; }
      }
    });
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
      if (base + u * 32 + lane < n) store_class(&res[base + u * 32 + lane], best[u]);
  }
};

}  // namespace
