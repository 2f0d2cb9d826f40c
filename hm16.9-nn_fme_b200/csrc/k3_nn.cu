// K3: NN_pred batched over PUs (replaces NN_pred(), TEncSearch.cpp:85-204, and the weight set selected
// in TEncSearch::init, TEncSearch.cpp:470-1073).
//
//   e   = float(err[0..8])                        (uint -> float, round to nearest; TEncSearch.cpp:88)
//   x9  = ((e - mean) / stdev) * gammaIn          (IEEE division; TEncSearch.cpp:89, 116)
//   IN  = [emb0[hIdx] | emb1[wIdx] | x9]          (height map 16->3, 12->4; width map 12->3, 16->4 (sic); :93-117)
//   X_l = relu(W_l X_{l-1} + b_l) * gamma_l + beta_l                                   (:120-127)
//   OUT = W_o X + b_o ; class = first argmax ; class -> (half, quarter) per axis        (:130-193)
//
// The arithmetic is plain fp32 with one rounding per operation and ascending-k dot products
// (__fmul_rn/__fadd_rn forbid FMA contraction), i.e. the operation order of the CPU restatement, so the
// logits are bit-identical to the oracle; against real Eigen the contract is 1e-5 relative (BASELINE.json).
//
// Fast path (architectures known at compile time: the shipped 17-22-20-49 nets and the 9-40-40-40-49
// "3-layer" shape): two PUs per thread, the layer input vector lives in registers (fully unrolled over k),
// weights are re-laid out once per CTA into shared memory with rows padded to 4 floats and fetched as
// broadcast LDS.128 (one wavefront for the whole warp), hidden activations go through a [unit][thread]
// shared-memory transpose so that the output-unit loop can stay a runtime loop (small code, no local memory).
// Generic path: any FMNN blob (up to 4 hidden layers of <= 64 units), one PU per thread, local-memory vectors.
#include "k3_common.cuh"

namespace {

constexpr int K3_THREADS = 128;
constexpr int K3_MAX_WIDTH = 64;  // widest layer supported (inputs or hidden units)

// ------------------------------------------------------------------------------------------------
// generic path
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K3_THREADS) k3_nn_pred(const fme_pu* __restrict__ pus, int n,
                                                        fme_result* __restrict__ res,
                                                        const float* __restrict__ blob, int blobWords) {
  extern __shared__ float s_w[];  // header (16 words) + payload
  for (int i = threadIdx.x; i < blobWords; i += blockDim.x) s_w[i] = blob[i];
  __syncthreads();
  const FmeNnHeader* H = reinterpret_cast<const FmeNnHeader*>(s_w);
  const int nErr = H->nErr, nEmb = H->nEmb, embRows = H->embRows, embDim = H->embDim;
  const int nHidden = H->nHidden, nOut = H->nOut;

  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const fme_pu p = pus[i];
    const float* w = s_w + 16;
    const float* mean = w; w += nErr;
    const float* stdev = w; w += nErr;
    const float* gin = w; w += nErr;
    float x[K3_MAX_WIDTH], y[K3_MAX_WIDTH];
    int nIn = 0;
    if (nEmb == 2) {
      const float* e0 = w + emb_index(p.h, true) * embDim;
      const float* e1 = w + embRows * embDim + emb_index(p.w, false) * embDim;
      for (int k = 0; k < embDim; ++k) x[nIn++] = e0[k];
      for (int k = 0; k < embDim; ++k) x[nIn++] = e1[k];
    }
    w += nEmb * embRows * embDim;
    for (int k = 0; k < nErr; ++k) {
      float e = __uint2float_rn(p.err[k]);
      e = __fdiv_rn(__fsub_rn(e, mean[k]), stdev[k]);
      x[nIn++] = __fmul_rn(e, gin[k]);
    }
    for (int l = 0; l < nHidden; ++l) {
      const int nOutL = H->hidden[l];
      const float* W = w; w += nOutL * nIn;
      const float* b = w; w += nOutL;
      const float* g = w; w += nOutL;
      const float* be = w; w += nOutL;
      for (int o = 0; o < nOutL; ++o) {
        float acc = __fmul_rn(W[o * nIn], x[0]);
        for (int k = 1; k < nIn; ++k) acc = __fadd_rn(acc, __fmul_rn(W[o * nIn + k], x[k]));
        acc = __fadd_rn(acc, b[o]);
        acc = acc < 0.0f ? 0.0f : acc;
        y[o] = __fadd_rn(__fmul_rn(acc, g[o]), be[o]);
      }
      for (int o = 0; o < nOutL; ++o) x[o] = y[o];
      nIn = nOutL;
    }
    const float* W = w; w += nOut * nIn;
    const float* b = w;
    int best = 0;
    float bestV = 0.0f;
    for (int o = 0; o < nOut; ++o) {
      float acc = __fmul_rn(W[o * nIn], x[0]);
      for (int k = 1; k < nIn; ++k) acc = __fadd_rn(acc, __fmul_rn(W[o * nIn + k], x[k]));
      acc = __fadd_rn(acc, b[o]);
      // outSigmoid (the reference's 3-layer backup network): the sigmoid is monotonic, the class is the first maximum
      // of the pre-activation -- no transcendental on the decision path
      // (evaluated in double, as the reference does, the sigmoid is exactly 1.0 from 53 ln 2 on: saturated outputs tie
      //  and the first wins -- the clamp reproduces that)
      if (H->outSigmoid) acc = fminf(acc, 36.7368f);
      if (o == 0 || acc > bestV) { bestV = acc; best = o; }
    }
    store_class(&res[i], best);
  }
}

// ------------------------------------------------------------------------------------------------
// fast path: compile-time layer sizes
// ------------------------------------------------------------------------------------------------
// NH hidden layers of sizes H1,H2,(H3); NEMB = 0 or 2 embedding tables of 8 x 4.
template <int NEMB, int H1, int H2, int H3, int NOUT, bool FMA>
__global__ void __launch_bounds__(K3F_THREADS) k3_nn_fixed(const fme_pu* __restrict__ pus, int n,
                                                          fme_result* __restrict__ res,
                                                          const float* __restrict__ blob, float outClamp) {
  constexpr int IN0 = 9 + 4 * NEMB;
  constexpr int HLAST = H3 > 0 ? H3 : H2;
  using L1 = LayerSmem<IN0, H1>;
  using L2 = LayerSmem<H1, H2>;
  using L3 = LayerSmem<H2, (H3 > 0 ? H3 : 1)>;
  using LO = LayerSmem<HLAST, NOUT>;
  extern __shared__ __align__(16) float s_mem[];
  float* s_in = s_mem;                                  // mean[9] stdev[9] gammaIn[9] emb[NEMB][8][4], padded to 4
  constexpr int IN_WORDS = pad4(27 + NEMB * 32);
  float* s_l1 = s_in + IN_WORDS;
  float* s_l2 = s_l1 + L1::WORDS;
  float* s_l3 = s_l2 + L2::WORDS;
  float* s_lo = s_l3 + (H3 > 0 ? L3::WORDS : 0);
  float* s_act = s_lo + LO::WORDS;                      // [HMAX][K3F_NPU * K3F_THREADS] activation transpose

  {  // stage + re-layout the weights once per CTA
    const float* p = blob + 16;
    for (int i = threadIdx.x; i < 27 + NEMB * 32; i += blockDim.x) s_in[i] = p[i];
    p += 27 + NEMB * 32;
    L1::load(s_l1, p, true); p += L1::blob_words_bn();
    L2::load(s_l2, p, true); p += L2::blob_words_bn();
    if (H3 > 0) { L3::load(s_l3, p, true); p += L3::blob_words_bn(); }
    LO::load(s_lo, p, false);
  }
  __syncthreads();

  const int tid = threadIdx.x;
  constexpr int ACT_STRIDE = K3F_NPU * K3F_THREADS;
  for (int base = blockIdx.x * ACT_STRIDE; base < n; base += gridDim.x * ACT_STRIDE) {
    int idx[K3F_NPU];
    float x0[K3F_NPU][IN0];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u) {
      idx[u] = base + u * K3F_THREADS + tid;
      const int ii = min(idx[u], n - 1);  // out-of-range lanes compute a valid PU and drop the result
      const fme_pu p = pus[ii];
      int c = 0;
      if (NEMB == 2) {
        const float* e0 = s_in + 27 + emb_index(p.h, true) * 4;
        const float* e1 = s_in + 27 + 32 + emb_index(p.w, false) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) x0[u][c++] = e0[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) x0[u][c++] = e1[k];
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float e = __uint2float_rn(p.err[k]);
        e = __fdiv_rn(__fsub_rn(e, s_in[k]), s_in[9 + k]);
        x0[u][c++] = __fmul_rn(e, s_in[18 + k]);
      }
    }
    // hidden layer 1
    dense_rows<FMA, IN0, H1>(L1::W(s_l1), L1::b(s_l1), x0, [&](int o, const float (&acc)[K3F_NPU]) {
      const float g = L1::g(s_l1)[o], be = L1::be(s_l1)[o];
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u) {
        float a = acc[u] < 0.0f ? 0.0f : acc[u];
        s_act[o * ACT_STRIDE + u * K3F_THREADS + tid] = __fadd_rn(__fmul_rn(a, g), be);
      }
    });
    float x1[K3F_NPU][H1];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
#pragma unroll
      for (int k = 0; k < H1; ++k) x1[u][k] = s_act[k * ACT_STRIDE + u * K3F_THREADS + tid];
    // hidden layer 2 (each thread only touches its own column of s_act: no barrier needed)
    dense_rows<FMA, H1, H2>(L2::W(s_l2), L2::b(s_l2), x1, [&](int o, const float (&acc)[K3F_NPU]) {
      const float g = L2::g(s_l2)[o], be = L2::be(s_l2)[o];
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u) {
        float a = acc[u] < 0.0f ? 0.0f : acc[u];
        s_act[o * ACT_STRIDE + u * K3F_THREADS + tid] = __fadd_rn(__fmul_rn(a, g), be);
      }
    });
    float x2[K3F_NPU][H2];
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
#pragma unroll
      for (int k = 0; k < H2; ++k) x2[u][k] = s_act[k * ACT_STRIDE + u * K3F_THREADS + tid];

    int best[K3F_NPU];
    float bestV[K3F_NPU];
    auto argmax = [&](int o, const float (&acc)[K3F_NPU]) {
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u)
      {  // first maximum (TEncSearch.cpp:134); outClamp = 53 ln 2 for sigmoid-output nets (double saturation ties), else +inf
        const float v = fminf(acc[u], outClamp);
        if (o == 0 || v > bestV[u]) { bestV[u] = v; best[u] = o; }
      }
    };
    if constexpr (H3 > 0) {
      dense_rows<FMA, H2, H3>(L3::W(s_l3), L3::b(s_l3), x2, [&](int o, const float (&acc)[K3F_NPU]) {
        const float g = L3::g(s_l3)[o], be = L3::be(s_l3)[o];
#pragma unroll
        for (int u = 0; u < K3F_NPU; ++u) {
          float a = acc[u] < 0.0f ? 0.0f : acc[u];
          s_act[o * ACT_STRIDE + u * K3F_THREADS + tid] = __fadd_rn(__fmul_rn(a, g), be);
        }
      });
      float x3[K3F_NPU][H3 > 0 ? H3 : 1];
#pragma unroll
      for (int u = 0; u < K3F_NPU; ++u)
#pragma unroll
        for (int k = 0; k < H3; ++k) x3[u][k] = s_act[k * ACT_STRIDE + u * K3F_THREADS + tid];
      dense_rows<FMA, (H3 > 0 ? H3 : 1), NOUT>(LO::W(s_lo), LO::b(s_lo), x3, argmax);
    } else {
      dense_rows<FMA, H2, NOUT>(LO::W(s_lo), LO::b(s_lo), x2, argmax);
    }
#pragma unroll
    for (int u = 0; u < K3F_NPU; ++u)
      if (idx[u] < n) store_class(&res[idx[u]], best[u]);
  }
}

template <int NEMB, int H1, int H2, int H3, int NOUT, bool FMA>
cudaError_t launch_fixed(const fme_pu* d_pus, int n, fme_result* d_res, const float* d_nn, bool outSigmoid, cudaStream_t s) {
  constexpr int IN0 = 9 + 4 * NEMB;
  constexpr int HLAST = H3 > 0 ? H3 : H2;
  constexpr int HMAX = (H1 > H2 ? (H1 > H3 ? H1 : H3) : (H2 > H3 ? H2 : H3));
  constexpr int words = pad4(27 + NEMB * 32) + LayerSmem<IN0, H1>::WORDS + LayerSmem<H1, H2>::WORDS +
                        (H3 > 0 ? LayerSmem<H2, (H3 > 0 ? H3 : 1)>::WORDS : 0) + LayerSmem<HLAST, NOUT>::WORDS +
                        HMAX * K3F_NPU * K3F_THREADS;
  const int smem = words * 4;
  if (smem > 48 * 1024) {  // per-device, per-function opt-in; idempotent and cheap, so no cached state shared between host threads
    cudaError_t e = cudaFuncSetAttribute(k3_nn_fixed<NEMB, H1, H2, H3, NOUT, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int blocks = (n + K3F_NPU * K3F_THREADS - 1) / (K3F_NPU * K3F_THREADS);
  // one resident wave of grid-striding CTAs (it was 6 per SM against 5 resident at 90 registers: 0.144 -> 0.138 ms)
  blocks = fme_one_wave(k3_nn_fixed<NEMB, H1, H2, H3, NOUT, FMA>, K3F_THREADS, smem, blocks);
  k3_nn_fixed<NEMB, H1, H2, H3, NOUT, FMA><<<blocks, K3F_THREADS, smem, s>>>(d_pus, n, d_res, d_nn,
                                                                             outSigmoid ? 36.7368f : 3.0e38f);
  return cudaGetLastError();
}

}  // namespace

cudaError_t fme_launch_k3(const fme_pu* d_pus, int n, fme_result* d_res, const float* d_nn, size_t nnBytes,
                          const FmeNnHeader& h, int fma, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return cudaSuccess;
  ++*launches;
  if (h.nOut == 49) {  // (a sigmoid on the outputs does not change the first maximum: same kernels)
    if (h.nEmb == 2 && h.embDim == 4 && h.nHidden == 2 && h.hidden[0] == 22 && h.hidden[1] == 20)
      return fma ? launch_fixed<2, 22, 20, 0, 49, true>(d_pus, n, d_res, d_nn, h.outSigmoid != 0, s)
                 : launch_fixed<2, 22, 20, 0, 49, false>(d_pus, n, d_res, d_nn, h.outSigmoid != 0, s);  // shipped per-QP nets (master)
    if (h.nEmb == 0 && h.nHidden == 3 && h.hidden[0] == 40 && h.hidden[1] == 40 && h.hidden[2] == 40)
      return fma ? launch_fixed<0, 40, 40, 40, 49, true>(d_pus, n, d_res, d_nn, h.outSigmoid != 0, s)
                 : launch_fixed<0, 40, 40, 40, 49, false>(d_pus, n, d_res, d_nn, h.outSigmoid != 0, s);  // "3-layer" 9-40-40-40-49 shape
  }
  int blocks = (n + K3_THREADS - 1) / K3_THREADS;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (nnBytes > 48 * 1024) {  // rare (generic path with a large blob): set on every launch, it is cheap
    cudaError_t e = cudaFuncSetAttribute(k3_nn_pred, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nnBytes);
    if (e != cudaSuccess) return e;
  }
  k3_nn_pred<<<blocks, K3_THREADS, nnBytes, s>>>(d_pus, n, d_res, d_nn, (int)(nnBytes / 4));
  return cudaGetLastError();
}
