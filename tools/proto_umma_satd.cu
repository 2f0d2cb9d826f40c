// Prototype / micro-benchmark for the tcgen05 SATD of K2 (FME_K2_PATH_UMMA): one CTA of 128 worker threads + an issuer warp,
//   D[M = 128 tile-candidates][N = 64 coefficients] (s32, TMEM) = A[M][K = 64 u8 pixels] x (H8 (x) H8)[N][K] (s8 +-1)
// accumulated over the candidate rows and the COMPLEMENTED source rows (H (c + 255 - o) = H (c - o) + 16320 e0), then
// tcgen05.ld + sum |.| per thread.  Part 1 checks every D element and the SATD against a CPU Hadamard for the canonical
// SWIZZLE_NONE K-major layout (8 rows x 16 bytes core matrices); part 2 times MMA groups and the epilogue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/proto_umma_satd tools/proto_umma_satd.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t smem_desc(unsigned addr, unsigned lboBytes, unsigned sboBytes) {
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
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    if (spin > (1 << 22)) __trap();
  }
}

constexpr int A_BYTES = 128 * 64, B8_BYTES = 64 * 64, B4_BYTES = 32 * 32;

// mode 0: correctness (8x8, N = 64, K = 64): writes D (128 x 64) and the per-row sums.
// mode 1: correctness (two 4x4 tiles per row, N = 32, K = 32).
// mode 2/3: timing of `iters` MMA groups with / without the epilogue (8x8); mode 4: 4x4 groups with epilogue.
__global__ void __launch_bounds__(160, 1)
proto(const uint8_t* __restrict__ cand, const uint8_t* __restrict__ org, int* __restrict__ dOut, unsigned* __restrict__ sums,
      int mode, int iters, long long* cycles, unsigned lboA, unsigned sboA) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* const sB8 = smem;
  uint8_t* const sB4 = smem + B8_BYTES;
  uint8_t* const sA = sB4 + B4_BYTES;          // two candidate buffers
  uint8_t* const sO = sA + 2 * A_BYTES;        // complemented source rows
  __shared__ __align__(8) unsigned long long s_full[2], s_done[2];
  __shared__ unsigned s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool four = mode == 1 || mode == 4;

  for (int i = tid; i < B8_BYTES; i += blockDim.x) {   // i -> (n, k) of the canonical layout
    const int kc = i / 1024, n = ((i % 1024) / 128) * 8 + ((i % 128) / 16), k = kc * 16 + (i % 16);
    const int u = n >> 3, v = n & 7, y = k >> 3, x = k & 7;
    sB8[i] = (uint8_t)(int8_t)(((__popc(u & y) + __popc(v & x)) & 1) ? -1 : 1);
  }
  for (int i = tid; i < B4_BYTES; i += blockDim.x) {
    const int kc = i / 512, n = ((i % 512) / 128) * 8 + ((i % 128) / 16), k = kc * 16 + (i % 16);
    int val = 0;
    if ((n >> 4) == (k >> 4)) {
      const int u = (n >> 2) & 3, v = n & 3, y = (k >> 2) & 3, x = k & 3;
      val = ((__popc(u & y) + __popc(v & x)) & 1) ? -1 : 1;
    }
    sB4[i] = (uint8_t)(int8_t)val;
  }
  const unsigned full0 = (unsigned)__cvta_generic_to_shared(&s_full[0]), done0 = (unsigned)__cvta_generic_to_shared(&s_done[0]);
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(full0 + 8 * b));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done0 + 8 * b));
    }
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows: thread m owns row m of A / O: 16-byte chunk kc at 16 m + kc * 2048
  if (tid < 128) {
    const int nch = four ? 2 : 4;
    for (int kc = 0; kc < nch; ++kc) {
      uint4 c = *reinterpret_cast<const uint4*>(cand + tid * 64 + kc * 16);
      uint4 o = *reinterpret_cast<const uint4*>(org + tid * 64 + kc * 16);
      o.x = ~o.x; o.y = ~o.y; o.z = ~o.z; o.w = ~o.w;
      *reinterpret_cast<uint4*>(sA + 16 * tid + kc * 2048) = c;
      *reinterpret_cast<uint4*>(sA + A_BYTES + 16 * tid + kc * 2048) = c;
      *reinterpret_cast<uint4*>(sO + 16 * tid + kc * 2048) = o;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = s_tmem;
  const unsigned aAddr = (unsigned)__cvta_generic_to_shared(sA), oAddr = (unsigned)__cvta_generic_to_shared(sO);
  const unsigned b8 = (unsigned)__cvta_generic_to_shared(sB8), b4 = (unsigned)__cvta_generic_to_shared(sB4);
  const unsigned N = four ? 32 : 64;
  const unsigned IDESC = (2u << 4) | (1u << 10) | ((N >> 3) << 17) | ((128u >> 4) << 24);   // u8 x s8 -> s32, K-major, M 128
  const bool sw = lboA < sboA;
  const int rounds = mode >= 2 ? iters : 1;
  long long t0 = clock64();
  if (warp == 4) {
    if (lane == 0) {
      for (int r = 0; r < rounds; ++r) {
        const int b = r & 1;
        mbar_wait(full0 + 8 * b, (r >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned d = tmem + b * 64;
        if (!four) {
          umma_i8(d, smem_desc(aAddr + b * A_BYTES, lboA, sboA), smem_desc(b8, sw ? 128 : 1024, sw ? 1024 : 128), IDESC, 0);
          umma_i8(d, smem_desc(aAddr + b * A_BYTES + 4096, lboA, sboA), smem_desc(b8 + 2048, sw ? 128 : 1024, sw ? 1024 : 128), IDESC, 1);
          umma_i8(d, smem_desc(oAddr, lboA, sboA), smem_desc(b8, sw ? 128 : 1024, sw ? 1024 : 128), IDESC, 1);
          umma_i8(d, smem_desc(oAddr + 4096, lboA, sboA), smem_desc(b8 + 2048, sw ? 128 : 1024, sw ? 1024 : 128), IDESC, 1);
        } else {
          umma_i8(d, smem_desc(aAddr + b * A_BYTES, lboA, sboA), smem_desc(b4, sw ? 128 : 512, sw ? 512 : 128), IDESC, 0);
          umma_i8(d, smem_desc(oAddr, lboA, sboA), smem_desc(b4, sw ? 128 : 512, sw ? 512 : 128), IDESC, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done0 + 8 * b) : "memory");
      }
    }
  } else {
    unsigned total = 0;
    for (int r = 0; r <= rounds; ++r) {
      if (r < rounds) {   // "repack" of round r is free here: the rows are already in both buffers
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8 * (r & 1)) : "memory");
      }
      if (r >= 1) {       // deferred epilogue of round r - 1
        const int q = r - 1, b = q & 1;
        mbar_wait(done0 + 8 * b, (q >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mode != 3) {
          const unsigned ta = tmem + ((unsigned)(32 * warp) << 16) + b * 64;
          unsigned s = 0, s2 = 0;
          int v[32];
          tmem_ld32(ta, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (mode < 2) for (int j = 0; j < 32; ++j) dOut[tid * 64 + j] = v[j];
          if (!four) {
            v[0] -= 16320;
#pragma unroll
            for (int j = 0; j < 32; ++j) s += (unsigned)abs(v[j]);
            tmem_ld32(ta + 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (mode < 2) for (int j = 0; j < 32; ++j) dOut[tid * 64 + 32 + j] = v[j];
#pragma unroll
            for (int j = 0; j < 32; ++j) s += (unsigned)abs(v[j]);
            s = (s + 2) >> 2;
          } else {
            v[0] -= 4080; v[16] -= 4080;
#pragma unroll
            for (int j = 0; j < 16; ++j) { s += (unsigned)abs(v[j]); s2 += (unsigned)abs(v[16 + j]); }
            s = ((s + 1) >> 1) + ((s2 + 1) >> 1);
          }
          total += s;
          if (mode < 2) sums[tid] = s;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    if (mode >= 2 && total == 0x12345678u) sums[tid] = total;
  }
  long long t1 = clock64();
  if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
  }
}

static int had8(const int* d, int* out) {   // out[u*8+v] = sum_{y,x} (-1)^(popc(u&y)+popc(v&x)) d[y*8+x]
  int s = 0;
  for (int u = 0; u < 8; ++u) for (int v = 0; v < 8; ++v) {
    int a = 0;
    for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) a += ((__builtin_popcount(u & y) + __builtin_popcount(v & x)) & 1) ? -d[y * 8 + x] : d[y * 8 + x];
    out[u * 8 + v] = a; s += abs(a);
  }
  return (s + 2) >> 2;
}
static int had4(const int* d) {
  int s = 0;
  for (int u = 0; u < 4; ++u) for (int v = 0; v < 4; ++v) {
    int a = 0;
    for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) a += ((__builtin_popcount(u & y) + __builtin_popcount(v & x)) & 1) ? -d[y * 4 + x] : d[y * 4 + x];
    s += abs(a);
  }
  return (s + 1) >> 1;
}

int main() {
  const int smemBytes = B8_BYTES + B4_BYTES + 3 * A_BYTES;
  cudaFuncSetAttribute(proto, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
  std::vector<uint8_t> hc(128 * 64), ho(128 * 64);
  srand(5);
  for (auto& v : hc) v = rand() & 255;
  for (auto& v : ho) v = rand() & 255;
  for (int i = 0; i < 64; ++i) { hc[i] = 255; ho[i] = 0; hc[64 + i] = 0; ho[64 + i] = 255; }   // extremes in rows 0, 1
  uint8_t *dc, *dob; int* dD; unsigned* dS; long long* dCyc;
  cudaMalloc(&dc, hc.size()); cudaMalloc(&dob, ho.size()); cudaMalloc(&dD, 128 * 64 * 4); cudaMalloc(&dS, 128 * 4); cudaMalloc(&dCyc, 1024 * 8);
  cudaMemcpy(dc, hc.data(), hc.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dob, ho.data(), ho.size(), cudaMemcpyHostToDevice);
  const unsigned layouts[2][2] = {{2048, 128}, {128, 2048}};
  for (int L = 0; L < 1; ++L) {
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(dD, 0, 128 * 64 * 4); cudaMemset(dS, 0, 128 * 4);
      proto<<<1, 160, smemBytes>>>(dc, dob, dD, dS, mode, 1, nullptr, layouts[L][0], layouts[L][1]);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("layout %d mode %d: CUDA error %s\n", L, mode, cudaGetErrorString(e)); return 1; }
      std::vector<int> D(128 * 64); std::vector<unsigned> S(128);
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(S.data(), dS, S.size() * 4, cudaMemcpyDeviceToHost);
      int badD = 0, badS = 0;
      for (int m = 0; m < 128; ++m) {
        int d[64], coef[64];
        for (int k = 0; k < 64; ++k) d[k] = (int)hc[m * 64 + k] - (int)ho[m * 64 + k];
        if (mode == 0) {
          const int want = had8(d, coef);
          coef[0] += 16320;
          for (int n = 0; n < 64; ++n) badD += D[m * 64 + n] != coef[n];
          badS += (int)S[m] != want;
          if (m < 2 || (badS && m < 6)) printf("   row %d: satd gpu %u cpu %d  D[0..3] gpu %d %d %d %d cpu %d %d %d %d\n", m, S[m], want, D[m * 64], D[m * 64 + 1], D[m * 64 + 2], D[m * 64 + 3], coef[0], coef[1], coef[2], coef[3]);
        } else {
          const int want = had4(d) + had4(d + 16);
          badS += (int)S[m] != want;
          if (m < 2) printf("   row %d: 2x satd4 gpu %u cpu %d\n", m, S[m], want);
        }
      }
      printf("layout LBO=%u SBO=%u mode %d (%s): %d wrong D elements, %d wrong SATDs of 128\n", layouts[L][0], layouts[L][1], mode, mode ? "2 x 4x4" : "8x8", badD, badS);
    }
  }
  // timing: one CTA per SM and two CTAs per SM
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int perSM = 1; perSM <= 2; ++perSM)
    for (int mode = 2; mode <= 4; ++mode) {
      const int iters = 4000, grid = 148 * perSM;
      proto<<<grid, 160, smemBytes>>>(dc, dob, dD, dS, mode, iters, dCyc, 2048, 128);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      proto<<<grid, 160, smemBytes>>>(dc, dob, dD, dS, mode, iters, dCyc, 2048, 128);
      cudaEventRecord(e1);
      cudaError_t e = cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cyc(grid);
      cudaMemcpy(cyc.data(), dCyc, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (auto c : cyc) avg += c; avg /= grid;
      printf("timing mode %d (%s) %d CTA/SM: %.3f ms for %d rounds of 128 tile-candidates per CTA -> %.0f clk per round per CTA, %.2f G tile-cand/s on the chip (%s)\n",
             mode, mode == 2 ? "8x8 mma + epilogue" : mode == 3 ? "8x8 mma only" : "2x4x4 mma + epilogue", perSM, ms, iters, avg / iters,
             (double)grid * iters * 128 / (ms * 1e-3) * 1e-9, cudaGetErrorString(e));
    }
  return 0;
}
