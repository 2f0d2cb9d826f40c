// Micro-benchmark: legacy mma.sync (IMMA u8 x s8 -> s32, HMMA f16 -> f32) rate per SM on sm_100a, alone and with
// concurrent integer ALU work.  Used to decide whether the SATD transform is worth moving to the tensor pipe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_mma tools/ubench_mma.cu && /tmp/ubench_mma
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void imma(int (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void hmma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// NACC independent accumulator chains per warp, ALU extra integer adds per mma
template <int NACC, int ALU>
__global__ void k_imma(int* out, unsigned seed) {
  int c[NACC][4];
  unsigned a[4], b[2];
  int x[8];
  for (int i = 0; i < 4; ++i) a[i] = seed * (i + 1) + threadIdx.x;
  b[0] = seed + 7 * threadIdx.x; b[1] = seed ^ threadIdx.x;
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) c[n][i] = 0;
  for (int i = 0; i < 8; ++i) x[i] = seed + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n) {
      imma(c[n], a, b);
#pragma unroll
      for (int j = 0; j < ALU; ++j) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[(n * ALU + j) & 7]) : "r"(b[0]));
    }
  }
  int s = 0;
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += c[n][i];
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_hmma(float* out, unsigned seed) {
  float c[NACC][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = 0x3c003c00u;
  b[0] = 0x3c003c00u; b[1] = 0x3c003c00u + (seed & 0);
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) c[n][i] = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n) hmma(c[n], a, b);
  }
  float s = 0;
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += c[n][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
static float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
template <int NACC, int ALU>
void run_imma(int warpsPerSM) {
  int* d; cudaMalloc(&d, 148 * 2048 * 4);
  int threads = warpsPerSM * 32 > 1024 ? 1024 : warpsPerSM * 32, blocksPerSM = warpsPerSM * 32 / threads;
  float ms = timeit([&] { k_imma<NACC, ALU><<<148 * blocksPerSM, threads>>>(d, 3u); });
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double mmas = 148.0 * warpsPerSM * ITERS * NACC;
  double macs = mmas * 16 * 8 * 32;
  printf("imma m16n8k32 acc=%d alu=%d warps/SM=%2d: %7.3f ms  %7.1f MAC/clk/SM  %6.2f mma/clk/SM  %6.1f Tops/s  (alu %5.1f lane-ops/clk/SM)\n", NACC, ALU, warpsPerSM, ms,
         macs / (ms * 1e-3) / 148 / (clk * 1e3), mmas / (ms * 1e-3) / 148 / (clk * 1e3), 2 * macs / ms / 1e9,
         mmas * ALU * 32 / (ms * 1e-3) / 148 / (clk * 1e3));
  cudaFree(d);
}
template <int NACC>
void run_hmma(int warpsPerSM) {
  float* d; cudaMalloc(&d, 148 * 2048 * 4);
  int threads = warpsPerSM * 32 > 1024 ? 1024 : warpsPerSM * 32, blocksPerSM = warpsPerSM * 32 / threads;
  float ms = timeit([&] { k_hmma<NACC><<<148 * blocksPerSM, threads>>>(d, 3u); });
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double mmas = 148.0 * warpsPerSM * ITERS * NACC;
  double macs = mmas * 16 * 8 * 16;
  printf("hmma m16n8k16 acc=%d warps/SM=%2d: %7.3f ms  %7.1f MAC/clk/SM  %6.2f mma/clk/SM  %6.1f TFLOP/s\n", NACC, warpsPerSM, ms,
         macs / (ms * 1e-3) / 148 / (clk * 1e3), mmas / (ms * 1e-3) / 148 / (clk * 1e3), 2 * macs / ms / 1e9);
  cudaFree(d);
}
int main() {
  run_imma<1, 0>(4); run_imma<4, 0>(4); run_imma<8, 0>(4); run_imma<4, 0>(8); run_imma<4, 0>(16); run_imma<8, 0>(16); run_imma<4, 0>(32);
  run_imma<4, 2>(16); run_imma<4, 4>(16); run_imma<4, 8>(16); run_imma<4, 16>(16);
  run_hmma<4>(4); run_hmma<4>(16); run_hmma<8>(16);
  return 0;
}
