// Micro-benchmark: issue cost of the legacy mma.sync shapes K1's tensor path can use, in SMSP cycles per instruction.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma2 tools/ubench_mma2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__device__ __forceinline__ void one(float (&c)[4], int (&ci)[4], unsigned (&h)[2], const unsigned (&a)[4], const unsigned (&b)[2]) {
  if (KIND == 0) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 1) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
  if (KIND == 2) asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
               : "+r"(h[0]), "+r"(h[1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 3) asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(ci[0]), "+r"(ci[1]), "+r"(ci[2]), "+r"(ci[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 4) asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+r"(ci[0]), "+r"(ci[1]), "+r"(ci[2]), "+r"(ci[3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
  if (KIND == 5) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  if (KIND == 6) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int KIND, int NACC>
__global__ void k(float* out, unsigned seed, int iters) {
  float c[NACC][4]; int ci[NACC][4]; unsigned h[NACC][2];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = seed * (i + 3) + threadIdx.x * 0x10001u;
  b[0] = seed + 7 * threadIdx.x; b[1] = seed ^ (threadIdx.x << 3);
  if (KIND != 3 && KIND != 4) { for (int i = 0; i < 4; ++i) a[i] = (a[i] & 0x03ff03ffu) | 0x3c003c00u; b[0] = (b[0] & 0x03ff03ffu) | 0x3c003c00u; b[1] = (b[1] & 0x03ff03ffu) | 0x3c003c00u; }
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) { c[n][i] = 0; ci[n][i] = 0; h[n][i & 1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < NACC; ++n) one<KIND>(c[n], ci[n], h[n], a, b);
  }
  float s = 0;
  for (int n = 0; n < NACC; ++n) for (int i = 0; i < 4; ++i) s += c[n][i] + ci[n][i] + h[n][i & 1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND, int NACC>
void run(const char* name, int warpsPerSM, int macs) {
  float* d; cudaMalloc(&d, 148 * 2048 * 4);
  int threads = warpsPerSM * 32 > 1024 ? 1024 : warpsPerSM * 32, blocksPerSM = warpsPerSM * 32 / threads;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float msv[2];
  for (int rep = 0; rep < 2; ++rep) {
    int iters = rep ? 16384 : 8192;
    k<KIND, NACC><<<148 * blocksPerSM, threads>>>(d, 3u, iters);
    cudaEventRecord(e0);
    k<KIND, NACC><<<148 * blocksPerSM, threads>>>(d, 3u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&msv[rep], e0, e1);
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ms = msv[1] - msv[0];  // 8192 extra iterations: launch overhead cancels
  double mmasPerSmsp = (double)warpsPerSM / 4 * 8192 * NACC;
  double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-28s acc=%d warps/SM=%2d: %8.3f ms per 8192 iters  %6.2f SMSP-cycles per mma  %7.1f MAC/clk/SM (err %s)\n", name, NACC, warpsPerSM, ms,
         cyc / mmasPerSmsp, 4.0 * macs / (cyc / mmasPerSmsp), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<0, 8>("hmma.16816 f16->f32", 16, 2048); run<0, 4>("hmma.16816 f16->f32", 8, 2048); run<0, 8>("hmma.16816 f16->f32", 4, 2048);
  run<1, 8>("hmma.1688 f16->f32", 16, 1024);
  run<2, 8>("hmma.16816 f16->f16", 16, 2048);
  run<5, 8>("hmma.16816 bf16->f32", 16, 2048);
  run<6, 8>("hmma.1688 tf32->f32", 16, 1024);
  run<3, 8>("imma.16832 s8*u8->s32", 16, 4096); run<3, 8>("imma.16832 s8*u8->s32", 4, 4096);
  run<4, 8>("imma.16816 s8*u8->s32", 16, 2048);
  return 0;
}
