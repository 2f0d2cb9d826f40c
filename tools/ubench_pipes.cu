// Micro-benchmark: per-SM throughput (lane-ops per clock) of the integer instructions K1/K2 lean on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int OP>
__device__ __forceinline__ void one(int& x, int b, int b0) {
  if (OP == 0) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(b), "r"(b0));
  if (OP == 1) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(b), "r"(b0));
  if (OP == 2) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(x) : "r"(b), "r"(b0));
  if (OP == 3) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(b));
  if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x) : "r"(b));
  if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(b0));
  if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(b0));
  if (OP == 7) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
  if (OP == 8) asm volatile("{.reg .f32 t; mov.b32 t, %0; fma.rn.f32 t, t, 1.0001, 0.5; mov.b32 %0, t;}" : "+r"(x));
  if (OP == 9) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(b0));
  if (OP == 10) asm volatile("{.reg .b32 t; mov.b32 t, %0; add.f16x2 t, t, t; mov.b32 %0, t;}" : "+r"(x));
  if (OP == 11) asm volatile("{.reg .f32 t; mov.b32 t, %0; mul.rn.f32 t, t, 1.0001; mov.b32 %0, t;}" : "+r"(x));
  if (OP == 12) asm volatile("{.reg .f32 t; mov.b32 t, %0; add.rn.f32 t, t, 0.5; mov.b32 %0, t;}" : "+r"(x));
}
// two instruction kinds interleaved 1:1 on independent registers: 2x the single-kind rate means different pipes
template <int A, int B>
__global__ void kp(int* out, int a0, int b0) {
  int a[8], b = b0 + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) { one<A>(a[i], b, b0); one<B>(a[i + 1], b, b0); }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int A, int B>
void runp(const char* name) {
  int* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kp<A, B><<<148 * 4, 512>>>(d, 1, 3);
  cudaEventRecord(e0);
  kp<A, B><<<148 * 4, 512>>>(d, 1, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 4 * 512 * ITERS * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-14s %8.3f ms  %6.1f lane-ops/clk/SM\n", name, ms, ops / (ms * 1e-3) / 148 / (clk * 1e3));
  cudaFree(d);
}
template <int OP>
__global__ void k(int* out, int a0, int b0) {
  int a[8], b = b0 + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = a0 + i + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(b0));
      if (OP == 1) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(b0));
      if (OP == 2) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(b0));
      if (OP == 3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a[i]) : "r"(b));
      if (OP == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(b0));
      if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(b0));
      if (OP == 7) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 8) asm volatile("{.reg .f32 t; mov.b32 t, %0; fma.rn.f32 t, t, 1.0001, 0.5; mov.b32 %0, t;}" : "+r"(a[i]));
      if (OP == 9) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(b0));
      if (OP == 10) asm volatile("{.reg .b32 t; mov.b32 t, %0; add.f16x2 t, t, t; mov.b32 %0, t;}" : "+r"(a[i]));
      if (OP == 11) asm volatile("{.reg .f32 t; mov.b32 t, %0; mul.rn.f32 t, t, 1.0001; mov.b32 %0, t;}" : "+r"(a[i]));
      if (OP == 12) asm volatile("{.reg .f32 t; mov.b32 t, %0; add.rn.f32 t, t, 0.5; mov.b32 %0, t;}" : "+r"(a[i]));
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name) {
  int* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<148 * 4, 512>>>(d, 1, 3);
  cudaEventRecord(e0);
  k<OP><<<148 * 4, 512>>>(d, 1, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 4 * 512 * ITERS * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-10s %8.3f ms  %7.2f Tops/s  %6.1f lane-ops/clk/SM (at %d MHz)\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(d);
}
int main() {
  run<0>("dp4a"); run<1>("dp2a"); run<2>("imad"); run<3>("iadd"); run<4>("prmt"); run<5>("shf"); run<6>("lop3");
  run<7>("vimnmx16x2"); run<8>("ffma"); run<9>("i2ip"); run<10>("hadd2"); run<11>("fmul"); run<12>("fadd");
  runp<11, 12>("fmul+fadd"); runp<11, 2>("fmul+imad"); runp<12, 3>("fadd+iadd");
  runp<0, 2>("dp4a+imad"); runp<0, 4>("dp4a+prmt"); runp<2, 4>("imad+prmt"); runp<2, 8>("imad+ffma"); runp<4, 8>("prmt+ffma");
  runp<1, 8>("dp2a+ffma"); runp<5, 4>("shf+prmt"); runp<9, 2>("i2ip+imad"); runp<9, 4>("i2ip+prmt"); runp<10, 4>("hadd2+prmt");
  runp<10, 2>("hadd2+imad"); runp<6, 2>("lop3+imad"); runp<6, 4>("lop3+prmt"); runp<0, 6>("dp4a+lop3"); runp<3, 4>("iadd+prmt");
  runp<3, 2>("iadd+imad"); runp<7, 4>("vimnmx+prmt"); runp<7, 2>("vimnmx+imad");
  return 0;
}
