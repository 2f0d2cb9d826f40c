// Store-only floor of K1: write 16 planes of Wp x Hp bytes (pitch rounded to 128) with 16-byte stores and nothing else,
// 32 launches cycling 4 slots like tools/k1_probe.py.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_store tools/ubench_store.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_fill(uint8_t* planes, size_t planeBytes, int pitch, int Wp, int Hp, unsigned v) {
  const int chunksPerRow = (Wp + 15) / 16;
  const long long total = (long long)chunksPerRow * Hp * 16;
  const uint4 val = make_uint4(v, v + 1, v + 2, v + 3);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % chunksPerRow);
    const long long r = i / chunksPerRow;
    const int y = (int)(r % Hp), p = (int)(r / Hp);
    *reinterpret_cast<uint4*>(planes + p * planeBytes + (size_t)y * pitch + c * 16) = val;
  }
}
// leaner: one (plane, 8-row band) job per CTA iteration, threads over the 16-byte chunks of a row -- no divisions per store
__global__ void k_fill2(uint8_t* planes, size_t planeBytes, int pitch, int Wp, int Hp, unsigned v) {
  const int chunksPerRow = (Wp + 15) / 16, bands = (Hp + 7) / 8;
  const uint4 val = make_uint4(v, v + 1, v + 2, v + 3);
  for (int job = blockIdx.x; job < bands * 16; job += gridDim.x) {
    const int p = job / bands, y0 = (job - p * bands) * 8;
    uint8_t* base = planes + p * planeBytes + (size_t)y0 * pitch;
    for (int i = threadIdx.x; i < chunksPerRow * 8; i += blockDim.x) {
      const int r = i / chunksPerRow, c = i - r * chunksPerRow;
      if (y0 + r < Hp) *reinterpret_cast<uint4*>(base + (size_t)r * pitch + c * 16) = val;
    }
  }
}
int main() {
  const int sizes[4][2] = {{1280, 720}, {1920, 1080}, {2560, 1440}, {3840, 2160}};
  for (auto& s : sizes) {
    const int Wp = s[0] + 160, Hp = s[1] + 160, pitch = (Wp + 127) / 128 * 128;
    const size_t planeBytes = (size_t)pitch * Hp, slot = planeBytes * 16;
    uint8_t* d; cudaMalloc(&d, slot * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148 * 4, 148 * 8, 148 * 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 32; ++i) k_fill<<<grid, 256>>>(d + (i % 4) * slot, planeBytes, pitch, Wp, Hp, i);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%dx%d grid %5d: %6.2f us per launch, %6.0f GB/s of plane bytes (%s)\n", s[0], s[1], grid, ms / 32 * 1e3,
             (double)Wp * Hp * 16 / (ms / 32 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 32; ++i) k_fill2<<<grid, 256>>>(d + (i % 4) * slot, planeBytes, pitch, Wp, Hp, i);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%dx%d grid %5d: %6.2f us per launch, %6.0f GB/s of plane bytes (band jobs)\n", s[0], s[1], grid, ms / 32 * 1e3,
             (double)Wp * Hp * 16 / (ms / 32 * 1e-3) / 1e9);
    }
    cudaFree(d);
  }
  // launch overhead: empty grid
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  uint8_t* d; cudaMalloc(&d, 1 << 20);
  cudaEventRecord(e0);
  for (int i = 0; i < 32; ++i) k_fill<<<592, 256>>>(d, 0, 128, 16, 1, i);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("empty launches back to back: %.2f us each\n", ms / 32 * 1e3);
  return 0;
}
