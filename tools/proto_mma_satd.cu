// Prototype: 8x8 SATD of 8 tile pairs at once on the tensor pipe (mma.sync m16n8k32 s8 x u8 -> s32).
//   D[m][n] = sum_k A[m][k] * B[k][n],  A = H8 (x) H8 (64x64, entries +-1), B[:, n] = the 64 pixels of tile n.
// K index -> pixel:  k-step ks in {0,1}, k in [0,32): row = 2*((k & 15) >> 2) + ks, col = 4*(k >> 4) + (k & 3),
// so thread (g, t) feeds rows 2t and 2t+1 of tile g.  SATD(n) = (sum_m |D_cand - D_org| + 2) >> 2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/proto_mma_satd tools/proto_mma_satd.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void imma(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte +1 / -1 of (H8 (x) H8)[m][pixel(y, x)], m = 8u + v
__device__ __forceinline__ unsigned had_byte(int m, int y, int x) {
  int u = m >> 3, v = m & 7;
  return ((__popc(u & y) + __popc(v & x)) & 1) ? 0xffu : 0x01u;
}
__device__ __forceinline__ void build_a(unsigned (&A)[4][2][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
  for (int mt = 0; mt < 4; ++mt)
    for (int ks = 0; ks < 2; ++ks)
      for (int r = 0; r < 4; ++r) {
        const int m = 16 * mt + g + 8 * (r & 1);       // a0,a2: row g; a1,a3: row g+8
        const int half = r >> 1;                        // a0,a1: k = 4t..4t+3; a2,a3: k = 16+4t..
        unsigned wv = 0;
        for (int j = 0; j < 4; ++j) wv |= had_byte(m, 2 * t + ks, 4 * half + j) << (8 * j);
        A[mt][ks][r] = wv;
      }
}
__global__ void k(const uint8_t* org, const uint8_t* cand, unsigned* out) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  unsigned A[4][2][4];
  build_a(A, lane);
  auto row = [&](const uint8_t* tile, int r, unsigned& lo, unsigned& hi) {
    const unsigned* p = reinterpret_cast<const unsigned*>(tile + 8 * r);
    lo = p[0]; hi = p[1];
  };
  int dO[4][4], dC[4][4];
  for (int mt = 0; mt < 4; ++mt) for (int i = 0; i < 4; ++i) { dO[mt][i] = 0; dC[mt][i] = 0; }
  unsigned b[2][2];
  row(org + 64 * g, 2 * t, b[0][0], b[0][1]); row(org + 64 * g, 2 * t + 1, b[1][0], b[1][1]);
  for (int mt = 0; mt < 4; ++mt) for (int ks = 0; ks < 2; ++ks) imma(dO[mt], A[mt][ks], b[ks][0], b[ks][1]);
  row(cand + 64 * g, 2 * t, b[0][0], b[0][1]); row(cand + 64 * g, 2 * t + 1, b[1][0], b[1][1]);
  for (int mt = 0; mt < 4; ++mt) for (int ks = 0; ks < 2; ++ks) imma(dC[mt], A[mt][ks], b[ks][0], b[ks][1]);
  unsigned s0 = 0, s1 = 0;  // columns 2t, 2t+1
  for (int mt = 0; mt < 4; ++mt) {
    s0 += abs(dC[mt][0] - dO[mt][0]) + abs(dC[mt][2] - dO[mt][2]);
    s1 += abs(dC[mt][1] - dO[mt][1]) + abs(dC[mt][3] - dO[mt][3]);
  }
  for (int d = 4; d < 32; d <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, d); s1 += __shfl_xor_sync(0xffffffffu, s1, d); }
  if (g == 0) { out[2 * t] = (s0 + 2) >> 2; out[2 * t + 1] = (s1 + 2) >> 2; }
}
static int ref_satd(const uint8_t* o, const uint8_t* c) {
  int d[64], tmp[64];
  for (int i = 0; i < 64; ++i) d[i] = (int)o[i] - (int)c[i];
  for (int u = 0; u < 8; ++u) for (int x = 0; x < 8; ++x) { int s = 0; for (int y = 0; y < 8; ++y) s += (__builtin_popcount(u & y) & 1 ? -1 : 1) * d[8 * y + x]; tmp[8 * u + x] = s; }
  int sum = 0;
  for (int u = 0; u < 8; ++u) for (int v = 0; v < 8; ++v) { int s = 0; for (int x = 0; x < 8; ++x) s += (__builtin_popcount(v & x) & 1 ? -1 : 1) * tmp[8 * u + x]; sum += abs(s); }
  return (sum + 2) >> 2;
}
int main() {
  uint8_t ho[512], hc[512];
  srand(5);
  for (int i = 0; i < 512; ++i) { ho[i] = rand() & 255; hc[i] = (i < 64) ? 255 - ((i * 37) & 1) * 255 : rand() & 255; }
  for (int i = 0; i < 64; ++i) ho[i] = ((i * 37) & 1) * 255;
  uint8_t *dorg, *dc; unsigned* dout;
  cudaMalloc(&dorg, 512); cudaMalloc(&dc, 512); cudaMalloc(&dout, 32);
  cudaMemcpy(dorg, ho, 512, cudaMemcpyHostToDevice); cudaMemcpy(dc, hc, 512, cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dorg, dc, dout);
  unsigned h[8]; cudaError_t e = cudaMemcpy(h, dout, 32, cudaMemcpyDeviceToHost);
  printf("%s\n", cudaGetErrorString(e));
  int bad = 0;
  for (int n = 0; n < 8; ++n) { int r = ref_satd(ho + 64 * n, hc + 64 * n); printf("tile %d: mma %u ref %d\n", n, h[n], r); bad += (int)h[n] != r; }
  printf(bad ? "MISMATCH\n" : "OK\n");
  return bad;
}
