// Micro-benchmark: small-box TMA (cp.async.bulk.tensor.3d) issue rate per SM on sm_100a.
// Models K2's staging: every lane of a warp fetches its own (BW x BH)-byte box at a pseudo-random position of a
// plane stack (u8, pitch 2176, 1240 rows, 64 planes), the warp waits on one mbarrier, repeats.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tma tools/ubench_tma.cu && tools/ubench_tma
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int PITCH = 2176, ROWS = 1240, PLANES = 64;

__device__ __forceinline__ void mbar_init(unsigned bar, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3d(unsigned dst, const CUtensorMap* m, int x, int y, int p, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(m), "r"(x), "r"(y), "r"(p), "r"(bar) : "memory");
}
// each warp: ITER rounds; in each round `lanesActive` lanes fetch one box each into the warp's smem slice
template <int BOXB>
__global__ void k_tma(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int iters, int lanesActive, int boxBytes, unsigned* out) {
  const CUtensorMap* mp = gmap ? gmap : &pmap;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* mine = smem + (size_t)warp * 32 * BOXB;
  unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[warp]);
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  unsigned rng = (blockIdx.x * 977u + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (lane == 0) mbar_expect(bar, (unsigned)(boxBytes * lanesActive));
    __syncwarp();
    if (lane < lanesActive) {
      rng = rng * 1664525u + 1013904223u;
      int x = (80 + ((rng >> 8) % 1900)) & ~15, y = 80 + ((rng >> 20) % 1000), p = (rng >> 3) & 63;
      tma3d((unsigned)__cvta_generic_to_shared(mine + lane * BOXB), mp, x, y, p, bar);
    }
    mbar_wait(bar, it & 1);
    acc += mine[lane * BOXB + (it & 15)];
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_ref(const __grid_constant__ CUtensorMap map, int x, int y, int p, unsigned* out) {
  __shared__ alignas(128) unsigned char buf[256];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cuda::device::experimental::fence_proxy_async_shared_cta(); }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
  if (threadIdx.x == 0) {
    cuda::device::experimental::cp_async_bulk_tensor_3d_global_to_shared(buf, &map, x, y, p, bar);
    tok = cuda::device::barrier_arrive_tx(bar, 1, 144);
  } else tok = bar.arrive();
  bar.wait(std::move(tok));
  out[threadIdx.x] = buf[threadIdx.x];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool g_useGlobal = false; static CUtensorMap* g_dmap = nullptr;
template <int BW, int BH>
int run(EncodeFn enc, uint8_t* planes, unsigned* out, int warps, int lanesActive) {
  constexpr int BOXB = ((BW * BH + 127) / 128) * 128;
  CUtensorMap map;
  cuuint64_t dims[3] = {PITCH, ROWS, PLANES}, strides[2] = {PITCH, (cuuint64_t)PITCH * ROWS};
  cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  if (g_useGlobal) { if (!g_dmap) cudaMalloc(&g_dmap, sizeof(CUtensorMap)); cudaMemcpy(g_dmap, &map, sizeof(map), cudaMemcpyHostToDevice); }
  size_t smem = (size_t)warps * 32 * BOXB;
  CK(cudaFuncSetAttribute(k_tma<BOXB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_tma<BOXB><<<148, warps * 32, smem>>>(map, g_useGlobal ? g_dmap : nullptr, iters, lanesActive, BW * BH, out);
  cudaEventRecord(e0);
  k_tma<BOXB><<<148, warps * 32, smem>>>(map, g_useGlobal ? g_dmap : nullptr, iters, lanesActive, BW * BH, out);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double boxes = 148.0 * warps * iters * lanesActive, cyc = ms * 1e-3 * clk * 1e3;
  printf("box %2dx%2d (%5d B slot) warps=%2d lanes=%2d: %7.3f ms  %6.2f clk/box/SM  %7.1f useful B/clk/SM  %7.1f GB/s chip\n", BW, BH, BOXB, warps, lanesActive,
         ms, cyc / (boxes / 148), boxes * BW * BH / 148 / cyc, boxes * BW * BH / ms / 1e6);
  return 0;
}
int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  uint8_t* planes; unsigned* out;
  CK(cudaMalloc(&planes, (size_t)PITCH * ROWS * PLANES));
  CK(cudaMemset(planes, 7, (size_t)PITCH * ROWS * PLANES));
  CK(cudaMalloc(&out, 148 * 1024 * 4));
  if (only == 5 || only == 6) {
    CUtensorMap map;
    cuuint64_t dims[3] = {PITCH, ROWS, PLANES}, strides[2] = {PITCH, (cuuint64_t)PITCH * ROWS};
    cuuint32_t box[3] = {16, 9, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, only == 5 ? 3 : 2, planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d\n", (int)r);
    const unsigned* w = reinterpret_cast<const unsigned*>(&map);
    for (int i = 0; i < 32; ++i) printf("%08x%c", w[i], i % 8 == 7 ? '\n' : ' ');
    if (only == 5) { k_ref<<<1, 32>>>(map, 100, 100, 3, out); CK(cudaDeviceSynchronize()); printf("k_ref ok\n"); }
    return 0;
  }
  if (only == 0) return run<16, 9>(enc, planes, out, 1, 1);
  if (only == 2) { g_useGlobal = true; return run<16, 9>(enc, planes, out, 1, 1); }
  if (only == 3) return run<16, 9>(enc, planes, out, 1, 0);
  if (only == 4) g_useGlobal = true;
  if (only == 1) return run<16, 9>(enc, planes, out, 4, 32);
  run<16, 9>(enc, planes, out, 4, 32);  run<16, 9>(enc, planes, out, 12, 32); run<16, 9>(enc, planes, out, 12, 8);
  run<32, 9>(enc, planes, out, 4, 32);  run<32, 9>(enc, planes, out, 8, 32);  run<32, 9>(enc, planes, out, 12, 16);
  run<32, 5>(enc, planes, out, 8, 32); run<32, 17>(enc, planes, out, 8, 16); run<48, 33>(enc, planes, out, 8, 8);
  run<48, 33>(enc, planes, out, 12, 4); run<64, 65>(enc, planes, out, 12, 2);  run<80, 65>(enc, planes, out, 12, 2);
  run<16, 9>(enc, planes, out, 24, 32);
  return 0;
}
