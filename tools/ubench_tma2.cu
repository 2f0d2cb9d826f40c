// TMA bring-up probe: one box through the libcu++ wrappers, geometry from argv:  rank bw bh dtype(0=u8,1=f32) x y
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
namespace cde = cuda::device::experimental;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
template <int RANK>
__global__ void k_ref(const __grid_constant__ CUtensorMap map, int x, int y, int p, int bytes, unsigned* out) {
  __shared__ alignas(1024) unsigned char buf[32768];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
  if (threadIdx.x == 0) {
    if (RANK == 3) cde::cp_async_bulk_tensor_3d_global_to_shared(buf, &map, x, y, p, bar);
    else cde::cp_async_bulk_tensor_2d_global_to_shared(buf, &map, x, y, bar);
    tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else tok = bar.arrive();
  bar.wait(std::move(tok));
  unsigned s = 0;
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += buf[i];
  out[threadIdx.x] = s;
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), dt = atoi(argv[4]), x = atoi(argv[5]), y = atoi(argv[6]);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int PITCH = 2176, ROWS = 1240, PLANES = 8;
  uint8_t* planes; unsigned* out;
  CK(cudaMalloc(&planes, (size_t)PITCH * ROWS * PLANES)); CK(cudaMemset(planes, 1, (size_t)PITCH * ROWS * PLANES)); CK(cudaMalloc(&out, 4096));
  int es_ = dt ? 4 : 1;
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)PITCH / es_, ROWS, PLANES}, strides[2] = {PITCH, (cuuint64_t)PITCH * ROWS};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&map, dt ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, planes, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d box %dx%d dt %d at (%d,%d): encode rc %d ... ", rank, bw, bh, dt, x, y, (int)r);
  int bytes = bw * bh * es_;
  if (rank == 3) k_ref<3><<<1, 32>>>(map, x, y, 1, bytes, out); else k_ref<2><<<1, 32>>>(map, x, y, 0, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h[32]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  unsigned s = 0; for (int i = 0; i < 32; ++i) s += h[i];
  printf("%s  sum=%u (expect %d)\n", cudaGetErrorString(e), s, bytes);
  return 0;
}
