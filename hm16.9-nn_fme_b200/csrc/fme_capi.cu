// C ABI of libfme_b200.so (include/fme_b200.h): context, device memory, streams, launch sequencing.
// There is deliberately no CPU path in this file: every entry point either runs CUDA work on an
// sm_100 device or fails with an error code.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "fme_common.cuh"

namespace {

thread_local std::string t_lastError;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_lastError = buf;
  return code;
}

#define CU_CHECK(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(FME_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                         \
  } while (0)

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

// Depth of the host<->device pipeline of the async entry points (record/result/picture buffer rings).  With two
// buffers the copy-in stream idles while the host waits for the oldest read-back (measured 1.16 ms per 1080p frame
// against a 1.05 ms PCIe floor); three keep it busy.
constexpr int FME_NBUF = 3;

// What FME_K2_PATH_AUTO resolves to: the fastest measured path (profiles/r2_k2_paths.txt).
#ifndef FME_K2_PATH_DEFAULT
#define FME_K2_PATH_DEFAULT FME_K2_PATH_SWAR
#endif

struct fme_ctx {
  fme_config cfg;
  FmeGeom g;
  int numSMs = 0;
  // `stream` carries every kernel.  With the ctx's own stream, host<->device copies of the async entry points run
  // on two extra streams (sIn, sOut) and are ordered against the kernels with events, so that the copies of one
  // frame overlap the kernels of its neighbours; with a caller-provided stream everything is issued on that stream.
  cudaStream_t ownStream = nullptr, stream = nullptr;
  cudaStream_t ownIn = nullptr, ownOut = nullptr, sIn = nullptr, sOut = nullptr;
  cudaEvent_t evIn[FME_NBUF] = {}, evDone[FME_NBUF] = {}, evOut[FME_NBUF] = {};  // per record/result buffer
  cudaEvent_t evPicIn[FME_NBUF] = {}, evPicFree[FME_NBUF] = {};                  // per picture staging buffer
  cudaEvent_t evOrgFree[FME_NBUF] = {};                                          // per source-picture buffer
  uint64_t submitSeq = 0, picSeq = 0, orgSeq = 0;
  int fifo[FME_NBUF] = {}, fifoCount = 0;                       // outstanding async submits (buffer ids), oldest first
  // device memory
  uint8_t* d_planes = nullptr;   // [slots][16][Hp][pitch]
  uint8_t* d_orgBuf[FME_NBUF] = {};  // ring of source pictures, [H + 2][orgPitch] each
  uint8_t* d_org = nullptr;      // the buffer the next submit reads
  uint8_t* d_picBuf[FME_NBUF] = {};  // ring of raw u8 picture staging buffers [H][picPitch]
  uint8_t* d_pic = nullptr;      // staging buffer K1 reads for the most recent upload
  int picPitch = 0;
  int16_t* d_pel = nullptr;      // Pel staging for uploads / block-level calls
  size_t pelCapacity = 0;        // in samples
  int16_t* d_pel2 = nullptr;
  size_t pel2Capacity = 0;
  int16_t* d_pelPic[FME_NBUF] = {};  // Pel picture staging for fme_upload_ref / fme_upload_org
  uint8_t* d_cb = nullptr;       // [slots][Hcp][cPitch]
  uint8_t* d_cr = nullptr;
  fme_pu* d_pusBuf[FME_NBUF] = {};
  fme_pu_head* d_headBuf[FME_NBUF] = {};  // allocated on the first fme_submit_heads*
  fme_err_grid* d_gridBuf[FME_NBUF] = {};  // allocated on the first fme_submit_heads_grids*
  fme_pu_compact* d_compactBuf[FME_NBUF] = {};  // allocated on the first fme_submit_compact*
  fme_result* d_resBuf[FME_NBUF] = {};
  fme_result8* d_res8Buf[FME_NBUF] = {};  // FME_MODE_RESULT8: packed copy that travels to the host
  fme_pu* d_pus = nullptr;       // = d_pusBuf[0], used by the synchronous helpers (fme_mc)
  fme_result* d_res = nullptr;
  float* d_nn = nullptr;
  size_t nnBytes = 0;
  FmeNnHeader nnHeader{};
  int* d_k1Counter = nullptr;    // K1's dynamic tile hand-out (2 ints, re-armed by the kernel)
  uint32_t* d_scratchU32 = nullptr;
  size_t scratchU32Capacity = 0;
  void* d_opBuf = nullptr;       // staging of the synchronous batched operators (fme_cand_cost, fme_mc_luma_compact)
  size_t opBufCap = 0;
  FmeK2Scratch k2{};
  // state
  std::vector<char> refValid;
  bool orgValid = false, sliceValid = false, nnValid = false;
  double lambda = 0.0;
  FmeCostLut costLut;            // the current slice's table; copied into every K2 launch as a kernel argument
  // profiling
  bool profiling = false;
  cudaEvent_t ev[8] = {};
  float lastMs[4] = {0, 0, 0, 0};
  int64_t launches = 0;
};

namespace {

int ensure_pel(int16_t** p, size_t* cap, size_t samples) {
  if (*cap >= samples) return FME_OK;
  if (*p) cudaFree(*p);  // cudaFree synchronises the device, so in-flight users of the old buffer have drained
  *p = nullptr;
  *cap = 0;
  CU_CHECK(cudaMalloc(p, samples * sizeof(int16_t)));
  *cap = samples;
  return FME_OK;
}

int ensure_u32(fme_ctx* c, size_t n) {
  if (c->scratchU32Capacity >= n) return FME_OK;
  if (c->d_scratchU32) cudaFree(c->d_scratchU32);
  c->d_scratchU32 = nullptr;
  c->scratchU32Capacity = 0;
  CU_CHECK(cudaMalloc(&c->d_scratchU32, n * sizeof(uint32_t)));
  c->scratchU32Capacity = n;
  return FME_OK;
}

// Copy a host picture (w x h, stride in samples) into a device u8 picture.  No CPU pass over the samples:
// u8 input is DMA'd straight into place; Pel (int16) input is DMA'd into a device staging plane and narrowed by a
// kernel.  The DMA runs on the copy-in stream; `stagingIdx` selects the Pel staging plane and the events that
// keep a staging buffer from being overwritten while an earlier convert/K1 still reads it.
int stage_picture(fme_ctx* c, const uint8_t* src, int stride, int w, int h, uint8_t* d_dst, int dstPitch, int stagingIdx) {
  if (!src || stride < w) return fail(FME_ERR_INVALID, "picture pointer/stride invalid");
  (void)stagingIdx;
  CU_CHECK(cudaMemcpy2DAsync(d_dst, dstPitch, src, stride, w, h, cudaMemcpyHostToDevice, c->sIn));
  return FME_OK;
}
int stage_picture(fme_ctx* c, const int16_t* src, int stride, int w, int h, uint8_t* d_dst, int dstPitch, int stagingIdx) {
  if (!src || stride < w) return fail(FME_ERR_INVALID, "picture pointer/stride invalid");
  CU_CHECK(cudaMemcpy2DAsync(c->d_pelPic[stagingIdx], (size_t)w * 2, src, (size_t)stride * 2, (size_t)w * 2, h,
                             cudaMemcpyHostToDevice, c->sIn));
  // the narrowing kernel runs on the kernel stream once the DMA has landed
  CU_CHECK(cudaEventRecord(c->evPicIn[stagingIdx], c->sIn));
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evPicIn[stagingIdx], 0));
  CU_CHECK(fme_launch_pel_to_u8(c->d_pelPic[stagingIdx], w, d_dst, dstPitch, w, h, c->stream, &c->launches));
  return FME_OK;
}

int check_slot(fme_ctx* c, int slot) {
  if (!c) return fail(FME_ERR_INVALID, "null ctx");
  if (slot < 0 || slot >= c->cfg.numRefSlots) return fail(FME_ERR_INVALID, "slot %d out of range", slot);
  return FME_OK;
}

// Event pair around one pass.  Events belong to this library's CUDA runtime instance; recording them on a
// stream created by another runtime instance (e.g. torch's) is rejected by the driver with "invalid resource
// handle", so a failed record switches profiling off instead of poisoning the next launch check.
struct StageTimer {
  fme_ctx* c;
  int idx;
  void rec(int e) {
    if (!c->profiling) return;
    if (cudaEventRecord(c->ev[e], c->stream) != cudaSuccess) {
      cudaGetLastError();
      c->profiling = false;
      for (float& m : c->lastMs) m = -1.f;
    }
  }
  // every pass is also an NVTX range ("fme:K1 interp", ...) for Nsight Systems / Compute timelines
  StageTimer(fme_ctx* c_, int i) : c(c_), idx(i) {
    static const char* const names[4] = {"fme:K1 interp", "fme:K2 refine", "fme:K3 nn_pred", "fme:K0 int surface"};
    nvtxRangePushA(names[idx & 3]);
    rec(2 * idx);
  }
  ~StageTimer() {
    rec(2 * idx + 1);
    nvtxRangePop();
  }
};
struct NvtxRange {  // host-side span of an ABI call (copies + launches it issues)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

// Debug build only (-DFME_STAMPS, tools/e2e_timeline.py): %globaltimer stamps on the kernel stream around the passes of a
// frame, to see where the kernel stream is busy and where it waits for copies.
#ifdef FME_STAMPS
__global__ void k_stamp(unsigned long long* buf, int idx, int tag) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  buf[idx] = (t << 4) | (unsigned long long)tag;
}
static unsigned long long* g_stamps = nullptr;
static int g_stampN = 0;
static void dbg_stamp(fme_ctx* c, int tag) {
  if (!g_stamps) cudaMalloc(&g_stamps, 8 * 65536);
  if (g_stampN < 65536) k_stamp<<<1, 1, 0, c->stream>>>(g_stamps, g_stampN++, tag);
}
#else
static inline void dbg_stamp(fme_ctx*, int) {}
#endif

int run_k1(fme_ctx* c, int slot, const uint8_t* d_pic = nullptr, int picPitch = 0, int rowBegin = 0, int rowEnd = 1 << 30) {
  dbg_stamp(c, 1);
  StageTimer t(c, 0);
  if (!d_pic) { d_pic = c->d_pic; picPitch = c->picPitch; }
  CU_CHECK(fme_launch_k1(c->g, d_pic, picPitch, c->d_planes + (size_t)slot * c->g.slotBytes, c->d_k1Counter, c->numSMs,
                         rowBegin, rowEnd, c->cfg.k1Path, c->stream,
                         &c->launches));
  dbg_stamp(c, 2);
  c->refValid[slot] = 1;
  return FME_OK;
}

// Pick the next picture staging buffer: its previous reader (convert / K1 on the kernel stream) must be done
// before the copy-in stream overwrites it.
int begin_picture_upload(fme_ctx* c, int* idx) {
  *idx = (int)(c->picSeq++ % FME_NBUF);
  CU_CHECK(cudaStreamWaitEvent(c->sIn, c->evPicFree[*idx], 0));
  c->d_pic = c->d_picBuf[*idx];
  return FME_OK;
}
// After the copy (and narrowing) have been issued: kernels wait for the DMA, and the buffer is released once the
// kernel stream has consumed it.
int end_picture_upload(fme_ctx* c, int idx, bool kernelsFollow) {
  CU_CHECK(cudaEventRecord(c->evPicIn[idx], c->sIn));
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evPicIn[idx], 0));
  (void)kernelsFollow;
  return FME_OK;
}
int release_picture(fme_ctx* c, int idx) {
  CU_CHECK(cudaEventRecord(c->evPicFree[idx], c->stream));
  return FME_OK;
}

// k0BehindK2: the unvalidated path may carry FME_PU_ERR_ON_GPU records (K3 needs their err[]); K2's binning pass counts
// them while it reads the records anyway, and the K0 pass launched behind K2 returns at once when there are none (a K0
// launched ahead of K2 has to scan every record to find that out: 0.05 ms per 858 000 records).
int run_search(fme_ctx* c, fme_pu* d_pus, int n, fme_result* d_out, int mode, bool k0BehindK2 = false) {
  if (mode < 1 || mode > 3) return fail(FME_ERR_INVALID, "mode must be FME_MODE_STD|NN|BOTH");
  if ((mode & FME_MODE_STD) && (!c->orgValid || !c->sliceValid))
    return fail(FME_ERR_STATE, "fme_submit(STD) needs fme_upload_org and fme_set_slice first");
  if ((mode & FME_MODE_NN) && !c->nnValid) return fail(FME_ERR_STATE, "fme_submit(NN) needs fme_set_nn_weights first");
  if (mode != FME_MODE_BOTH) CU_CHECK(fme_launch_clear_results(d_out, n, c->stream, &c->launches));
  // fme_config.k3Fuse (experimental, off by default): both passes over the same records, the shipped net shape, no K0
  // between them -> K3's work items ride inside the K2 kernel (k2_refine.cu)
  bool nnDone = false;
  const FmeNnHeader& nh = c->nnHeader;
  const bool fusable = mode == FME_MODE_BOTH && !k0BehindK2 && c->cfg.k3Fuse == 1 && nh.nOut == 49 && nh.nEmb == 2 &&
                       nh.embDim == 4 && nh.nHidden == 2 && nh.hidden[0] == 22 && nh.hidden[1] == 20;
  const FmeK2NnFuse nnFuse = {c->d_nn, c->cfg.nnFma, nh.outSigmoid ? 36.7368f : 3.0e38f};
  if (mode & FME_MODE_STD) {
    StageTimer t(c, 1);
    CU_CHECK(fme_launch_k2(c->g, c->d_planes, c->d_org, d_pus, n, d_out, c->costLut, c->cfg.useHadME, c->cfg.biPred,
                           c->cfg.k2Path == FME_K2_PATH_AUTO ? FME_K2_PATH_DEFAULT : c->cfg.k2Path, c->k2, c->numSMs, c->stream,
                           &c->launches, fusable ? &nnFuse : nullptr, &nnDone));
  }
  if (k0BehindK2) {
    StageTimer t(c, 3);
    CU_CHECK(fme_launch_k0(c->g, c->d_planes, c->d_org, d_pus, n, c->cfg.fen, c->stream, &c->launches, c->k2.workCounter + 1));
  }
  dbg_stamp(c, 5);
  if ((mode & FME_MODE_NN) && !nnDone) {
    StageTimer t(c, 2);
    CU_CHECK(fme_launch_k3(d_pus, n, d_out, c->d_nn, c->nnBytes, c->nnHeader, c->cfg.nnFma, c->stream, &c->launches));
  }
  dbg_stamp(c, 6);
  return FME_OK;
}

int sync_all(fme_ctx* c) {
  CU_CHECK(cudaStreamSynchronize(c->sIn));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  CU_CHECK(cudaStreamSynchronize(c->sOut));
  c->fifoCount = 0;
  return FME_OK;
}

void collect_ms(fme_ctx* c) {
  if (!c->profiling) return;
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    if (cudaEventQuery(c->ev[2 * i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev[2 * i], c->ev[2 * i + 1]) == cudaSuccess)
      c->lastMs[i] = ms;
  }
}

}  // namespace

#ifdef FME_STAMPS
extern "C" int fme_debug_stamps(unsigned long long* out, int max) {
  cudaDeviceSynchronize();
  const int n = g_stampN < max ? g_stampN : max;
  if (n > 0) cudaMemcpy(out, g_stamps, 8 * (size_t)n, cudaMemcpyDeviceToHost);
  g_stampN = 0;
  return n;
}
#endif

extern "C" {

const char* fme_last_error(void) { return t_lastError.c_str(); }
const char* fme_version(void) { return "fme_b200 0.1 (sm_100a)"; }

int fme_create(const fme_config* cfg, fme_ctx** out) {
  if (!cfg || !out) return fail(FME_ERR_INVALID, "null argument");
  *out = nullptr;
  // at least one CTU: the kernels clamp a PU's source block into [0, W - w] x [0, H - h] and rely on w <= W, h <= H
  if (cfg->width < 64 || cfg->height < 64 || cfg->width > 16384 || cfg->height > 16384)
    return fail(FME_ERR_INVALID, "unsupported picture size %dx%d (64x64 .. 16384x16384)", cfg->width, cfg->height);
  if ((cfg->width & 7) || (cfg->height & 7))
    return fail(FME_ERR_INVALID, "picture size %dx%d is not a multiple of the minimum CU size 8", cfg->width, cfg->height);
  if (cfg->margin < 16 || (cfg->margin & 15)) return fail(FME_ERR_INVALID, "margin must be a multiple of 16, >= 16");
  if (cfg->bitDepth != 8) return fail(FME_ERR_INVALID, "frame-level passes support bitDepth 8 only");
  if (cfg->numRefSlots < 1 || cfg->numRefSlots > 64) return fail(FME_ERR_INVALID, "numRefSlots out of range");
  if (cfg->maxPUs < 1) return fail(FME_ERR_INVALID, "maxPUs must be positive");
  if (cfg->k2Path < FME_K2_PATH_AUTO || cfg->k2Path > FME_K2_PATH_UMMA) return fail(FME_ERR_INVALID, "k2Path out of range");
  if (cfg->k1Path < FME_K1_PATH_AUTO || cfg->k1Path > FME_K1_PATH_UMMA) return fail(FME_ERR_INVALID, "k1Path out of range");
  if (cfg->k3Fuse < 0 || cfg->k3Fuse > 1) return fail(FME_ERR_INVALID, "k3Fuse must be 0 or 1");

  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
    cudaGetLastError();
    return fail(FME_ERR_NO_DEVICE, "no CUDA device visible; this engine has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= nDev) return fail(FME_ERR_INVALID, "device %d out of range", cfg->device);
  cudaDeviceProp prop;
  CU_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(FME_ERR_NO_DEVICE, "device %d is sm_%d%d; libfme_b200 is built for sm_100a only", cfg->device, prop.major,
                prop.minor);
  CU_CHECK(cudaSetDevice(cfg->device));

  fme_ctx* c = new fme_ctx();
  c->cfg = *cfg;
  c->numSMs = prop.multiProcessorCount;
  FmeGeom& g = c->g;
  g.W = cfg->width; g.H = cfg->height; g.M = cfg->margin;
  g.Wp = g.W + 2 * g.M; g.Hp = g.H + 2 * g.M;
  g.pitch = round_up(g.Wp, 128);
  g.planeBytes = (size_t)g.Hp * g.pitch;
  g.slotBytes = g.planeBytes * FME_NUM_PLANES;
  g.orgPitch = round_up(g.W + 16, 128);
  g.Wc = g.W / 2; g.Hc = g.H / 2; g.Mc = g.M / 2;
  g.Wcp = g.Wc + 2 * g.Mc; g.Hcp = g.Hc + 2 * g.Mc;
  g.cPitch = round_up(g.Wcp, 128);
  g.cPlaneBytes = (size_t)g.Hcp * g.cPitch;
  g.numSlots = cfg->numRefSlots;
  if (g.slotBytes > (size_t)INT32_MAX) {  // the kernels form plane offsets (plane * planeBytes + row * pitch) in 32-bit int
    int rc = fail(FME_ERR_INVALID, "picture %dx%d with margin %d: a 16-plane slot of %zu bytes exceeds the 2^31 offset range",
                  g.W, g.H, g.M, g.slotBytes);
    delete c;
    return rc;
  }
  c->picPitch = round_up(g.W, 256);
  c->refValid.assign(cfg->numRefSlots, 0);

#define CREATE_CHECK(expr)                                                                    \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      int rc = fail(FME_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));                   \
      fme_destroy(c);                                                                         \
      return rc;                                                                              \
    }                                                                                         \
  } while (0)
  CREATE_CHECK(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
  CREATE_CHECK(cudaStreamCreateWithFlags(&c->ownIn, cudaStreamNonBlocking));
  CREATE_CHECK(cudaStreamCreateWithFlags(&c->ownOut, cudaStreamNonBlocking));
  c->stream = c->ownStream; c->sIn = c->ownIn; c->sOut = c->ownOut;
  for (auto& e : c->ev) CREATE_CHECK(cudaEventCreate(&e));
  for (int b = 0; b < FME_NBUF; ++b) {
    for (cudaEvent_t* e : {&c->evIn[b], &c->evDone[b], &c->evOut[b], &c->evPicIn[b], &c->evPicFree[b], &c->evOrgFree[b]})
      CREATE_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  // +1 plane of slack so that word-granular staging reads past the last row stay inside the allocation
  CREATE_CHECK(cudaMalloc(&c->d_planes, g.slotBytes * cfg->numRefSlots + g.pitch * 2));
  CREATE_CHECK(cudaMemsetAsync(c->d_planes, 0, g.slotBytes * cfg->numRefSlots + g.pitch * 2, c->stream));
  for (int b = 0; b < FME_NBUF; ++b) {
    CREATE_CHECK(cudaMalloc(&c->d_orgBuf[b], (size_t)(g.H + 2) * g.orgPitch));
    CREATE_CHECK(cudaMemsetAsync(c->d_orgBuf[b], 0, (size_t)(g.H + 2) * g.orgPitch, c->stream));
    CREATE_CHECK(cudaMalloc(&c->d_picBuf[b], (size_t)g.H * c->picPitch));
    CREATE_CHECK(cudaMalloc(&c->d_pelPic[b], (size_t)g.W * g.H * sizeof(int16_t)));
    CREATE_CHECK(cudaMalloc(&c->d_pusBuf[b], sizeof(fme_pu) * (size_t)cfg->maxPUs));
    CREATE_CHECK(cudaMalloc(&c->d_resBuf[b], sizeof(fme_result) * (size_t)cfg->maxPUs));
    CREATE_CHECK(cudaMalloc(&c->d_res8Buf[b], sizeof(fme_result8) * (size_t)cfg->maxPUs));
  }
  c->d_org = c->d_orgBuf[0]; c->d_pic = c->d_picBuf[0]; c->d_pus = c->d_pusBuf[0]; c->d_res = c->d_resBuf[0];
  CREATE_CHECK(cudaMalloc(&c->d_k1Counter, 2 * sizeof(int)));
  CREATE_CHECK(cudaMemsetAsync(c->d_k1Counter, 0, 2 * sizeof(int), c->stream));
  CREATE_CHECK(cudaMalloc(&c->k2.classCount, sizeof(int) * (4 * FME_K2_KEYS + 16)));
  c->k2.classCursor = c->k2.classCount + FME_K2_KEYS;
  c->k2.workCounter = c->k2.classCursor + FME_K2_KEYS;  // counts, cursors and the work counter are cleared together
  c->k2.classOffset = c->k2.workCounter + 4;
  c->k2.packOffset = c->k2.classOffset + FME_K2_KEYS + 1;
  CREATE_CHECK(cudaMalloc(&c->k2.order, sizeof(int) * (size_t)cfg->maxPUs));
  CREATE_CHECK(cudaMalloc(&c->k2.keys, sizeof(short) * (size_t)cfg->maxPUs));
  CREATE_CHECK(cudaStreamSynchronize(c->stream));
#undef CREATE_CHECK
  *out = c;
  return FME_OK;
}

void fme_destroy(fme_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  cudaFree(c->d_planes); cudaFree(c->d_pel); cudaFree(c->d_pel2);
  for (int b = 0; b < FME_NBUF; ++b) {
    cudaFree(c->d_orgBuf[b]); cudaFree(c->d_picBuf[b]); cudaFree(c->d_pelPic[b]); cudaFree(c->d_pusBuf[b]); cudaFree(c->d_resBuf[b]); cudaFree(c->d_res8Buf[b]);
    cudaFree(c->d_headBuf[b]); cudaFree(c->d_gridBuf[b]); cudaFree(c->d_compactBuf[b]);
    for (cudaEvent_t e : {c->evIn[b], c->evDone[b], c->evOut[b], c->evPicIn[b], c->evPicFree[b], c->evOrgFree[b]})
      if (e) cudaEventDestroy(e);
  }
  cudaFree(c->d_cb); cudaFree(c->d_cr); cudaFree(c->d_nn); cudaFree(c->d_opBuf);
  cudaFree(c->d_k1Counter); cudaFree(c->d_scratchU32); cudaFree(c->k2.classCount); cudaFree(c->k2.order); cudaFree(c->k2.keys);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->ownStream) cudaStreamDestroy(c->ownStream);
  if (c->ownIn) cudaStreamDestroy(c->ownIn);
  if (c->ownOut) cudaStreamDestroy(c->ownOut);
  delete c;
}

int fme_set_stream(fme_ctx* c, void* s) {
  if (!c) return fail(FME_ERR_INVALID, "null ctx");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  int rc = sync_all(c);
  if (rc) return rc;
  if (s) {  // caller-provided stream: everything, copies included, is issued on it
    c->stream = c->sIn = c->sOut = static_cast<cudaStream_t>(s);
  } else {
    c->stream = c->ownStream; c->sIn = c->ownIn; c->sOut = c->ownOut;
  }
  return FME_OK;
}

int fme_synchronize(fme_ctx* c) {
  if (!c) return fail(FME_ERR_INVALID, "null ctx");
  int rc = sync_all(c);
  if (rc) return rc;
  collect_ms(c);
  return FME_OK;
}

int fme_wait_oldest(fme_ctx* c) {
  if (!c) return fail(FME_ERR_INVALID, "null ctx");
  if (c->fifoCount == 0) return FME_OK;
  CU_CHECK(cudaEventSynchronize(c->evOut[c->fifo[0]]));
  for (int i = 1; i < c->fifoCount; ++i) c->fifo[i - 1] = c->fifo[i];
  --c->fifoCount;
  return FME_OK;
}

int fme_set_profiling(fme_ctx* c, int on) {
  if (!c) return fail(FME_ERR_INVALID, "null ctx");
  c->profiling = on != 0;
  return FME_OK;
}

int fme_last_kernel_ms(fme_ctx* c, float k[4]) {
  if (!c || !k) return fail(FME_ERR_INVALID, "null argument");
  collect_ms(c);
  for (int i = 0; i < 4; ++i) k[i] = c->lastMs[i];
  return FME_OK;
}

int64_t fme_launch_count(fme_ctx* c) { return c ? c->launches : 0; }

// ---- NN weights ----------------------------------------------------------------------------
static size_t nn_payload_floats(const FmeNnHeader& h) {
  size_t n = 3 * (size_t)h.nErr + (size_t)h.nEmb * h.embRows * h.embDim;
  int in = h.nErr + h.nEmb * h.embDim;
  for (int l = 0; l < h.nHidden; ++l) {
    n += (size_t)h.hidden[l] * in + 3 * (size_t)h.hidden[l];
    in = h.hidden[l];
  }
  return n + (size_t)h.nOut * in + h.nOut;
}

int fme_set_nn_weights(fme_ctx* c, const void* blob, size_t bytes) {
  if (!c || !blob || bytes < sizeof(FmeNnHeader)) return fail(FME_ERR_INVALID, "bad weight blob");
  FmeNnHeader h;
  memcpy(&h, blob, sizeof(h));
  if (h.magic != FME_NN_MAGIC) return fail(FME_ERR_INVALID, "weight blob: bad magic");
  if (h.nErr != 9 || (h.nEmb != 0 && h.nEmb != 2) || h.nHidden < 1 || h.nHidden > 4 || h.nOut < 1 || h.nOut > 64 ||
      (h.nEmb == 2 && (h.embRows != 8 || h.embDim < 1 || h.embDim > 8)))
    return fail(FME_ERR_INVALID, "weight blob: unsupported architecture");
  for (int l = 0; l < h.nHidden; ++l)
    if (h.hidden[l] < 1 || h.hidden[l] > 64) return fail(FME_ERR_INVALID, "weight blob: hidden width out of range");
  size_t need = sizeof(FmeNnHeader) + 4 * nn_payload_floats(h);
  if (bytes != need) return fail(FME_ERR_INVALID, "weight blob: size %zu, expected %zu", bytes, need);
  if (need > 200 * 1024) return fail(FME_ERR_INVALID, "weight blob too large for shared memory");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  {
    int rc = sync_all(c);
    if (rc) return rc;
  }
  if (c->d_nn) cudaFree(c->d_nn);
  c->d_nn = nullptr;
  CU_CHECK(cudaMalloc(&c->d_nn, need));
  CU_CHECK(cudaMemcpy(c->d_nn, blob, need, cudaMemcpyHostToDevice));
  c->nnBytes = need;
  c->nnHeader = h;
  c->nnValid = true;
  return FME_OK;
}

// DL/blowing/<qp> layout: files "<k>.<name>.csv", k = 1..14 (DL/edit.sh:13-17)
static int read_csv(const std::string& path, std::vector<std::vector<double>>& rows) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return fail(FME_ERR_IO, "cannot open %s", path.c_str());
  std::string text;
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  rows.clear();
  std::vector<double> cur;
  const char* p = text.c_str();
  while (*p) {
    if (*p == '\n') {
      if (!cur.empty()) rows.push_back(cur);
      cur.clear();
      ++p;
    } else if (*p == ',' || *p == ';' || *p == ' ' || *p == '\t' || *p == '\r') {
      ++p;
    } else {
      char* end = nullptr;
      double v = strtod(p, &end);
      if (end == p) return fail(FME_ERR_IO, "parse error in %s", path.c_str());
      cur.push_back(v);
      p = end;
    }
  }
  if (!cur.empty()) rows.push_back(cur);
  return FME_OK;
}

int fme_load_nn_csv_dir(fme_ctx* c, const char* dir) {
  if (!c || !dir) return fail(FME_ERR_INVALID, "null argument");
  static const char* names[14] = {"1.emb0-weight", "2.emb1-weight", "3.lins0-weight", "4.lins1-weight", "5.outp-weight",
                                  "6.lins0-bias", "7.lins1-bias", "8.outp-bias", "9.bn-weight", "10.bns0-weight",
                                  "11.bns1-weight", "12.bns0-bias", "13.bns1-bias", nullptr};
  std::vector<std::vector<double>> t[14];
  for (int i = 0; i < 13; ++i) {
    int rc = read_csv(std::string(dir) + "/" + names[i] + ".csv", t[i]);
    if (rc) return rc;
  }
  // 14.mapper_<qp>.csv: the QP is part of the name; try the four shipped ones
  int rc = FME_ERR_IO;
  for (int qp : {22, 27, 32, 37}) {
    rc = read_csv(std::string(dir) + "/14.mapper_" + std::to_string(qp) + ".csv", t[13]);
    if (rc == FME_OK) break;
  }
  if (rc) return rc;
  auto flat = [](const std::vector<std::vector<double>>& r, std::vector<float>& out) {
    for (auto& row : r)
      for (double v : row) out.push_back((float)v);
  };
  if (t[13].size() != 2 || t[13][0].size() != 9) return fail(FME_ERR_IO, "mapper csv malformed");
  FmeNnHeader h{};
  h.magic = FME_NN_MAGIC; h.version = 1; h.nErr = 9; h.nEmb = 2; h.embRows = 8; h.embDim = 4; h.nHidden = 2;
  h.hidden[0] = (int)t[2].size(); h.hidden[1] = (int)t[3].size(); h.nOut = (int)t[4].size();
  std::vector<float> p;
  flat({t[13][0]}, p); flat({t[13][1]}, p); flat(t[8], p);       // mean, stdev, gammaIn
  flat(t[0], p); flat(t[1], p);                                    // embeddings
  flat(t[2], p); flat(t[5], p); flat(t[9], p); flat(t[11], p);     // layer 0: W b gamma beta
  flat(t[3], p); flat(t[6], p); flat(t[10], p); flat(t[12], p);    // layer 1
  flat(t[4], p); flat(t[7], p);                                    // output
  if (p.size() != nn_payload_floats(h)) return fail(FME_ERR_IO, "csv directory has unexpected shapes");
  std::vector<char> blob(sizeof(h) + 4 * p.size());
  memcpy(blob.data(), &h, sizeof(h));
  memcpy(blob.data() + sizeof(h), p.data(), 4 * p.size());
  return fme_set_nn_weights(c, blob.data(), blob.size());
}

// ---- slice lambda ----------------------------------------------------------------------------
int fme_set_slice(fme_ctx* c, double lambda) {
  if (!c || !(lambda > 0.0)) return fail(FME_ERR_INVALID, "lambda must be positive");
  c->lambda = lambda;
  // TComRdCost.cpp:108 m_dLambdaMotionSAD[0]; TComRdCost.h:159 selectMotionLambda(true, 0, false);
  // TComRdCost.h:165-169 Distortion((m_motionLambda * bits) / 65536.0)
  // (FME_PU_LOSSLESS records use the same table: bit-exact for the standard lossy and the all-lossless cost modes;
  //  COST_MIXED_LOSSLESS_LOSSY_CODING would select m_dLambdaMotionSAD[1] for transquant-bypass PUs, TComRdCost.cpp:110-116)
  const double motionLambda = 65536.0 * sqrt(lambda);
  for (unsigned b = 0; b < FME_COST_LUT_SIZE; ++b) c->costLut.v[b] = (uint32_t)((motionLambda * b) / 65536.0);
  // No synchronisation and no device copy: submits already issued carry the previous table in their launch arguments.
  c->sliceValid = true;
  return FME_OK;
}

int fme_mv_cost(fme_ctx* c, int x, int y, int scale, int predX, int predY, uint32_t* out) {
  if (!c || !out) return fail(FME_ERR_INVALID, "null argument");
  if (!c->sliceValid) return fail(FME_ERR_STATE, "fme_set_slice not called");
  auto bits = [](int v) {
    unsigned len = 1, t = (v <= 0) ? (((unsigned)(-v)) << 1) + 1 : ((unsigned)v << 1);
    while (t != 1) { t >>= 1; len += 2; }
    return len;
  };
  unsigned b = bits((x << scale) - predX) + bits((y << scale) - predY);
  if (b >= FME_COST_LUT_SIZE) return fail(FME_ERR_INVALID, "vector out of range");
  *out = c->costLut.v[b];
  return FME_OK;
}

// ---- frame data ----------------------------------------------------------------------------
extern "C++" {
template <typename T>
static int upload_ref_host(fme_ctx* c, int slot, const T* y, int stride) {
  NvtxRange nvtxSpan("fme_upload_ref");
  int rc = check_slot(c, slot);
  if (rc) return rc;
  CU_CHECK(cudaSetDevice(c->cfg.device));
  int idx;
  if ((rc = begin_picture_upload(c, &idx))) return rc;
  if ((rc = stage_picture(c, y, stride, c->g.W, c->g.H, c->d_pic, c->picPitch, idx))) return rc;
  if ((rc = end_picture_upload(c, idx, true))) return rc;
  if ((rc = run_k1(c, slot))) return rc;
  return release_picture(c, idx);
}
}  // extern "C++"
int fme_upload_ref(fme_ctx* c, int slot, const int16_t* y, int stride) { return upload_ref_host(c, slot, y, stride); }
int fme_upload_ref_u8(fme_ctx* c, int slot, const uint8_t* y, int stride) { return upload_ref_host(c, slot, y, stride); }

int fme_upload_ref_device_u8(fme_ctx* c, int slot, const uint8_t* d_y, int pitch) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!d_y || pitch < c->g.W) return fail(FME_ERR_INVALID, "bad device picture");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  // device-resident input, ordered on the kernel stream.  K1 reads 32-bit words: a 4-byte aligned picture is
  // interpolated straight from the caller's buffer (which must stay unchanged until the work issued here has run, as
  // with any stream-ordered call); anything else goes through a staging copy.
  if (((size_t)d_y & 3) == 0 && (pitch & 3) == 0) return run_k1(c, slot, d_y, pitch);
  int idx = (int)(c->picSeq++ % FME_NBUF);
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evPicFree[idx], 0));
  c->d_pic = c->d_picBuf[idx];
  CU_CHECK(cudaMemcpy2DAsync(c->d_pic, c->picPitch, d_y, pitch, c->g.W, c->g.H, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = run_k1(c, slot))) return rc;
  return release_picture(c, idx);
}

// Banded multi-GPU mode: a rank interpolates only the rows its own PUs can reference.  picRowBegin / picRowEnd are PICTURE
// rows (may lie outside [0, H): the margin); the sub-pel planes are produced for these rows, every other row of the slot
// keeps whatever it held.  The caller derives the range from its band's records: a PU at row y with integer MV my and
// height h reads plane rows y + my - 1 .. y + my + h (quarter-pel candidates one row up / down, TEncSearch.cpp:1613-1623).
int fme_upload_ref_device_u8_rows(fme_ctx* c, int slot, const uint8_t* d_y, int pitch, int picRowBegin, int picRowEnd) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!d_y || pitch < c->g.W || ((size_t)d_y & 3) || (pitch & 3)) return fail(FME_ERR_INVALID, "bad device picture (4-byte aligned base and pitch required)");
  if (picRowEnd <= picRowBegin) return fail(FME_ERR_INVALID, "empty row range");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  return run_k1(c, slot, d_y, pitch, picRowBegin + c->g.M, picRowEnd + c->g.M);
}

int fme_interp_slot(fme_ctx* c, int slot) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!c->refValid[slot]) return fail(FME_ERR_STATE, "fme_interp_slot: slot %d holds no picture", slot);
  CU_CHECK(cudaSetDevice(c->cfg.device));
  // the slot's own picture = the picture area of its plane 0 (integer-pel copy, margin M on every side)
  int idx = (int)(c->picSeq++ % FME_NBUF);
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evPicFree[idx], 0));
  c->d_pic = c->d_picBuf[idx];
  const uint8_t* plane0 = c->d_planes + (size_t)slot * c->g.slotBytes + (size_t)c->g.M * c->g.pitch + c->g.M;
  CU_CHECK(cudaMemcpy2DAsync(c->d_pic, c->picPitch, plane0, c->g.pitch, c->g.W, c->g.H, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = run_k1(c, slot))) return rc;
  return release_picture(c, idx);
}

// The source picture lives in a ring: a new upload goes to a buffer no in-flight submit reads.
static int begin_org_upload(fme_ctx* c, cudaStream_t copyStream, int* idx) {
  *idx = (int)(c->orgSeq++ % FME_NBUF);
  CU_CHECK(cudaStreamWaitEvent(copyStream, c->evOrgFree[*idx], 0));  // last submit that read this buffer is done
  c->d_org = c->d_orgBuf[*idx];
  return FME_OK;
}
extern "C++" {
template <typename T>
static int upload_org_host(fme_ctx* c, const T* y, int stride) {
  if (!c) return fail(FME_ERR_INVALID, "null ctx");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  int idx, rc;
  if ((rc = begin_org_upload(c, c->sIn, &idx))) return rc;
  // the narrowing kernel (Pel input) also writes d_org on the kernel stream: make that stream respect the same hazard
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evOrgFree[idx], 0));
  int pidx = (int)(c->picSeq++ % FME_NBUF);
  CU_CHECK(cudaStreamWaitEvent(c->sIn, c->evPicFree[pidx], 0));
  if ((rc = stage_picture(c, y, stride, c->g.W, c->g.H, c->d_org, c->g.orgPitch, pidx))) return rc;
  CU_CHECK(cudaEventRecord(c->evPicIn[pidx], c->sIn));
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evPicIn[pidx], 0));
  CU_CHECK(cudaEventRecord(c->evPicFree[pidx], c->stream));
  c->orgValid = true;
  return FME_OK;
}
}  // extern "C++"
int fme_upload_org(fme_ctx* c, const int16_t* y, int stride) { return upload_org_host(c, y, stride); }
int fme_upload_org_u8(fme_ctx* c, const uint8_t* y, int stride) { return upload_org_host(c, y, stride); }

int fme_upload_org_device_u8(fme_ctx* c, const uint8_t* d_y, int pitch) {
  if (!c || !d_y || pitch < c->g.W) return fail(FME_ERR_INVALID, "bad device picture");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  int idx, rc;
  if ((rc = begin_org_upload(c, c->stream, &idx))) return rc;
  CU_CHECK(cudaMemcpy2DAsync(c->d_org, c->g.orgPitch, d_y, pitch, c->g.W, c->g.H, cudaMemcpyDeviceToDevice, c->stream));
  c->orgValid = true;
  return FME_OK;
}

int fme_upload_org_device_u8_rows(fme_ctx* c, const uint8_t* d_y, int pitch, int rowBegin, int rowEnd) {
  if (!c || !d_y || pitch < c->g.W) return fail(FME_ERR_INVALID, "bad device picture");
  if (rowBegin < 0) rowBegin = 0;
  if (rowEnd > c->g.H) rowEnd = c->g.H;
  if (rowEnd <= rowBegin) return fail(FME_ERR_INVALID, "empty row range");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  int idx, rc;
  if ((rc = begin_org_upload(c, c->stream, &idx))) return rc;
  CU_CHECK(cudaMemcpy2DAsync(c->d_org + (size_t)rowBegin * c->g.orgPitch, c->g.orgPitch, d_y + (size_t)rowBegin * pitch, pitch,
                             c->g.W, rowEnd - rowBegin, cudaMemcpyDeviceToDevice, c->stream));
  c->orgValid = true;
  return FME_OK;
}

extern "C++" {
template <typename T>
static int upload_ref_chroma_host(fme_ctx* c, int slot, const T* cb, const T* cr, int stride) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!cb || !cr) return fail(FME_ERR_INVALID, "null chroma plane");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  const FmeGeom& g = c->g;
  if (!c->d_cb) {
    CU_CHECK(cudaMalloc(&c->d_cb, g.cPlaneBytes * c->cfg.numRefSlots + g.cPitch));
    CU_CHECK(cudaMalloc(&c->d_cr, g.cPlaneBytes * c->cfg.numRefSlots + g.cPitch));
  }
  for (int k = 0; k < 2; ++k) {
    int idx;
    if ((rc = begin_picture_upload(c, &idx))) return rc;
    if ((rc = stage_picture(c, k ? cr : cb, stride, g.Wc, g.Hc, c->d_pic, c->picPitch, idx))) return rc;
    if ((rc = end_picture_upload(c, idx, true))) return rc;
    CU_CHECK(fme_launch_pad_chroma(g, c->d_pic, c->picPitch, (k ? c->d_cr : c->d_cb) + (size_t)slot * g.cPlaneBytes,
                                   c->stream, &c->launches));
    if ((rc = release_picture(c, idx))) return rc;
  }
  return FME_OK;
}
}  // extern "C++"
int fme_upload_ref_chroma(fme_ctx* c, int slot, const int16_t* cb, const int16_t* cr, int stride) {
  return upload_ref_chroma_host(c, slot, cb, cr, stride);
}
int fme_upload_ref_chroma_u8(fme_ctx* c, int slot, const uint8_t* cb, const uint8_t* cr, int stride) {
  return upload_ref_chroma_host(c, slot, cb, cr, stride);
}

// One frame of a raw planar 8-bit 4:2:0 file (TVideoIOYuv layout: Y, then Cb, then Cr, no padding) straight into the
// padded device planes of a reference slot: three DMAs from the caller's frame buffer, no host pass over the samples.
int fme_upload_ref_yuv420_u8(fme_ctx* c, int slot, const uint8_t* frame, int withChroma) {
  if (!c || !frame) return fail(FME_ERR_INVALID, "null argument");
  const int W = c->cfg.width, H = c->cfg.height;
  int rc = fme_upload_ref_u8(c, slot, frame, W);
  if (rc || !withChroma) return rc;
  const uint8_t* cb = frame + (size_t)W * H;
  return fme_upload_ref_chroma_u8(c, slot, cb, cb + (size_t)(W / 2) * (H / 2), W / 2);
}
int fme_upload_org_yuv420_u8(fme_ctx* c, const uint8_t* frame) {
  if (!c || !frame) return fail(FME_ERR_INVALID, "null argument");
  return fme_upload_org_u8(c, frame, c->cfg.width);  // fractional ME searches on luma only (TEncSearch.cpp:4474-4479)
}

// Inter PU shapes of HEVC (fme_hevc_pu_shape): K2's lane units are 8x8 tiles or PAIRS of 4x4 tiles, and every HEVC
// shape that is 4x4-tiled has an even tile count.
static bool valid_pu_size(int w, int h) { return fme_hevc_pu_shape(w, h); }

// ---- the batched search ------------------------------------------------------------------------
static int submit_common(fme_ctx* c, const fme_pu* pus, int n, fme_result* out, int mode, bool sync,
                         const fme_pu_head* heads = nullptr, const fme_err_grid* grids = nullptr, int nGrids = 0,
                         const fme_pu_compact* compact = nullptr) {
  if (!c || (!pus && !heads && !compact) || !out) return fail(FME_ERR_INVALID, "null argument");
  if (nGrids < 0 || (nGrids > 0 && (!grids || (!heads && !compact)))) return fail(FME_ERR_INVALID, "bad error-grid list");
  if (nGrids > c->cfg.maxPUs) return fail(FME_ERR_INVALID, "nGrids=%d exceeds maxPUs=%d", nGrids, c->cfg.maxPUs);
  NvtxRange nvtxSpan(heads ? "fme_submit_heads" : compact ? "fme_submit_compact" : "fme_submit");
  const bool packed = (mode & FME_MODE_RESULT8) != 0;  // `out` is an fme_result8 array
  mode &= ~FME_MODE_RESULT8;
  if (n < 0 || n > c->cfg.maxPUs) return fail(FME_ERR_INVALID, "n=%d exceeds maxPUs=%d", n, c->cfg.maxPUs);
  if (n == 0) return FME_OK;
  CU_CHECK(cudaSetDevice(c->cfg.device));
  bool needK0 = false;
  if (heads) {
    needK0 = true;
    if (sync) {
      for (int i = 0; i < n; ++i) {
        if (heads[i].refSlot >= c->cfg.numRefSlots || !c->refValid[heads[i].refSlot])
          return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, heads[i].refSlot);
        if (!valid_pu_size(heads[i].w, heads[i].h))
          return fail(FME_ERR_INVALID, "PU %d: %dx%d is not an HEVC PU size", i, heads[i].w, heads[i].h);
        if (heads[i].flags & FME_PU_BI)
          return fail(FME_ERR_INVALID, "PU %d: FME_PU_BI needs a full record (the other list's slot and MV live in err[])", i);
      }
      for (int j = 0; j < nGrids; ++j)
        if (grids[j].pu < 0 || grids[j].pu >= n) return fail(FME_ERR_INVALID, "error grid %d names PU %d of %d", j, grids[j].pu, n);
    }
    for (int b = 0; b < FME_NBUF; ++b)
      if (!c->d_headBuf[b]) CU_CHECK(cudaMalloc(&c->d_headBuf[b], sizeof(fme_pu_head) * (size_t)c->cfg.maxPUs));
    for (int b = 0; b < FME_NBUF && nGrids > 0; ++b)
      if (!c->d_gridBuf[b]) CU_CHECK(cudaMalloc(&c->d_gridBuf[b], sizeof(fme_err_grid) * (size_t)c->cfg.maxPUs));
  } else if (compact) {   // grids travel with the records (or in the big list): no K0 pass
    if (sync) {
      for (int i = 0; i < n; ++i) {
        const fme_pu_head& h = compact[i].head;
        if ((mode & FME_MODE_STD) && (h.refSlot >= c->cfg.numRefSlots || !c->refValid[h.refSlot]))
          return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, h.refSlot);
        if ((mode & FME_MODE_STD) && !valid_pu_size(h.w, h.h))
          return fail(FME_ERR_INVALID, "PU %d: %dx%d is not an HEVC PU size", i, h.w, h.h);
        if (h.flags & FME_PU_BI)
          return fail(FME_ERR_INVALID, "PU %d: FME_PU_BI needs a full record (the other list's slot and MV live in err[])", i);
      }
      for (int j = 0; j < nGrids; ++j)
        if (grids[j].pu < 0 || grids[j].pu >= n) return fail(FME_ERR_INVALID, "error grid %d names PU %d of %d", j, grids[j].pu, n);
    }
    for (int b = 0; b < FME_NBUF; ++b)
      if (!c->d_compactBuf[b]) CU_CHECK(cudaMalloc(&c->d_compactBuf[b], sizeof(fme_pu_compact) * (size_t)c->cfg.maxPUs));
    for (int b = 0; b < FME_NBUF && nGrids > 0; ++b)
      if (!c->d_gridBuf[b]) CU_CHECK(cudaMalloc(&c->d_gridBuf[b], sizeof(fme_err_grid) * (size_t)c->cfg.maxPUs));
  } else if (sync) {  // full validation on the synchronous path
    for (int i = 0; i < n; ++i) {
      if (pus[i].flags & FME_PU_ERR_ON_GPU) needK0 = true;
      if ((mode & FME_MODE_STD) || (pus[i].flags & FME_PU_ERR_ON_GPU)) {
        if (pus[i].refSlot >= c->cfg.numRefSlots || !c->refValid[pus[i].refSlot])
          return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, pus[i].refSlot);
        if (!valid_pu_size(pus[i].w, pus[i].h))
          return fail(FME_ERR_INVALID, "PU %d: %dx%d is not an HEVC PU size", i, pus[i].w, pus[i].h);
        if (pus[i].flags & FME_PU_BI) {
          if (!c->cfg.biPred) return fail(FME_ERR_INVALID, "PU %d carries FME_PU_BI but the ctx was created with biPred = 0", i);
          if (pus[i].flags & FME_PU_ERR_ON_GPU)
            return fail(FME_ERR_INVALID, "PU %d: FME_PU_BI records carry the other list's prediction in err[], not an error grid", i);
          const int os = (int)(pus[i].err[0] & 0xff);
          if (os >= c->cfg.numRefSlots || !c->refValid[os])
            return fail(FME_ERR_STATE, "PU %d: other-list slot %d holds no picture", i, os);
        }
      }
    }
  } else {
    // unvalidated path: nothing is read from the caller's (pinned) records here.  K0 is launched whenever NN_pred runs --
    // it skips records without FME_PU_ERR_ON_GPU itself -- so a mixed batch can never feed K3 an unfilled err[].
    needK0 = (mode & FME_MODE_NN) != 0;
  }
  if (needK0 && !c->orgValid) return fail(FME_ERR_STATE, "FME_PU_ERR_ON_GPU needs fme_upload_org first");
  if (c->fifoCount == FME_NBUF) {  // at most FME_NBUF submits in flight: retire the oldest before reusing its buffers
    int rc = fme_wait_oldest(c);
    if (rc) return rc;
  }
  const int b = (int)(c->submitSeq++ % FME_NBUF);
  fme_pu* d_pus = c->d_pusBuf[b];
  fme_result* d_res = c->d_resBuf[b];
  // records in: the kernels that read d_pus[b] FME_NBUF submits ago must be done
  CU_CHECK(cudaStreamWaitEvent(c->sIn, c->evDone[b], 0));
  if (heads) CU_CHECK(cudaMemcpyAsync(c->d_headBuf[b], heads, sizeof(fme_pu_head) * (size_t)n, cudaMemcpyHostToDevice, c->sIn));
  else if (compact) CU_CHECK(cudaMemcpyAsync(c->d_compactBuf[b], compact, sizeof(fme_pu_compact) * (size_t)n, cudaMemcpyHostToDevice, c->sIn));
  else CU_CHECK(cudaMemcpyAsync(d_pus, pus, sizeof(fme_pu) * (size_t)n, cudaMemcpyHostToDevice, c->sIn));
  if (nGrids > 0)
    CU_CHECK(cudaMemcpyAsync(c->d_gridBuf[b], grids, sizeof(fme_err_grid) * (size_t)nGrids, cudaMemcpyHostToDevice, c->sIn));
  CU_CHECK(cudaEventRecord(c->evIn[b], c->sIn));
  // kernels: need the records, and the previous read-out of d_res[b] must be done
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evIn[b], 0));
  CU_CHECK(cudaStreamWaitEvent(c->stream, c->evOut[b], 0));
  dbg_stamp(c, 3);
  if (heads) CU_CHECK(fme_launch_expand_heads(c->d_headBuf[b], n, d_pus, c->stream, &c->launches));
  if (compact) CU_CHECK(fme_launch_expand_compact(c->d_compactBuf[b], n, d_pus, c->stream, &c->launches));
  if (nGrids > 0) CU_CHECK(fme_launch_apply_grids(c->d_gridBuf[b], nGrids, d_pus, n, c->stream, &c->launches));
  // full records on the unvalidated path with both passes: K0 goes behind K2's record count (see run_search)
  const bool k0BehindK2 = needK0 && !sync && !heads && !compact && mode == FME_MODE_BOTH;
  if (needK0 && !k0BehindK2) {
    StageTimer t(c, 3);
    CU_CHECK(fme_launch_k0(c->g, c->d_planes, c->d_org, d_pus, n, c->cfg.fen, c->stream, &c->launches));
  }
  dbg_stamp(c, 4);
  int rc = run_search(c, d_pus, n, d_res, mode, k0BehindK2);
  if (rc) return rc;
  if (packed) CU_CHECK(fme_launch_pack_results(d_res, n, c->d_res8Buf[b], c->stream, &c->launches));
  dbg_stamp(c, 7);
  CU_CHECK(cudaEventRecord(c->evDone[b], c->stream));
  CU_CHECK(cudaEventRecord(c->evOrgFree[(c->orgSeq + FME_NBUF - 1) % FME_NBUF], c->stream));  // the source buffer this submit read
  // results out
  CU_CHECK(cudaStreamWaitEvent(c->sOut, c->evDone[b], 0));
  if (packed) CU_CHECK(cudaMemcpyAsync(out, c->d_res8Buf[b], sizeof(fme_result8) * (size_t)n, cudaMemcpyDeviceToHost, c->sOut));
  else CU_CHECK(cudaMemcpyAsync(out, d_res, sizeof(fme_result) * (size_t)n, cudaMemcpyDeviceToHost, c->sOut));
  CU_CHECK(cudaEventRecord(c->evOut[b], c->sOut));
  c->fifo[c->fifoCount++] = b;
  if (sync) {
    rc = sync_all(c);
    if (rc) return rc;
    collect_ms(c);
  }
  return FME_OK;
}

int fme_submit(fme_ctx* c, const fme_pu* pus, int n, fme_result* out, int mode) {
  return submit_common(c, pus, n, out, mode, true);
}
int fme_submit_async(fme_ctx* c, const fme_pu* pus, int n, fme_result* out, int mode) {
  return submit_common(c, pus, n, out, mode, false);
}

int fme_submit_heads(fme_ctx* c, const fme_pu_head* heads, int n, fme_result* out, int mode) {
  return submit_common(c, nullptr, n, out, mode, true, heads);
}
int fme_submit_heads_async(fme_ctx* c, const fme_pu_head* heads, int n, fme_result* out, int mode) {
  return submit_common(c, nullptr, n, out, mode, false, heads);
}

int fme_submit_heads_grids(fme_ctx* c, const fme_pu_head* heads, int n, const fme_err_grid* grids, int nGrids,
                           fme_result* out, int mode) {
  return submit_common(c, nullptr, n, out, mode, true, heads, grids, nGrids);
}
int fme_submit_heads_grids_async(fme_ctx* c, const fme_pu_head* heads, int n, const fme_err_grid* grids, int nGrids,
                                 fme_result* out, int mode) {
  return submit_common(c, nullptr, n, out, mode, false, heads, grids, nGrids);
}

int fme_submit_compact(fme_ctx* c, const fme_pu_compact* recs, int n, const fme_err_grid* big, int nBig, fme_result* out,
                       int mode) {
  return submit_common(c, nullptr, n, out, mode, true, nullptr, big, nBig, recs);
}
int fme_submit_compact_async(fme_ctx* c, const fme_pu_compact* recs, int n, const fme_err_grid* big, int nBig,
                             fme_result* out, int mode) {
  return submit_common(c, nullptr, n, out, mode, false, nullptr, big, nBig, recs);
}

int fme_submit_device(fme_ctx* c, const fme_pu* d_pus, int n, fme_result* d_out, int mode) {
  if (!c || !d_pus || !d_out) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0 || n > c->cfg.maxPUs) return fail(FME_ERR_INVALID, "n=%d exceeds maxPUs=%d", n, c->cfg.maxPUs);
  if (n == 0) return FME_OK;
  CU_CHECK(cudaSetDevice(c->cfg.device));
  const bool packed = (mode & FME_MODE_RESULT8) != 0;  // d_out is a device fme_result8 array
  mode &= ~FME_MODE_RESULT8;
  int rc = run_search(c, const_cast<fme_pu*>(d_pus), n, packed ? c->d_resBuf[0] : d_out, mode);
  if (rc) return rc;
  if (packed) CU_CHECK(fme_launch_pack_results(c->d_resBuf[0], n, reinterpret_cast<fme_result8*>(d_out), c->stream, &c->launches));
  CU_CHECK(cudaEventRecord(c->evOrgFree[(c->orgSeq + FME_NBUF - 1) % FME_NBUF], c->stream));
  return FME_OK;
}

int fme_int_surface_device(fme_ctx* c, fme_pu* d_pus, int n) {
  if (!c || !d_pus) return fail(FME_ERR_INVALID, "null argument");
  if (!c->orgValid) return fail(FME_ERR_STATE, "fme_int_surface_device needs fme_upload_org first");
  if (n <= 0) return FME_OK;
  CU_CHECK(cudaSetDevice(c->cfg.device));
  {
    StageTimer t(c, 3);
    CU_CHECK(fme_launch_k0(c->g, c->d_planes, c->d_org, d_pus, n, c->cfg.fen, c->stream, &c->launches));
  }
  CU_CHECK(cudaEventRecord(c->evOrgFree[(c->orgSeq + FME_NBUF - 1) % FME_NBUF], c->stream));
  return FME_OK;
}

// ---- block-level parity entry points -------------------------------------------------------------
static int filter_common(fme_ctx* c, int isVertical, int comp, const int16_t* src, int srcStride, int16_t* dst,
                         int dstStride, int w, int h, int frac, int isFirst, int isLast, int bitDepth) {
  if (!c || !src || !dst) return fail(FME_ERR_INVALID, "null argument");
  if (w < 1 || h < 1 || w > 4096 || h > 4096 || srcStride < w || dstStride < w) return fail(FME_ERR_INVALID, "bad block geometry");
  if (bitDepth < 8 || bitDepth > 12) return fail(FME_ERR_INVALID, "bitDepth out of range");
  bool isLuma = comp == 0;
  int ntaps = isLuma ? 8 : 4;
  if (frac < 0 || frac >= (isLuma ? 4 : 8)) return fail(FME_ERR_INVALID, "frac out of range");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  // copy the block plus the taps' support; like the reference, the caller guarantees it is readable
  int before = frac ? ntaps / 2 - 1 : 0, after = frac ? ntaps / 2 : 0;
  int bx = isVertical ? 0 : before, ax = isVertical ? 0 : after;
  int by = isVertical ? before : 0, ay = isVertical ? after : 0;
  int sw = w + bx + ax, sh = h + by + ay;
  int rc = ensure_pel(&c->d_pel, &c->pelCapacity, (size_t)sw * sh);
  if (rc) return rc;
  rc = ensure_pel(&c->d_pel2, &c->pel2Capacity, (size_t)w * h);
  if (rc) return rc;
  const int16_t* s0 = src - (ptrdiff_t)by * srcStride - bx;
  CU_CHECK(cudaMemcpy2DAsync(c->d_pel, sw * 2, s0, (size_t)srcStride * 2, (size_t)sw * 2, sh, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(fme_launch_filter(isVertical, ntaps, isFirst, isLast, bitDepth, c->d_pel + (size_t)by * sw + bx, sw, c->d_pel2,
                             w, w, h, frac, isLuma, c->stream, &c->launches));
  CU_CHECK(cudaMemcpy2DAsync(dst, (size_t)dstStride * 2, c->d_pel2, (size_t)w * 2, (size_t)w * 2, h, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

int fme_filter_hor(fme_ctx* c, int comp, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h,
                   int frac, int isLast, int bitDepth) {
  return filter_common(c, 0, comp, src, srcStride, dst, dstStride, w, h, frac, 1, isLast, bitDepth);
}
int fme_filter_ver(fme_ctx* c, int comp, const int16_t* src, int srcStride, int16_t* dst, int dstStride, int w, int h,
                   int frac, int isFirst, int isLast, int bitDepth) {
  return filter_common(c, 1, comp, src, srcStride, dst, dstStride, w, h, frac, isFirst, isLast, bitDepth);
}

int fme_dist(fme_ctx* c, int kind, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h,
             int bitDepth, int subShift, int nBlocks, uint32_t* out) {
  if (!c || !org || !cur || !out) return fail(FME_ERR_INVALID, "null argument");
  if (kind < 0 || kind > 2 || w < 2 || h < 2 || w > 64 || h > 64 || orgStride < w || curStride < w || nBlocks < 1 ||
      bitDepth < 8 || bitDepth > 12 || subShift < 0 || subShift > 4)
    return fail(FME_ERR_INVALID, "bad fme_dist arguments");
  CU_CHECK(cudaSetDevice(c->cfg.device));
  size_t no = (size_t)nBlocks * h * orgStride, nc = (size_t)nBlocks * h * curStride;
  int rc = ensure_pel(&c->d_pel, &c->pelCapacity, no);
  if (rc) return rc;
  rc = ensure_pel(&c->d_pel2, &c->pel2Capacity, nc);
  if (rc) return rc;
  rc = ensure_u32(c, nBlocks);
  if (rc) return rc;
  CU_CHECK(cudaMemcpyAsync(c->d_pel, org, no * 2, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(c->d_pel2, cur, nc * 2, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(fme_launch_dist(kind, c->d_pel, orgStride, c->d_pel2, curStride, w, h, bitDepth, subShift, nBlocks,
                           c->d_scratchU32, c->stream, &c->launches));
  CU_CHECK(cudaMemcpyAsync(out, c->d_scratchU32, sizeof(uint32_t) * nBlocks, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

// ---- motion compensation ---------------------------------------------------------------------------
int fme_mc(fme_ctx* c, const fme_mc_pu* pus, int n, int16_t* dstY, int16_t* dstCb, int16_t* dstCr) {
  if (!c || !pus || !dstY) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0 || n > c->cfg.maxPUs) return fail(FME_ERR_INVALID, "n out of range");
  if (n == 0) return FME_OK;
  bool chroma = dstCb && dstCr;
  if (chroma && !c->d_cb) return fail(FME_ERR_STATE, "fme_mc with chroma needs fme_upload_ref_chroma first");
  for (int i = 0; i < n; ++i) {
    if (pus[i].refSlot >= c->cfg.numRefSlots || !c->refValid[pus[i].refSlot])
      return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, pus[i].refSlot);
    if (pus[i].w > 64 || pus[i].h > 64 || pus[i].w < 4 || pus[i].h < 4) return fail(FME_ERR_INVALID, "PU %d: bad size", i);
  }
  CU_CHECK(cudaSetDevice(c->cfg.device));
  {
    int rc0 = sync_all(c);  // the record staging buffer is shared with the async submit path
    if (rc0) return rc0;
  }
  size_t ySamples = (size_t)n * 64 * 64, cSamples = (size_t)n * 32 * 32;
  int rc = ensure_pel(&c->d_pel, &c->pelCapacity, ySamples);
  if (rc) return rc;
  if (chroma) {
    rc = ensure_pel(&c->d_pel2, &c->pel2Capacity, 2 * cSamples);
    if (rc) return rc;
  }
  static_assert(sizeof(fme_mc_pu) <= sizeof(fme_pu), "record staging reuses d_pus");
  CU_CHECK(cudaMemcpyAsync(c->d_pus, pus, sizeof(fme_mc_pu) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(fme_launch_mc(c->g, c->d_planes, chroma ? c->d_cb : nullptr, chroma ? c->d_cr : nullptr,
                         reinterpret_cast<const fme_mc_pu*>(c->d_pus), n, c->d_pel, chroma ? c->d_pel2 : nullptr,
                         chroma ? c->d_pel2 + cSamples : nullptr, c->stream, &c->launches));
  CU_CHECK(cudaMemcpyAsync(dstY, c->d_pel, ySamples * 2, cudaMemcpyDeviceToHost, c->stream));
  if (chroma) {
    CU_CHECK(cudaMemcpyAsync(dstCb, c->d_pel2, cSamples * 2, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(cudaMemcpyAsync(dstCr, c->d_pel2 + cSamples, cSamples * 2, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

int fme_mc_bi(fme_ctx* c, const fme_mc_bi_pu* pus, int n, int16_t* dstY, int16_t* dstCb, int16_t* dstCr) {
  if (!c || !pus || !dstY) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0 || n > c->cfg.maxPUs) return fail(FME_ERR_INVALID, "n out of range");
  if (n == 0) return FME_OK;
  bool chroma = dstCb && dstCr;
  if (chroma && !c->d_cb) return fail(FME_ERR_STATE, "fme_mc_bi with chroma needs fme_upload_ref_chroma first");
  for (int i = 0; i < n; ++i) {
    const int sl[2] = {pus[i].refSlot0, pus[i].refSlot1};
    for (int l = 0; l < 2; ++l)
      if (sl[l] >= c->cfg.numRefSlots || !c->refValid[sl[l]])
        return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, sl[l]);
    if (pus[i].w > 64 || pus[i].h > 64 || pus[i].w < 4 || pus[i].h < 4) return fail(FME_ERR_INVALID, "PU %d: bad size", i);
  }
  CU_CHECK(cudaSetDevice(c->cfg.device));
  {
    int rc0 = sync_all(c);  // the record staging buffer is shared with the async submit path
    if (rc0) return rc0;
  }
  size_t ySamples = (size_t)n * 64 * 64, cSamples = (size_t)n * 32 * 32;
  int rc = ensure_pel(&c->d_pel, &c->pelCapacity, ySamples);
  if (rc) return rc;
  if (chroma) {
    rc = ensure_pel(&c->d_pel2, &c->pel2Capacity, 2 * cSamples);
    if (rc) return rc;
  }
  static_assert(sizeof(fme_mc_bi_pu) <= sizeof(fme_pu), "record staging reuses d_pus");
  CU_CHECK(cudaMemcpyAsync(c->d_pus, pus, sizeof(fme_mc_bi_pu) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(fme_launch_mc_bi(c->g, c->d_planes, chroma ? c->d_cb : nullptr, chroma ? c->d_cr : nullptr,
                            reinterpret_cast<const fme_mc_bi_pu*>(c->d_pus), n, c->d_pel, chroma ? c->d_pel2 : nullptr,
                            chroma ? c->d_pel2 + cSamples : nullptr, c->stream, &c->launches));
  CU_CHECK(cudaMemcpyAsync(dstY, c->d_pel, ySamples * 2, cudaMemcpyDeviceToHost, c->stream));
  if (chroma) {
    CU_CHECK(cudaMemcpyAsync(dstCb, c->d_pel2, cSamples * 2, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(cudaMemcpyAsync(dstCr, c->d_pel2 + cSamples, cSamples * 2, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

int fme_pred_error(fme_ctx* c, const fme_mc_pu* pus, int n, uint32_t* out) {
  if (!c || !pus || !out) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0 || n > c->cfg.maxPUs) return fail(FME_ERR_INVALID, "n out of range");
  if (n == 0) return FME_OK;
  if (!c->orgValid) return fail(FME_ERR_STATE, "fme_pred_error needs fme_upload_org first");
  for (int i = 0; i < n; ++i) {
    if (pus[i].refSlot >= c->cfg.numRefSlots || !c->refValid[pus[i].refSlot])
      return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, pus[i].refSlot);
    if (!fme_hevc_pu_shape(pus[i].w, pus[i].h)) return fail(FME_ERR_INVALID, "PU %d: %dx%d is not an HEVC PU size", i, pus[i].w, pus[i].h);
  }
  CU_CHECK(cudaSetDevice(c->cfg.device));
  int rc = sync_all(c);  // the record staging buffer is shared with the async submit path
  if (rc) return rc;
  if ((rc = ensure_u32(c, n))) return rc;
  CU_CHECK(cudaMemcpyAsync(c->d_pus, pus, sizeof(fme_mc_pu) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(fme_launch_pred_error(c->g, c->d_planes, c->d_org, reinterpret_cast<const fme_mc_pu*>(c->d_pus), n,
                                 c->cfg.useHadME, c->d_scratchU32, c->stream, &c->launches));
  CU_CHECK(cudaMemcpyAsync(out, c->d_scratchU32, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

// ---- batched candidate costs / compact MC ---------------------------------------------------------------
static int ensure_bytes(void** p, size_t* cap, size_t bytes) {
  if (*cap >= bytes) return FME_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CU_CHECK(cudaMalloc(p, bytes));
  *cap = bytes;
  return FME_OK;
}

int fme_cand_cost_device(fme_ctx* c, const fme_cand_pu* d_cands, int n, uint32_t* d_cost, int32_t* d_best) {
  if (!c || !d_cands || !d_cost) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0) return fail(FME_ERR_INVALID, "n out of range");
  if (!c->orgValid || !c->sliceValid) return fail(FME_ERR_STATE, "fme_cand_cost needs fme_upload_org and fme_set_slice first");
  if (n == 0) return FME_OK;
  CU_CHECK(cudaSetDevice(c->cfg.device));
  CU_CHECK(fme_launch_cand_cost(c->g, c->d_planes, c->d_org, d_cands, n, c->costLut, c->cfg.useHadME, d_cost, d_best, c->stream,
                                &c->launches));
  CU_CHECK(cudaEventRecord(c->evOrgFree[(c->orgSeq + FME_NBUF - 1) % FME_NBUF], c->stream));
  return FME_OK;
}

int fme_cand_cost(fme_ctx* c, const fme_cand_pu* cands, int n, uint32_t* cost, int32_t* best) {
  if (!c || !cands || !cost) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0) return fail(FME_ERR_INVALID, "n out of range");
  if (n == 0) return FME_OK;
  for (int i = 0; i < n; ++i) {
    if (cands[i].refSlot >= c->cfg.numRefSlots || !c->refValid[cands[i].refSlot])
      return fail(FME_ERR_STATE, "candidate %d references slot %d which holds no picture", i, cands[i].refSlot);
    if (!fme_hevc_pu_shape(cands[i].w, cands[i].h)) return fail(FME_ERR_INVALID, "candidate %d: %dx%d is not an HEVC PU size", i, cands[i].w, cands[i].h);
    if (cands[i].bits >= FME_COST_LUT_SIZE) return fail(FME_ERR_INVALID, "candidate %d: bits out of range", i);
  }
  CU_CHECK(cudaSetDevice(c->cfg.device));
  const size_t inB = sizeof(fme_cand_pu) * (size_t)n, outB = 4 * (size_t)n;
  int rc = ensure_bytes(&c->d_opBuf, &c->opBufCap, inB + 2 * outB);
  if (rc) return rc;
  fme_cand_pu* d_c = static_cast<fme_cand_pu*>(c->d_opBuf);
  uint32_t* d_cost = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(c->d_opBuf) + inB);
  int32_t* d_best = reinterpret_cast<int32_t*>(d_cost + n);
  // stream-ordered on the kernel stream; only this call's own copies are waited for (no pipeline drain)
  CU_CHECK(cudaMemcpyAsync(d_c, cands, inB, cudaMemcpyHostToDevice, c->stream));
  if ((rc = fme_cand_cost_device(c, d_c, n, d_cost, best ? d_best : nullptr))) return rc;
  CU_CHECK(cudaMemcpyAsync(cost, d_cost, outB, cudaMemcpyDeviceToHost, c->stream));
  if (best) CU_CHECK(cudaMemcpyAsync(best, d_best, outB, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

int fme_mc_luma_compact_device(fme_ctx* c, const fme_mc_pu* d_pus, int n, const uint32_t* d_offsets, uint8_t* d_out) {
  if (!c || !d_pus || !d_offsets || !d_out) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0) return fail(FME_ERR_INVALID, "n out of range");
  if (n == 0) return FME_OK;
  CU_CHECK(cudaSetDevice(c->cfg.device));
  CU_CHECK(fme_launch_mc_luma_compact(c->g, c->d_planes, d_pus, n, d_offsets, d_out, c->stream, &c->launches));
  return FME_OK;
}

int fme_mc_luma_compact(fme_ctx* c, const fme_mc_pu* pus, int n, const uint32_t* offsets, uint8_t* out, size_t outBytes) {
  if (!c || !pus || !offsets || !out) return fail(FME_ERR_INVALID, "null argument");
  if (n < 0) return fail(FME_ERR_INVALID, "n out of range");
  if (n == 0) return FME_OK;
  for (int i = 0; i < n; ++i) {
    if (pus[i].refSlot >= c->cfg.numRefSlots || !c->refValid[pus[i].refSlot])
      return fail(FME_ERR_STATE, "PU %d references slot %d which holds no picture", i, pus[i].refSlot);
    if (pus[i].w > 64 || pus[i].h > 64 || pus[i].w < 4 || pus[i].h < 4 || (pus[i].w & 3)) return fail(FME_ERR_INVALID, "PU %d: bad size", i);
    if ((offsets[i] & 3) || (size_t)offsets[i] + (size_t)pus[i].w * pus[i].h > outBytes)
      return fail(FME_ERR_INVALID, "PU %d: offset %u misaligned or block beyond outBytes", i, offsets[i]);
  }
  CU_CHECK(cudaSetDevice(c->cfg.device));
  const size_t inB = (sizeof(fme_mc_pu) * (size_t)n + 15) & ~(size_t)15, offB = (4 * (size_t)n + 15) & ~(size_t)15;
  int rc = ensure_bytes(&c->d_opBuf, &c->opBufCap, inB + offB + outBytes);
  if (rc) return rc;
  uint8_t* base = static_cast<uint8_t*>(c->d_opBuf);
  CU_CHECK(cudaMemcpyAsync(base, pus, sizeof(fme_mc_pu) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(cudaMemcpyAsync(base + inB, offsets, 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  if ((rc = fme_mc_luma_compact_device(c, reinterpret_cast<const fme_mc_pu*>(base), n, reinterpret_cast<const uint32_t*>(base + inB),
                                       base + inB + offB)))
    return rc;
  CU_CHECK(cudaMemcpyAsync(out, base + inB + offB, outBytes, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

// ---- introspection -----------------------------------------------------------------------------------
int fme_download_plane(fme_ctx* c, int slot, int fy, int fx, uint8_t* dst, int dstStride) {
  int rc = check_slot(c, slot);
  if (rc) return rc;
  if (!dst || fy < 0 || fy > 3 || fx < 0 || fx > 3 || dstStride < c->g.Wp) return fail(FME_ERR_INVALID, "bad argument");
  if (!c->refValid[slot]) return fail(FME_ERR_STATE, "slot %d holds no picture", slot);
  CU_CHECK(cudaSetDevice(c->cfg.device));
  const uint8_t* src = c->d_planes + (size_t)slot * c->g.slotBytes + (size_t)(fy * 4 + fx) * c->g.planeBytes;
  CU_CHECK(cudaMemcpy2DAsync(dst, dstStride, src, c->g.pitch, c->g.Wp, c->g.Hp, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(cudaStreamSynchronize(c->stream));
  return FME_OK;
}

}  // extern "C"
