// edt_capi.cu -- host side of the C ABI declared in include/edt_b200.h.
//
// Plays the role of the reference's volume drivers (pyedt::_edt3dsq / _edt2dsq,
// src/edt.hpp:411-484, 632-678): it owns the pass order X -> Y -> Z over one float32
// volume that is transformed in place, but the "thread pool" is the CUDA grid and the
// passes are stream-ordered kernel launches.  No CPU fallback exists in this file.
#include "../../include/edt_b200.h"
#include "edt_kernels.cuh"
#include "edt_voxel_graph.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace {

thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return fail(e__ == cudaErrorMemoryAllocation ? EDTB200_ENOMEM : EDTB200_ECUDA, \
                  "%s failed: %s", #expr, cudaGetErrorString(e__));                 \
  } while (0)

constexpr int kMaxDevices = 64;

// cuTensorMapEncodeTiled, fetched through the runtime so that libcuda is not a link-time
// dependency (the library must load on machines without a driver, e.g. for the build check).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_encode_tried = false;

EncodeTiledFn tensor_map_encoder() {
  if (!g_encode_tried) {
    g_encode_tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    else
      cudaGetLastError();
  }
  return g_encode;
}

// Per-device cached state for calls that take host buffers.
struct DeviceCache {
  void* labels = nullptr;
  size_t labels_bytes = 0;
  float* dist = nullptr;
  size_t dist_bytes = 0;
  cudaStream_t stream = nullptr;
  // second slot + copy streams + events of edtb200_transform_batch (slot 0 is labels / dist above)
  void* labels2 = nullptr;
  size_t labels2_bytes = 0;
  float* dist2 = nullptr;
  size_t dist2_bytes = 0;
  cudaStream_t stream_up = nullptr, stream_down = nullptr;
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_down[2] = {nullptr, nullptr};
  // step tables T[k] of the first-axis pass, keyed by the weight's bits (see step_table_kernel)
  struct Table { float* data = nullptr; int count = 0; uint32_t wbits = 0; cudaEvent_t ready = nullptr;
                 cudaStream_t built_on = nullptr; uint64_t stamp = 0; };
  Table tables[8];
  uint64_t table_clock = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  bool probed = false;
};

std::mutex g_mutex;
DeviceCache g_cache[kMaxDevices];

// ---- host <-> device staging for pageable host memory --------------------------------
// numpy arrays are pageable: a plain cudaMemcpy of 512 MiB then runs at a fraction of the PCIe
// rate (the driver stages it through small internal buffers, and a freshly allocated output
// array is also page-faulted in by that single thread).  Pinned callers are copied directly;
// pageable ones go through three 32 MiB pinned buffers that a few host threads fill / drain in
// parallel while the DMA engine moves the previous chunk.
class CopyPool {
 public:
  explicit CopyPool(int n) : n_(n) {
    for (int i = 0; i < n_; ++i) threads_.emplace_back([this, i] { loop(i); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; ++epoch_; }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int size() const { return n_; }
  // run fn(i) for i in [0, n) on the pool and wait
  void run(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> l(m_);
    fn_ = &fn; pending_ = n_; ++epoch_;
    cv_.notify_all();
    done_.wait(l, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }
 private:
  void loop(int i) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int)>* fn;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        fn = fn_;
      }
      (*fn)(i);
      {
        std::lock_guard<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  unsigned long epoch_ = 0;
  bool stop_ = false;
};

// Two pools and two sets of staging buffers: [0] for copies towards the device, [1] for copies
// back, so that a batch (edtb200_transform_batch) can drive both directions at once from two
// host threads.  Created on first use; only ever touched under g_mutex or by the batch's helper.
CopyPool* g_pools[2] = {nullptr, nullptr};

void parallel_memcpy(void* dst, const void* src, size_t bytes, int dir) {
  CopyPool*& g_pool = g_pools[dir];
  if (!g_pool) {
    unsigned hw = std::thread::hardware_concurrency();
    // measured on the B200 host (2 x 64 threads), 512 MiB each way: 8 threads 40 ms, 16 threads
    // 31 ms, 32 threads 38 ms per numpy-to-numpy call; populating the fresh output array's pages
    // from helper threads during the upload was tried and only made it slower (44-68 ms)
    int n = hw >= 32 ? 16 : (hw >= 16 ? 8 : (hw >= 4 ? 4 : 1));
    if (const char* e = getenv("EDTB200_COPY_THREADS")) n = std::max(1, std::min(64, atoi(e)));
    g_pool = new CopyPool(n);
  }
  const int n = g_pool->size();
  const size_t slice = ((bytes + n - 1) / n + 4095) & ~size_t(4095);
  g_pool->run([&](int i) {
    const size_t off = slice * (size_t)i;
    if (off < bytes) memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off,
                            bytes - off < slice ? bytes - off : slice);
  });
}

constexpr size_t kStageBytes = size_t(32) << 20;
constexpr int kStages = 3;
struct StageBuffers {
  void* buf[kStages] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[kStages] = {nullptr, nullptr, nullptr};
};
StageBuffers g_stage[2][kMaxDevices];

bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int ensure_stage(int device, int dir) {
  StageBuffers& sb = g_stage[dir][device];
  for (int i = 0; i < kStages; ++i) {
    if (!sb.buf[i]) CUDA_TRY(cudaHostAlloc(&sb.buf[i], kStageBytes, cudaHostAllocDefault));
    if (!sb.ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&sb.ev[i], cudaEventDisableTiming));
  }
  return 0;
}

// host (pageable or pinned) -> device, stream-ordered; returns after the last chunk is queued
int upload(void* dst_dev, const void* src_host, size_t bytes, int device, cudaStream_t stream) {
  if (bytes < (size_t(4) << 20) || host_pointer_is_pinned(src_host)) {
    CUDA_TRY(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  }
  int rc = ensure_stage(device, 0);
  if (rc) return rc;
  StageBuffers& sb = g_stage[0][device];
  size_t off = 0;
  for (int k = 0; off < bytes; ++k) {
    const int b = k % kStages;
    const size_t n = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    // the DMA out of this buffer must have finished -- also the one queued by an EARLIER upload:
    // inside a batch the previous volume's copies may still be waiting in the stream
    CUDA_TRY(cudaEventSynchronize(sb.ev[b]));
    parallel_memcpy(sb.buf[b], static_cast<const char*>(src_host) + off, n, 0);
    CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(dst_dev) + off, sb.buf[b], n, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaEventRecord(sb.ev[b], stream));
    off += n;
  }
  return 0;
}

// device -> host (pageable or pinned); complete on return for the pageable case
int download(void* dst_host, const void* src_dev, size_t bytes, int device, cudaStream_t stream) {
  if (bytes < (size_t(4) << 20) || host_pointer_is_pinned(dst_host)) {
    CUDA_TRY(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream));
    return 0;
  }
  int rc = ensure_stage(device, 1);
  if (rc) return rc;
  StageBuffers& sb = g_stage[1][device];
  size_t off = 0, prev_off = 0, prev_n = 0;
  int prev_b = -1;
  for (int k = 0; off < bytes || prev_b >= 0; ++k) {
    int b = -1;
    size_t n = 0;
    if (off < bytes) {
      b = k % kStages;
      n = bytes - off < kStageBytes ? bytes - off : kStageBytes;
      CUDA_TRY(cudaMemcpyAsync(sb.buf[b], static_cast<const char*>(src_dev) + off, n, cudaMemcpyDeviceToHost, stream));
      CUDA_TRY(cudaEventRecord(sb.ev[b], stream));
    }
    if (prev_b >= 0) {                                             // drain the chunk queued one step earlier
      CUDA_TRY(cudaEventSynchronize(sb.ev[prev_b]));
      parallel_memcpy(static_cast<char*>(dst_host) + prev_off, sb.buf[prev_b], prev_n, 1);
    }
    prev_b = b; prev_off = off; prev_n = n;
    off += n;
  }
  return 0;
}

// Restores the calling thread's current CUDA device when an entry point returns (callers such as
// PyTorch rely on "their" current device staying put).
struct DeviceGuard {
  int saved = -1;
  DeviceGuard() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
  ~DeviceGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

int probe(int device, DeviceCache** out) {
  if (device < 0 || device >= kMaxDevices) return fail(EDTB200_EINVAL, "bad device %d", device);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(EDTB200_ECUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device >= count) return fail(EDTB200_EINVAL, "device %d out of range (%d visible)", device, count);
  DeviceCache& dc = g_cache[device];
  CUDA_TRY(cudaSetDevice(device));
  if (!dc.probed) {
    CUDA_TRY(cudaDeviceGetAttribute(&dc.sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&dc.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // keep stream-ordered allocations (the code plane, long-line temporaries) in the pool between
    // calls instead of returning them to the driver at every synchronisation
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    dc.probed = true;
  }
  *out = &dc;
  return 0;
}

int check_dims(int label_bytes, int ndim, int64_t& sx, int64_t& sy, int64_t& sz) {
  if (!(label_bytes == 1 || label_bytes == 2 || label_bytes == 4 || label_bytes == 8))
    return fail(EDTB200_EINVAL, "label_bytes must be 1, 2, 4 or 8 (got %d)", label_bytes);
  if (ndim < 1 || ndim > 3) return fail(EDTB200_EINVAL, "ndim must be 1, 2 or 3 (got %d)", ndim);
  if (ndim < 3) sz = 1;
  if (ndim < 2) sy = 1;
  if (sx < 0 || sy < 0 || sz < 0) return fail(EDTB200_EINVAL, "negative size");
  const int64_t lim = (int64_t)1 << 30;
  if (sx > lim || sy > lim || sz > lim) return fail(EDTB200_ELIMIT, "axis longer than 2^30");
  return 0;
}

// Device table T[0..count) for weight w, cached per device.  Built once on `stream`; other
// streams wait on the build event, so no host synchronisation and no per-call allocation.
int step_table(DeviceCache& dc, float w, int count, cudaStream_t stream, const float** out) {
  uint32_t wbits;
  memcpy(&wbits, &w, sizeof(wbits));
  DeviceCache::Table* hit = nullptr;
  DeviceCache::Table* victim = &dc.tables[0];
  for (auto& t : dc.tables) {
    if (t.data && t.wbits == wbits && t.count >= count) { hit = &t; break; }
    if (t.stamp < victim->stamp) victim = &t;
  }
  if (!hit) {
    DeviceCache::Table& t = *victim;
    if (t.data) { CUDA_TRY(cudaDeviceSynchronize()); cudaFree(t.data); t.data = nullptr; }
    const int cap = count < 4096 ? 4096 : count;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&t.data), sizeof(float) * (size_t)cap));
    if (!t.ready) CUDA_TRY(cudaEventCreateWithFlags(&t.ready, cudaEventDisableTiming));
    edtb200::step_table_kernel<<<1, 32, 0, stream>>>(w, cap, t.data);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(t.ready, stream));
    t.count = cap; t.wbits = wbits; t.built_on = stream;
    hit = &t;
  } else if (hit->built_on != stream) {
    CUDA_TRY(cudaStreamWaitEvent(stream, hit->ready, 0));
  }
  hit->stamp = ++dc.table_clock;
  *out = hit->data;
  return 0;
}

// ---- launches ---------------------------------------------------------------------

// codes != nullptr asks for the one-byte neighbour codes (see first_axis_vec_kernel); *codes_done
// reports whether they were produced (only the vector kernel can).
template <int Bytes>
int launch_first(const void* labels, float* f, int64_t nlines, int64_t sx, float w, int border,
                 int flags, DeviceCache& dc, cudaStream_t stream, uint8_t* codes = nullptr, int64_t sy = 1,
                 bool* codes_done = nullptr) {
  if (codes_done) *codes_done = false;
  using namespace edtb200;
  const float* table = nullptr;
  int trc = step_table(dc, w, (int)sx + 1, stream, &table);
  if (trc) return trc;

  using LT = typename LabelOf<Bytes>::type;
  // register-resident vector kernel when rows are short and 16-byte aligned
  if (sx % 4 == 0 && sx <= 1024 && reinterpret_cast<uintptr_t>(labels) % (4 * Bytes) == 0 &&
      reinterpret_cast<uintptr_t>(f) % 16 == 0) {
    const size_t smem = sizeof(float) * (size_t)(sx + 2);
    int64_t blocks = (nlines + 7) / 8;
    const int64_t cap = (int64_t)dc.sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const LT* lab = static_cast<const LT*>(labels);
    const bool want_codes = codes != nullptr && reinterpret_cast<uintptr_t>(codes) % 4 == 0;
#define EDT_LAUNCH_VEC(KK)                                                                         \
  do {                                                                                           \
    if (want_codes) {                                                                            \
      if (flags == 0)                                                                            \
        first_axis_vec_kernel<Bytes, KK, true, true><<<(unsigned)blocks, 256, smem, stream>>>(    \
            lab, f, nlines, (int)sx, table, border, flags, codes, (int)sy);                      \
      else                                                                                       \
        first_axis_vec_kernel<Bytes, KK, false, true><<<(unsigned)blocks, 256, smem, stream>>>(   \
            lab, f, nlines, (int)sx, table, border, flags, codes, (int)sy);                      \
    } else {                                                                                     \
      if (flags == 0)                                                                            \
        first_axis_vec_kernel<Bytes, KK, true, false><<<(unsigned)blocks, 256, smem, stream>>>(   \
            lab, f, nlines, (int)sx, table, border, flags, nullptr, 1);                          \
      else                                                                                       \
        first_axis_vec_kernel<Bytes, KK, false, false><<<(unsigned)blocks, 256, smem, stream>>>(  \
            lab, f, nlines, (int)sx, table, border, flags, nullptr, 1);                          \
    }                                                                                            \
  } while (0)
    if (sx <= 128)      EDT_LAUNCH_VEC(1);
    else if (sx <= 256) EDT_LAUNCH_VEC(2);
    else if (sx <= 512) EDT_LAUNCH_VEC(4);
    else                EDT_LAUNCH_VEC(8);
#undef EDT_LAUNCH_VEC
    CUDA_TRY(cudaGetLastError());
    if (codes_done) *codes_done = want_codes;
    return 0;
  }
  const int nwords = (int)(sx >> 5) + 1;
  const size_t per_warp = sizeof(uint32_t) * 4 * (size_t)nwords;
  int warps = 8;
  while (warps > 1 && per_warp * warps > (size_t)dc.max_smem_optin) warps >>= 1;
  if (per_warp * warps > (size_t)dc.max_smem_optin) {
    return fail(EDTB200_ELIMIT, "first axis of %lld voxels exceeds the shared-memory line buffer",
                (long long)sx);
  }
  const size_t smem = per_warp * warps;
  auto kern = first_axis_kernel<Bytes>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (nlines + warps - 1) / warps;
  const int64_t cap = (int64_t)dc.sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, warps * 32, smem, stream>>>(
      static_cast<const typename LabelOf<Bytes>::type*>(labels), f, nlines, (int)sx, table, border, flags);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// Tensor map over the distance volume for one later-axis pass: dims (adjacent lines, line
// length, outer), box = tx lines x box_rows.
bool make_tile_map(CUtensorMap* map, float* f, const edtb200::LineGeom& g, int tx, int box_rows) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)g.inner_count, (cuuint64_t)g.n, (cuuint64_t)g.outer_count};
  const cuuint64_t strides[2] = {(cuuint64_t)g.line_stride * sizeof(float),
                                 (cuuint64_t)(g.outer_count > 1 ? g.outer_stride : g.line_stride * (int64_t)g.n) *
                                     sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)tx, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, f, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Shared memory of one tile of `tx` lines (see later_axis_tile_kernel).
size_t tile_smem_bytes(int n, int tx, int rows_alloc) {
  const int nchunks = (n + 31) >> 5;
  return (size_t)rows_alloc * tx * 4 + (size_t)nchunks * tx * 12 + (size_t)((n + 3) & ~1) * 4 + 16 +
         (size_t)nchunks * tx;     // + one flag byte per (chunk, line)
}

template <int Bytes, int TX>
int launch_tile(const void* labels, float* f, edtb200::LineGeom g, float w2, int border_lo, int border_hi,
                int flags, bool use_tma, cudaStream_t stream, int code_bit = 0, bool pdl = false) {
  using namespace edtb200;
  using LT = typename LabelOf<Bytes>::type;
  const int nchunks = (g.n + 31) >> 5;
  TileBoxes tb;
  tb.nboxes = (g.n + 255) / 256;
  tb.box_rows = (g.n + tb.nboxes - 1) / tb.nboxes;
  if (tb.nboxes > 1) tb.box_rows = (tb.box_rows + 3) & ~3;      // keeps every box 128-byte aligned
  g.tiles_per_outer = (int)((g.inner_count + TX - 1) / TX);
  const int64_t tiles = (int64_t)g.tiles_per_outer * g.outer_count;
  if (tiles > 0x7fffffffLL) return fail(EDTB200_ELIMIT, "too many line tiles");
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (use_tma && !make_tile_map(&map, f, g, TX, tb.box_rows)) use_tma = false;
  const int rows_alloc = use_tma ? tb.box_rows * tb.nboxes : g.n;
  const size_t smem = tile_smem_bytes(g.n, TX, rows_alloc);
  constexpr int SUBS = 32 / TX;
  int warps = (nchunks + SUBS - 1) / SUBS;
  const bool wide = warps > 16;                 // long lines: one tile per SM, so give it 32 warps
  if (warps > 32) warps = 32;
  const LT* lab = static_cast<const LT*>(labels);
  // Programmatic dependent launch: this pass may begin (label staging) while the previous pass of
  // the stream drains its last wave; the kernel itself waits before touching the distances.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3((unsigned)(warps * 32));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = pdl_attr;
  static const bool pdl_off = getenv("EDTB200_NO_PDL") != nullptr;    // A/B switch for measurements
  cfg.numAttrs = (pdl && !pdl_off) ? 1 : 0;   // only when the previous kernel of the stream is our own pass
#define EDT_LAUNCH_TILE(EPI, TMA, CODES)                                                            \
  do {                                                                                              \
    if (wide) {                                                                                     \
      auto kern = later_axis_tile_kernel<Bytes, TX, EPI, TMA, CODES, true>;                         \
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, map, lab, f, g, tb, w2, border_lo, border_hi, flags, code_bit)); \
    } else {                                                                                        \
      auto kern = later_axis_tile_kernel<Bytes, TX, EPI, TMA, CODES, false>;                        \
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, map, lab, f, g, tb, w2, border_lo, border_hi, flags, code_bit)); \
    }                                                                                               \
  } while (0)
  if constexpr (Bytes == 1) {
    if (code_bit) {
      if (flags) { if (use_tma) EDT_LAUNCH_TILE(true, true, true); else EDT_LAUNCH_TILE(true, false, true); }
      else       { if (use_tma) EDT_LAUNCH_TILE(false, true, true); else EDT_LAUNCH_TILE(false, false, true); }
      CUDA_TRY(cudaGetLastError());
      return 0;
    }
  }
  if (flags) { if (use_tma) EDT_LAUNCH_TILE(true, true, false); else EDT_LAUNCH_TILE(true, false, false); }
  else       { if (use_tma) EDT_LAUNCH_TILE(false, true, false); else EDT_LAUNCH_TILE(false, false, false); }
#undef EDT_LAUNCH_TILE
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// code_bit != 0: `labels` are the one-byte codes of the first-axis pass (Bytes must be 1); the
// caller must have checked tile_path_ok() because only the tile kernel understands codes.
template <int Bytes>
int launch_later(const void* labels, float* f, const edtb200::LineGeom& g0, float w, int border_lo,
                 int border_hi, int flags, const DeviceCache& dc, cudaStream_t stream, int code_bit = 0,
                 bool pdl = false) {
  using namespace edtb200;
  using LT = typename LabelOf<Bytes>::type;
  LineGeom g = g0;
  const float w2 = w * w;                       // float product, as src/edt.hpp:181

  // ---- shared-memory tile kernel: whole lines x TX adjacent lines per CTA ----
  const bool fits32 = (int64_t)g.n * g.line_stride + 64 < (1LL << 32);
  if (fits32 && g.n <= 4096 && g.inner_count < (1LL << 31)) {
    const bool aligned = reinterpret_cast<uintptr_t>(f) % 16 == 0 && g.line_stride % 4 == 0 &&
                         (g.outer_count <= 1 || g.outer_stride % 4 == 0);
    // Tile width: 128-byte rows (TX = 32) keep DRAM pages and L2 lines whole and measured
    // fastest even at one CTA per SM; narrower tiles only when a 32-wide tile cannot fit.
    int tx = 0;
    for (int cand = 32; cand >= 8 && !tx; cand >>= 1) {
      const int nb = (g.n + 255) / 256;
      int br = (g.n + nb - 1) / nb;
      if (nb > 1) br = (br + 3) & ~3;
      if (tile_smem_bytes(g.n, cand, br * nb) <= (size_t)dc.max_smem_optin) tx = cand;
    }
    if (tx) {
      const bool use_tma = aligned && g.inner_count >= tx;
      switch (tx) {
        case 32: return launch_tile<Bytes, 32>(labels, f, g, w2, border_lo, border_hi, flags, use_tma, stream, code_bit, pdl);
        case 16: return launch_tile<Bytes, 16>(labels, f, g, w2, border_lo, border_hi, flags, use_tma, stream, code_bit, pdl);
        default: return launch_tile<Bytes, 8>(labels, f, g, w2, border_lo, border_hi, flags, use_tma, stream, code_bit, pdl);
      }
    }
  }
  if (code_bit) return fail(EDTB200_ELIMIT, "internal: neighbour codes need the tile kernel");
  // ---- lines too long for a shared-memory tile: out of place through a temporary volume ----
  const int64_t lines = g.inner_count * g.outer_count;
  const size_t bytes = sizeof(float) * (size_t)lines * (size_t)g.n;
  float* tmp = nullptr;
  int* hull = nullptr;
  CUDA_TRY(cudaMallocAsync(&tmp, bytes, stream));
  if (cudaMallocAsync(&hull, bytes, stream) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeAsync(tmp, stream);
    return fail(EDTB200_ENOMEM, "no device memory for the long-line scratch volumes");
  }
  const int64_t blocks = (lines + 127) / 128;
  if (blocks > 0x7fffffffLL) {
    cudaFreeAsync(tmp, stream); cudaFreeAsync(hull, stream);
    return fail(EDTB200_ELIMIT, "too many lines");
  }
  later_axis_long_kernel<Bytes><<<(unsigned)blocks, 128, 0, stream>>>(
      static_cast<const LT*>(labels), f, tmp, hull, g, w2, border_lo, border_hi, flags);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(f, tmp, bytes, cudaMemcpyDeviceToDevice, stream));
  CUDA_TRY(cudaFreeAsync(hull, stream));
  CUDA_TRY(cudaFreeAsync(tmp, stream));
  return 0;
}

int dispatch_first(int label_bytes, const void* labels, float* f, int64_t nlines, int64_t sx, float w,
                   int border, int flags, DeviceCache& dc, cudaStream_t s, uint8_t* codes = nullptr,
                   int64_t sy = 1, bool* codes_done = nullptr) {
  switch (label_bytes) {
    case 1: return launch_first<1>(labels, f, nlines, sx, w, border, flags, dc, s, codes, sy, codes_done);
    case 2: return launch_first<2>(labels, f, nlines, sx, w, border, flags, dc, s, codes, sy, codes_done);
    case 4: return launch_first<4>(labels, f, nlines, sx, w, border, flags, dc, s, codes, sy, codes_done);
    default: return launch_first<8>(labels, f, nlines, sx, w, border, flags, dc, s, codes, sy, codes_done);
  }
}

// Will launch_later() take the shared-memory tile kernel for this geometry?  (same conditions)
bool tile_path_ok(const edtb200::LineGeom& g, const DeviceCache& dc) {
  if (!((int64_t)g.n * g.line_stride + 64 < (1LL << 32) && g.n <= 4096 && g.inner_count < (1LL << 31))) return false;
  const int nb = (g.n + 255) / 256;
  int br = (g.n + nb - 1) / nb;
  if (nb > 1) br = (br + 3) & ~3;
  return tile_smem_bytes(g.n, 8, br * nb) <= (size_t)dc.max_smem_optin;
}

int dispatch_later(int label_bytes, const void* labels, float* f, const edtb200::LineGeom& g, float w,
                   int lo, int hi, int flags, const DeviceCache& dc, cudaStream_t s, int code_bit = 0,
                   bool pdl = false) {
  switch (label_bytes) {
    case 1: return launch_later<1>(labels, f, g, w, lo, hi, flags, dc, s, code_bit, pdl);
    case 2: return launch_later<2>(labels, f, g, w, lo, hi, flags, dc, s, code_bit, pdl);
    case 4: return launch_later<4>(labels, f, g, w, lo, hi, flags, dc, s, code_bit, pdl);
    default: return launch_later<8>(labels, f, g, w, lo, hi, flags, dc, s, code_bit, pdl);
  }
}

edtb200::LineGeom geom_for_axis(int axis, int64_t sx, int64_t sy, int64_t sz) {
  edtb200::LineGeom g;
  if (axis == 1) {
    g.outer_count = sz; g.outer_stride = sx * sy; g.inner_count = sx; g.line_stride = sx; g.n = (int)sy;
  } else {
    g.outer_count = 1; g.outer_stride = 0; g.inner_count = sx * sy; g.line_stride = sx * sy; g.n = (int)sz;
  }
  g.tiles_per_outer = 0;
  return g;
}

// Optional per-pass timing for bench.py: with profiling on, every transform of this thread records
// four CUDA events (before / between / after its axis passes) into a ring, so a timed loop can be
// analysed afterwards without any synchronisation inside it (edtb200_profile_passes /
// edtb200_pass_ms).
constexpr int kProfileRing = 256;
thread_local bool g_profile = false;
thread_local cudaEvent_t g_pass_events[kProfileRing][4];
thread_local bool g_pass_events_init = false;
thread_local int g_pass_marks[kProfileRing];       // highest mark index recorded in the slot
thread_local long g_pass_seq = 0;                  // transforms profiled so far

void mark_pass(int idx, cudaStream_t stream) {
  if (!g_profile) return;
  if (!g_pass_events_init) {
    for (auto& slot : g_pass_events) for (auto& e : slot) e = nullptr;
    g_pass_events_init = true;
  }
  if (idx == 0) ++g_pass_seq;
  const int slot = (int)((g_pass_seq - 1) % kProfileRing);
  cudaEvent_t& e = g_pass_events[slot][idx];
  if (!e && cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
  cudaEventRecord(e, stream);
  g_pass_marks[slot] = idx;
}

// All passes of one transform on device-resident buffers.
int run_passes(const void* labels, int label_bytes, int ndim, int64_t sx, int64_t sy, int64_t sz,
               float wx, float wy, float wz, int border, int flags, float* f,
               DeviceCache& dc, cudaStream_t stream) {
  using namespace edtb200;
  // sqrt / sign are applied by whichever pass is the last one; background-as-label (sdf)
  // changes the first pass only -- later passes treat every run alike.
  const int epilogue = ((flags & EDTB200_SQRT) ? kSqrt : 0) | ((flags & EDTB200_SIGNED) ? kNegate : 0);
  const int zero_label = (flags & EDTB200_SIGNED) ? kZeroLabel : 0;
  // EXPERIMENT, off by default (EDTB200_USE_CODES=1 turns it on): read labels wider than one byte
  // ONCE -- the first-axis pass leaves a one-byte code per voxel (differs from its y / z
  // neighbour, is background) and the later passes read that instead (3L+20 -> L+22 bytes per
  // voxel of HBM traffic).  Measured on B200, 512^3 uint32: Y 0.261 -> 0.244 ms, Z 0.280 -> 0.250 ms,
  // but the first-axis pass 0.180 -> 0.350 ms (two more label rows through L2 + 75 registers), a
  // net loss (0.84 vs 0.73 ms), so the passes read the labels by default.
  // The later passes are launched with programmatic stream serialization: the kernel before them
  // in the stream is our own previous pass, which never writes the labels, so their label staging
  // (before griddepcontrol.wait) may overlap its tail.  The per-axis entry points do not do this:
  // there the previous kernel is the caller's and may be the one producing the labels.
  int rc = 0;
  uint8_t* codes = nullptr;
  bool codes_done = false;
  const edtb200::LineGeom gy = geom_for_axis(1, sx, sy, sz), gz = geom_for_axis(2, sx, sy, sz);
  static const bool codes_enabled = getenv("EDTB200_USE_CODES") != nullptr;
  if (codes_enabled && label_bytes > 1 && ndim >= 2 && tile_path_ok(gy, dc) && (ndim < 3 || tile_path_ok(gz, dc)))
    CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&codes), (size_t)(sx * sy * sz), stream));
  mark_pass(0, stream);
  rc = dispatch_first(label_bytes, labels, f, sy * sz, sx, wx, border, zero_label | (ndim == 1 ? epilogue : 0), dc,
                      stream, codes, sy, &codes_done);
  mark_pass(1, stream);
  if (!rc && ndim >= 2) {
    rc = codes_done ? dispatch_later(1, codes, f, gy, wy, border, border, ndim == 2 ? epilogue : 0, dc, stream, 1)
                    : dispatch_later(label_bytes, labels, f, gy, wy, border, border, ndim == 2 ? epilogue : 0, dc,
                                     stream, 0, /*pdl=*/true);
    mark_pass(2, stream);
  }
  if (!rc && ndim >= 3) {
    rc = codes_done ? dispatch_later(1, codes, f, gz, wz, border, border, epilogue, dc, stream, 2)
                    : dispatch_later(label_bytes, labels, f, gz, wz, border, border, epilogue, dc, stream, 0,
                                     /*pdl=*/true);
    mark_pass(3, stream);
  }
  if (codes) cudaFreeAsync(codes, stream);
  return rc;
}


}  // namespace

extern "C" {

int edtb200_version(void) { return EDTB200_VERSION; }

const char* edtb200_last_error(void) { return g_error; }

int edtb200_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

int edtb200_transform(const void* labels, int label_bytes, int ndim, int64_t sx, int64_t sy, int64_t sz,
                      float wx, float wy, float wz, int black_border, int flags, float* out,
                      int device, void* stream_v) {
  int rc = check_dims(label_bytes, ndim, sx, sy, sz);
  if (rc) return rc;
  const int64_t total = sx * sy * sz;
  if (total == 0) return 0;
  if (!labels || !out) return fail(EDTB200_EINVAL, "null pointer");

  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dc = nullptr;
  rc = probe(device, &dc);
  if (rc) return rc;

  const bool lab_dev = flags & EDTB200_LABELS_ON_DEVICE;
  const bool out_dev = flags & EDTB200_OUT_ON_DEVICE;
  const int border = black_border != 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

  if (lab_dev && out_dev)
    return run_passes(labels, label_bytes, ndim, sx, sy, sz, wx, wy, wz, border, flags, out, *dc, stream);

  // host memory involved: stage through cached device buffers, synchronous on return
  if (!stream) {
    if (!dc->stream) CUDA_TRY(cudaStreamCreateWithFlags(&dc->stream, cudaStreamNonBlocking));
    stream = dc->stream;
  }
  const size_t lab_bytes = (size_t)total * (size_t)label_bytes;
  const size_t out_bytes = (size_t)total * sizeof(float);
  const void* d_labels = labels;
  float* d_out = out;
  if (!lab_dev) {
    if (dc->labels_bytes < lab_bytes) {
      if (dc->labels) cudaFree(dc->labels);
      dc->labels = nullptr; dc->labels_bytes = 0;
      CUDA_TRY(cudaMalloc(&dc->labels, lab_bytes));
      dc->labels_bytes = lab_bytes;
    }
    rc = upload(dc->labels, labels, lab_bytes, device, stream);
    if (rc) return rc;
    d_labels = dc->labels;
  }
  if (!out_dev) {
    if (dc->dist_bytes < out_bytes) {
      if (dc->dist) cudaFree(dc->dist);
      dc->dist = nullptr; dc->dist_bytes = 0;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&dc->dist), out_bytes));
      dc->dist_bytes = out_bytes;
    }
    d_out = dc->dist;
  }
  rc = run_passes(d_labels, label_bytes, ndim, sx, sy, sz, wx, wy, wz, border, flags, d_out, *dc, stream);
  if (rc) return rc;
  if (!out_dev) {
    rc = download(out, d_out, out_bytes, device, stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(stream));
  return 0;
}

int edtb200_transform_batch(const void* const* labels, float* const* outs, int count, int label_bytes, int ndim,
                            int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                            int black_border, int flags, int device) {
  if (count < 0) return fail(EDTB200_EINVAL, "negative count");
  if (flags & (EDTB200_LABELS_ON_DEVICE | EDTB200_OUT_ON_DEVICE))
    return fail(EDTB200_EINVAL, "edtb200_transform_batch pipelines HOST buffers; device-resident volumes are "
                                "already asynchronous through edtb200_transform");
  int rc = check_dims(label_bytes, ndim, sx, sy, sz);
  if (rc) return rc;
  const int64_t total = sx * sy * sz;
  if (total == 0 || count == 0) return 0;
  if (!labels || !outs) return fail(EDTB200_EINVAL, "null pointer");
  for (int k = 0; k < count; ++k)
    if (!labels[k] || !outs[k]) return fail(EDTB200_EINVAL, "null pointer in volume %d", k);

  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dcp = nullptr;
  rc = probe(device, &dcp);
  if (rc) return rc;
  DeviceCache& dc = *dcp;
  const size_t lab_bytes = (size_t)total * (size_t)label_bytes, out_bytes = (size_t)total * sizeof(float);
  auto grow = [](void** p, size_t* have, size_t need) -> cudaError_t {
    if (*have >= need) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *have = need;
    return e;
  };
  CUDA_TRY(grow(&dc.labels, &dc.labels_bytes, lab_bytes));
  CUDA_TRY(grow(reinterpret_cast<void**>(&dc.dist), &dc.dist_bytes, out_bytes));
  if (count > 1) {
    CUDA_TRY(grow(&dc.labels2, &dc.labels2_bytes, lab_bytes));
    CUDA_TRY(grow(reinterpret_cast<void**>(&dc.dist2), &dc.dist2_bytes, out_bytes));
  }
  if (!dc.stream) CUDA_TRY(cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking));
  if (!dc.stream_up) CUDA_TRY(cudaStreamCreateWithFlags(&dc.stream_up, cudaStreamNonBlocking));
  if (!dc.stream_down) CUDA_TRY(cudaStreamCreateWithFlags(&dc.stream_down, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    if (!dc.ev_up[i]) CUDA_TRY(cudaEventCreateWithFlags(&dc.ev_up[i], cudaEventDisableTiming));
    if (!dc.ev_comp[i]) CUDA_TRY(cudaEventCreateWithFlags(&dc.ev_comp[i], cudaEventDisableTiming));
    if (!dc.ev_down[i]) CUDA_TRY(cudaEventCreateWithFlags(&dc.ev_down[i], cudaEventDisableTiming));
  }
  void* d_labels[2] = {dc.labels, dc.labels2};
  float* d_dist[2] = {dc.dist, dc.dist2};
  const int border = black_border != 0;

  // Volume k uses slot k % 2.  Three streams: uploads (fed by a helper thread, so that pageable
  // buffers can be staged in both directions at once), the passes, downloads (this thread).
  //   upload k    waits for the passes of k-2 (they read the slot's labels)
  //   passes k    wait for upload k and for download k-2 (it reads the slot's distances)
  //   download k  waits for the passes of k
  // Host-side counters make sure an event has been RECORDED before someone waits on it.
  std::mutex m;
  std::condition_variable cv;
  int uploads_queued = 0, passes_queued = 0, failed = 0;
  char upload_error[sizeof(g_error)] = "";
  std::thread uploader([&] {
    cudaSetDevice(device);
    for (int k = 0; k < count; ++k) {
      const int slot = k & 1;
      int urc = 0;
      if (k >= 2) {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&] { return passes_queued >= k - 1 || failed; });
        if (failed) return;
        if (cudaStreamWaitEvent(dc.stream_up, dc.ev_comp[slot], 0) != cudaSuccess) urc = EDTB200_ECUDA;
      }
      if (!urc) urc = upload(d_labels[slot], labels[k], lab_bytes, device, dc.stream_up);
      if (!urc && cudaEventRecord(dc.ev_up[slot], dc.stream_up) != cudaSuccess) urc = EDTB200_ECUDA;
      std::lock_guard<std::mutex> l(m);
      if (urc) {
        failed = urc;
        snprintf(upload_error, sizeof(upload_error), "upload of volume %d failed: %.400s", k,
                 urc == EDTB200_ECUDA ? cudaGetErrorString(cudaGetLastError()) : g_error);
      } else {
        uploads_queued = k + 1;
      }
      cv.notify_all();
      if (urc) return;
    }
  });
  auto bail = [&](int code) {
    { std::lock_guard<std::mutex> l(m); if (!failed) failed = code; }
    cv.notify_all();
    uploader.join();
    cudaStreamSynchronize(dc.stream_up); cudaStreamSynchronize(dc.stream); cudaStreamSynchronize(dc.stream_down);
    return code;
  };
  #define BATCH_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
      fail(EDTB200_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); return bail(EDTB200_ECUDA); } } while (0)
  for (int k = 0; k < count; ++k) {
    const int slot = k & 1;
    {
      std::unique_lock<std::mutex> l(m);
      cv.wait(l, [&] { return uploads_queued >= k + 1 || failed; });
      if (failed) { l.unlock(); uploader.join(); cudaDeviceSynchronize(); return fail(failed, "%s", upload_error); }
    }
    BATCH_TRY(cudaStreamWaitEvent(dc.stream, dc.ev_up[slot], 0));
    if (k >= 2) BATCH_TRY(cudaStreamWaitEvent(dc.stream, dc.ev_down[slot], 0));
    rc = run_passes(d_labels[slot], label_bytes, ndim, sx, sy, sz, wx, wy, wz, border, flags, d_dist[slot], dc,
                    dc.stream);
    if (rc) return bail(rc);
    BATCH_TRY(cudaEventRecord(dc.ev_comp[slot], dc.stream));
    { std::lock_guard<std::mutex> l(m); passes_queued = k + 1; }
    cv.notify_all();
    BATCH_TRY(cudaStreamWaitEvent(dc.stream_down, dc.ev_comp[slot], 0));
    rc = download(outs[k], d_dist[slot], out_bytes, device, dc.stream_down);
    if (rc) return bail(rc);
    BATCH_TRY(cudaEventRecord(dc.ev_down[slot], dc.stream_down));
  }
  #undef BATCH_TRY
  uploader.join();
  CUDA_TRY(cudaStreamSynchronize(dc.stream_down));
  return 0;
}

int edtb200_transform_voxel_graph(const void* labels, int label_bytes, const unsigned char* graph, int ndim,
                                  int64_t sx, int64_t sy, int64_t sz, float wx, float wy, float wz,
                                  int black_border, int flags, float* out, int device, void* stream_v) {
  using namespace edtb200;
  if (ndim != 2 && ndim != 3)
    return fail(EDTB200_EINVAL, "a voxel graph needs a 2-D or 3-D volume (got ndim %d)", ndim);
  if (flags & EDTB200_SIGNED)
    return fail(EDTB200_EINVAL, "EDTB200_SIGNED is not defined with a voxel graph: subtract two transforms");
  if ((flags & EDTB200_LABELS_FLOAT) && label_bytes != 4 && label_bytes != 8)
    return fail(EDTB200_EINVAL, "EDTB200_LABELS_FLOAT needs 4- or 8-byte labels");
  int rc = check_dims(label_bytes, ndim, sx, sy, sz);
  if (rc) return rc;
  const int64_t total = sx * sy * sz;
  if (total == 0) return 0;
  if (!labels || !graph || !out) return fail(EDTB200_EINVAL, "null pointer");
  int64_t sx2 = 2 * sx, sy2 = 2 * sy, sz2 = ndim == 3 ? 2 * sz : 1;
  rc = check_dims(1, ndim, sx2, sy2, sz2);
  if (rc) return rc;
  const int64_t total2 = sx2 * sy2 * sz2;

  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dc = nullptr;
  rc = probe(device, &dc);
  if (rc) return rc;
  const bool in_dev = flags & EDTB200_LABELS_ON_DEVICE;
  const bool out_dev = flags & EDTB200_OUT_ON_DEVICE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!stream && !(in_dev && out_dev)) {
    if (!dc->stream) CUDA_TRY(cudaStreamCreateWithFlags(&dc->stream, cudaStreamNonBlocking));
    stream = dc->stream;
  }

  // stream-ordered scratch: the inputs when they come from the host, the doubled byte mask, the
  // doubled distance volume, and the result when it goes back to the host
  void *d_labels = nullptr, *d_graph = nullptr, *d_cells = nullptr, *d_doubled = nullptr, *d_result = nullptr;
  auto release = [&]() {
    void* all[] = {d_labels, d_graph, d_cells, d_doubled, d_result};
    for (void* p : all) if (p) cudaFreeAsync(p, stream);
  };
  #define VG_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { release(); \
      return fail(e_ == cudaErrorMemoryAllocation ? EDTB200_ENOMEM : EDTB200_ECUDA, "%s: %s", #expr, \
                  cudaGetErrorString(e_)); } } while (0)
  const void* lab = labels;
  const uint8_t* gr = graph;
  if (!in_dev) {
    VG_TRY(cudaMallocAsync(&d_labels, (size_t)total * label_bytes, stream));
    VG_TRY(cudaMallocAsync(&d_graph, (size_t)total, stream));
    rc = upload(d_labels, labels, (size_t)total * label_bytes, device, stream);
    if (!rc) rc = upload(d_graph, graph, (size_t)total, device, stream);
    if (rc) { release(); return rc; }
    lab = d_labels;
    gr = static_cast<const uint8_t*>(d_graph);
  }
  VG_TRY(cudaMallocAsync(&d_cells, (size_t)total2, stream));
  VG_TRY(cudaMallocAsync(&d_doubled, (size_t)total2 * sizeof(float), stream));
  float* result = out;
  if (!out_dev) {
    VG_TRY(cudaMallocAsync(&d_result, (size_t)total * sizeof(float), stream));
    result = static_cast<float*>(d_result);
  }

  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, (int64_t)dc->sm_count * 32);
  const int border = black_border != 0, as_float = (flags & EDTB200_LABELS_FLOAT) ? 1 : 0;
  uint8_t* cells = static_cast<uint8_t*>(d_cells);
  switch (label_bytes) {
    case 1: voxel_graph_expand_kernel<1><<<blocks, threads, 0, stream>>>(lab, gr, cells, sx, sy, sz, ndim, border, as_float); break;
    case 2: voxel_graph_expand_kernel<2><<<blocks, threads, 0, stream>>>(lab, gr, cells, sx, sy, sz, ndim, border, as_float); break;
    case 4: voxel_graph_expand_kernel<4><<<blocks, threads, 0, stream>>>(lab, gr, cells, sx, sy, sz, ndim, border, as_float); break;
    default: voxel_graph_expand_kernel<8><<<blocks, threads, 0, stream>>>(lab, gr, cells, sx, sy, sz, ndim, border, as_float); break;
  }
  VG_TRY(cudaGetLastError());
  // half the anisotropy (vg:102-107, 199-204); the sqrt is taken by the gather instead
  rc = run_passes(cells, 1, ndim, sx2, sy2, sz2, wx / 2, wy / 2, wz / 2, border, 0,
                  static_cast<float*>(d_doubled), *dc, stream);
  if (rc) { release(); return rc; }
  voxel_graph_gather_kernel<<<blocks, threads, 0, stream>>>(static_cast<const float*>(d_doubled), result, sx, sy, sz,
                                                            (flags & EDTB200_SQRT) ? 1 : 0);
  VG_TRY(cudaGetLastError());
  if (!out_dev) {
    rc = download(out, result, (size_t)total * sizeof(float), device, stream);
    if (rc) { release(); return rc; }
  }
  release();
  if (!(in_dev && out_dev)) VG_TRY(cudaStreamSynchronize(stream));
  #undef VG_TRY
  return 0;
}

int edtb200_pass_first(const void* labels_dev, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
                       float wx, int black_border, int flags, float* f_dev, int device, void* stream) {
  int rc = check_dims(label_bytes, 3, sx, sy, sz);
  if (rc) return rc;
  if (sx * sy * sz == 0) return 0;
  if (!labels_dev || !f_dev) return fail(EDTB200_EINVAL, "null pointer");
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dc = nullptr;
  rc = probe(device, &dc);
  if (rc) return rc;
  const int kflags = ((flags & EDTB200_SQRT) ? edtb200::kSqrt : 0) |
                     ((flags & EDTB200_SIGNED) ? edtb200::kZeroLabel : 0);
  return dispatch_first(label_bytes, labels_dev, f_dev, sy * sz, sx, wx, black_border != 0, kflags, *dc,
                        static_cast<cudaStream_t>(stream));
}

int edtb200_pass_later(const void* labels_dev, int label_bytes, int axis, int64_t sx, int64_t sy, int64_t sz,
                       float w, int border_lo, int border_hi, int flags, float* f_dev, int device,
                       void* stream) {
  int rc = check_dims(label_bytes, 3, sx, sy, sz);
  if (rc) return rc;
  if (axis != 1 && axis != 2) return fail(EDTB200_EINVAL, "axis must be 1 (Y) or 2 (Z)");
  if (sx * sy * sz == 0) return 0;
  if (!labels_dev || !f_dev) return fail(EDTB200_EINVAL, "null pointer");
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dc = nullptr;
  rc = probe(device, &dc);
  if (rc) return rc;
  return dispatch_later(label_bytes, labels_dev, f_dev, geom_for_axis(axis, sx, sy, sz), w,
                        border_lo != 0, border_hi != 0,
                        ((flags & EDTB200_SQRT) ? edtb200::kSqrt : 0) |
                            ((flags & EDTB200_SIGNED) ? edtb200::kNegate : 0),
                        *dc, static_cast<cudaStream_t>(stream));
}

int edtb200_slab_face_runs(const void* labels_dev, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
                           int high_face, int halo, int flags, unsigned char* m_dev, int* overflow_dev,
                           int device, void* stream_v) {
  int rc = check_dims(label_bytes, 3, sx, sy, sz);
  if (rc) return rc;
  if (halo < 1 || halo > 254) return fail(EDTB200_EINVAL, "halo must be in 1..254");
  if (sx * sy * sz == 0) return 0;
  if (!labels_dev || !m_dev || !overflow_dev) return fail(EDTB200_EINVAL, "null pointer");
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dc = nullptr;
  rc = probe(device, &dc);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int64_t plane = sx * sy;
  const unsigned blocks = (unsigned)((plane + 255) / 256);
  const int zl = (flags & EDTB200_SIGNED) ? 1 : 0;
  using namespace edtb200;
  switch (label_bytes) {
    case 1: face_runs_kernel<1><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(labels_dev), plane, (int)sz, high_face, halo, zl, m_dev, overflow_dev); break;
    case 2: face_runs_kernel<2><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(labels_dev), plane, (int)sz, high_face, halo, zl, m_dev, overflow_dev); break;
    case 4: face_runs_kernel<4><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t*>(labels_dev), plane, (int)sz, high_face, halo, zl, m_dev, overflow_dev); break;
    default: face_runs_kernel<8><<<blocks, 256, 0, stream>>>(static_cast<const uint64_t*>(labels_dev), plane, (int)sz, high_face, halo, zl, m_dev, overflow_dev); break;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int edtb200_slab_face_fixup(const void* labels_dev, int label_bytes, int64_t sx, int64_t sy, int64_t sz,
                            int high_face, int halo, float wz, int flags, const void* nb_label_dev,
                            const unsigned char* nb_m_dev, const float* nb_f_dev, float* f_dev, int* inexact_dev,
                            int device, void* stream_v) {
  int rc = check_dims(label_bytes, 3, sx, sy, sz);
  if (rc) return rc;
  if (halo < 1 || halo > 254) return fail(EDTB200_EINVAL, "halo must be in 1..254");
  if (sx * sy * sz == 0) return 0;
  if (!labels_dev || !nb_label_dev || !nb_m_dev || !nb_f_dev || !f_dev) return fail(EDTB200_EINVAL, "null pointer");
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  DeviceCache* dc = nullptr;
  rc = probe(device, &dc);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int64_t plane = sx * sy;
  const unsigned blocks = (unsigned)((plane + 255) / 256);
  using namespace edtb200;
  const int kflags = ((flags & EDTB200_SQRT) ? kSqrt : 0) | ((flags & EDTB200_SIGNED) ? (kNegate | kZeroLabel) : 0);
  const float w2 = wz * wz;
  switch (label_bytes) {
    case 1: face_fixup_kernel<1><<<blocks, 256, 0, stream>>>(static_cast<const uint8_t*>(labels_dev), f_dev, plane, (int)sz, high_face, halo, w2, static_cast<const uint8_t*>(nb_label_dev), nb_m_dev, nb_f_dev, kflags, inexact_dev); break;
    case 2: face_fixup_kernel<2><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(labels_dev), f_dev, plane, (int)sz, high_face, halo, w2, static_cast<const uint16_t*>(nb_label_dev), nb_m_dev, nb_f_dev, kflags, inexact_dev); break;
    case 4: face_fixup_kernel<4><<<blocks, 256, 0, stream>>>(static_cast<const uint32_t*>(labels_dev), f_dev, plane, (int)sz, high_face, halo, w2, static_cast<const uint32_t*>(nb_label_dev), nb_m_dev, nb_f_dev, kflags, inexact_dev); break;
    default: face_fixup_kernel<8><<<blocks, 256, 0, stream>>>(static_cast<const uint64_t*>(labels_dev), f_dev, plane, (int)sz, high_face, halo, w2, static_cast<const uint64_t*>(nb_label_dev), nb_m_dev, nb_f_dev, kflags, inexact_dev); break;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int edtb200_profile_passes(int enable) {
  g_profile = enable != 0;
  if (enable) g_pass_seq = 0;
  return 0;
}

int edtb200_pass_ms(int steps_back, float* ms3) {
  if (!ms3) return fail(EDTB200_EINVAL, "null pointer");
  ms3[0] = ms3[1] = ms3[2] = 0.0f;
  if (steps_back < 0 || steps_back >= kProfileRing || steps_back >= g_pass_seq)
    return fail(EDTB200_EINVAL, "no profiled transform %d steps back", steps_back);
  const int slot = (int)((g_pass_seq - 1 - steps_back) % kProfileRing);
  for (int i = 0; i < g_pass_marks[slot] && i < 3; ++i) {
    if (!g_pass_events[slot][i] || !g_pass_events[slot][i + 1]) break;
    CUDA_TRY(cudaEventSynchronize(g_pass_events[slot][i + 1]));
    CUDA_TRY(cudaEventElapsedTime(&ms3[i], g_pass_events[slot][i], g_pass_events[slot][i + 1]));
  }
  return 0;
}

int edtb200_release(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceGuard restore_device;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  for (int d = 0; d < count && d < kMaxDevices; ++d) {
    DeviceCache& dc = g_cache[d];
    bool any = dc.labels || dc.dist || dc.stream || dc.labels2 || dc.dist2 || dc.stream_up;
    for (auto& t : dc.tables) any = any || t.data;
    if (!any) continue;
    cudaSetDevice(d);
    cudaDeviceSynchronize();
    for (auto& t : dc.tables) { if (t.data) cudaFree(t.data); if (t.ready) cudaEventDestroy(t.ready); }
    if (dc.stream) { cudaStreamSynchronize(dc.stream); cudaStreamDestroy(dc.stream); }
    if (dc.labels) cudaFree(dc.labels);
    if (dc.dist) cudaFree(dc.dist);
    if (dc.labels2) cudaFree(dc.labels2);
    if (dc.dist2) cudaFree(dc.dist2);
    if (dc.stream_up) cudaStreamDestroy(dc.stream_up);
    if (dc.stream_down) cudaStreamDestroy(dc.stream_down);
    for (int i = 0; i < 2; ++i) {
      if (dc.ev_up[i]) cudaEventDestroy(dc.ev_up[i]);
      if (dc.ev_comp[i]) cudaEventDestroy(dc.ev_comp[i]);
      if (dc.ev_down[i]) cudaEventDestroy(dc.ev_down[i]);
    }
    dc = DeviceCache();
  }
  for (int dir = 0; dir < 2; ++dir)
    for (int d = 0; d < count && d < kMaxDevices; ++d) {
      StageBuffers& sb = g_stage[dir][d];
      for (int i = 0; i < kStages; ++i) {
        if (sb.buf[i]) { cudaFreeHost(sb.buf[i]); sb.buf[i] = nullptr; }
        if (sb.ev[i]) { cudaEventDestroy(sb.ev[i]); sb.ev[i] = nullptr; }
      }
    }
  return 0;
}

}  // extern "C"
