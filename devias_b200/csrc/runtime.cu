// Library runtime: error reporting, launch counter, TMA tensor-map encoding through the driver entry point
// (resolved with cudaGetDriverEntryPoint so the .so has no link-time dependency on libcuda).
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace dv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* what, const char* detail, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s:%d)", what, detail, file, line);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DEVIAS_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_nd(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, const void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled", "driver entry point unavailable (no CUDA driver?)", __FILE__, __LINE__);
    return DEVIAS_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof(msg), "CUresult %d (rank %u dims %llu,%llu stride %llu box %u,%u)", (int)r, rank,
             (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
    set_last_error("cuTensorMapEncodeTiled", msg, __FILE__, __LINE__);
    return DEVIAS_ERR_CUDA;
  }
  return DEVIAS_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t str[1] = {row_stride_bytes};
  const uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_nd(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// ---------------------------------------------------------------- optional per-kernel timing (bench.py roofline leg)
struct ProfRec { cudaEvent_t a, b; int kind; double work; };
static std::vector<ProfRec> g_prof;
static size_t g_prof_used = 0;
static bool g_prof_on = false;
static bool g_prof_capture_only = false;

// Inside a stream capture the record becomes an EXTERNAL event-record node: every replay of the captured graph re-records the
// event, and cudaEventElapsedTime reads the times of the last replay -- the kernels are timed inside the replayed step, under its
// clocks, caches and back-to-back launch conditions.
static void record(cudaEvent_t e, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive)
    cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);
  else
    cudaEventRecord(e, s);
}

int prof_begin(int kind, double work, cudaStream_t s) {
  if (!g_prof_on) return -1;
  if (g_prof_capture_only) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusActive) return -1;
  }
  if (g_prof_used == g_prof.size()) {
    ProfRec r{};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    g_prof.push_back(r);
  }
  ProfRec& r = g_prof[g_prof_used];
  r.kind = kind;
  r.work = work;
  record(r.a, s);
  return (int)g_prof_used++;
}
void prof_end(int id, cudaStream_t s) {
  if (id >= 0) record(g_prof[id].b, s);
}

}  // namespace dv

extern "C" int devias_profile_begin(void) {
  dv::g_prof_used = 0;
  dv::g_prof_on = true;
  dv::g_prof_capture_only = false;
  return DEVIAS_OK;
}
extern "C" int devias_profile_begin_capture(void) {
  dv::g_prof_used = 0;
  dv::g_prof_on = true;
  dv::g_prof_capture_only = true;
  return DEVIAS_OK;
}
extern "C" int devias_profile_pause(void) {
  dv::g_prof_on = false;
  return DEVIAS_OK;
}
extern "C" int devias_profile_end(int kind, double* total_ms, double* total_work, int64_t* launches) {
  using namespace dv;
  g_prof_on = false;      // the records stay: a captured graph keeps re-recording them, so this may be called after every replay
  DV_CHECK_CUDA(cudaDeviceSynchronize());
  double ms = 0, work = 0;
  long long n = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    if (g_prof[i].kind != kind) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b) == cudaSuccess) {
      ms += t;
      work += g_prof[i].work;
      ++n;
    }
  }
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  if (launches) *launches = n;
  return DEVIAS_OK;
}

extern "C" int devias_abi_version(void) { return 1; }
extern "C" const char* devias_last_error(void) { return dv::g_err; }
extern "C" int64_t devias_launch_count(void) { return dv::g_launches.load(std::memory_order_relaxed); }
