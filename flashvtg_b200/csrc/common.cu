#include "common.cuh"

#include <stdlib.h>
#include <vector>

namespace fvtg {

HostState& host_state() {
  static thread_local HostState s = {{0}, 0};
  return s;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(host_state().err, sizeof(host_state().err), fmt, ap);
  va_end(ap);
  return code;
}

struct ProfRec { cudaEvent_t a, b; int cls; };
struct ProfState {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
};
static ProfState& prof_state() {
  static thread_local ProfState s;
  return s;
}
#ifdef FVTG_DEBUG_HOOKS
static thread_local long long* g_trace = nullptr;
long long* dbg_trace() { return g_trace; }
void set_dbg_trace(long long* p) { g_trace = p; }
#endif
bool prof_on() { return prof_state().on; }
static cudaEvent_t prof_event() {
  ProfState& p = prof_state();
  if (!p.pool.empty()) {
    cudaEvent_t e = p.pool.back();
    p.pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(cudaStream_t st, int cls) {
  ProfRec r;
  r.a = prof_event();
  r.b = prof_event();
  r.cls = cls;
  cudaEventRecord(r.a, st);
  prof_state().recs.push_back(r);
}
void prof_end(cudaStream_t st) {
  ProfState& p = prof_state();
  if (!p.recs.empty()) cudaEventRecord(p.recs.back().b, st);
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("FVTG_PDL"); return !e || atoi(e) != 0; }();
  return on;
}

int check_arch() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(FVTG_EARCH, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return fail(FVTG_EARCH, "flashvtg_b200 needs an sm_100 device, found sm_%d%d", major, minor);
  return FVTG_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t rows, cols, pitch;
  uint32_t box_rows, box_cols;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && pitch == o.pitch &&
           box_rows == o.box_rows && box_cols == o.box_cols;
  }
};
struct TmapEntry { TmapKey key; CUtensorMap map; };

// Descriptors are pure functions of (pointer, geometry): the same few hundred recur every step
// (workspace carving is deterministic), so encoding is cached per thread.
static std::vector<TmapEntry>& tmap_cache() {
  static thread_local std::vector<TmapEntry> c;
  return c;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                   uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  const TmapKey key{base, rows, cols, pitch_elems, box_rows, box_cols};
  std::vector<TmapEntry>& cache = tmap_cache();
  for (const TmapEntry& e : cache)
    if (e.key == key) {
      *map = e.map;
      return FVTG_OK;
    }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(FVTG_ELAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((pitch_elems * 2) & 15))
    return fail(FVTG_EINVAL, "TMA operand must be 16-byte aligned with a 16-byte pitch");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(FVTG_ELAUNCH, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu pitch=%llu",
                (int)r, (unsigned long long)rows, (unsigned long long)cols,
                (unsigned long long)pitch_elems);
  if (cache.size() >= 1024) cache.clear();
  cache.push_back(TmapEntry{key, *map});
  return FVTG_OK;
}

}  // namespace fvtg

extern "C" {

const char* fvtg_last_error(void) { return fvtg::host_state().err; }
int64_t fvtg_last_launch_count(void) { return fvtg::host_state().launches; }
int32_t fvtg_abi_version(void) { return FVTG_ABI_VERSION; }

#ifdef FVTG_DEBUG_HOOKS
void fvtg_dbg_set_trace(void* device_buf) { fvtg::set_dbg_trace(static_cast<long long*>(device_buf)); }
#endif

void fvtg_prof_enable(int32_t on) { fvtg::prof_state().on = on != 0; }

int32_t fvtg_prof_collect(double* ms, int64_t* launches, int32_t n_classes) {
  using namespace fvtg;
  ProfState& p = prof_state();
  for (int i = 0; i < n_classes; ++i) {
    if (ms) ms[i] = 0.0;
    if (launches) launches[i] = 0;
  }
  for (ProfRec& r : p.recs) {
    cudaError_t e = cudaEventSynchronize(r.b);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) return fail(FVTG_ELAUNCH, "prof_collect: %s", cudaGetErrorString(e));
    if (r.cls >= 0 && r.cls < n_classes) {
      if (ms) ms[r.cls] += t;
      if (launches) launches[r.cls] += 1;
    }
    p.pool.push_back(r.a);
    p.pool.push_back(r.b);
  }
  p.recs.clear();
  return FVTG_OK;
}
}
