// Host-side plumbing shared by every translation unit: error string, launch
// counter, the TMA tensor-map encoder (driver entry point fetched at run time,
// so the library links against cudart only) and the row-space geometry structs
// that host code and kernels agree on.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/flashvtg_b200.h"

namespace fvtg {

typedef __nv_bfloat16 bf16;

// -------------------------------------------------------------- host state --
struct HostState {
  char err[512];
  int64_t launches;
};
HostState& host_state();

int fail(int code, const char* fmt, ...);
inline void count_launch(int n = 1) { host_state().launches += n; }

#define FVTG_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return fvtg::fail(FVTG_ELAUNCH, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                            \
  } while (0)

#define FVTG_TRY(expr)             \
  do {                             \
    int _rc = (expr);              \
    if (_rc != FVTG_OK) return _rc; \
  } while (0)

#define FVTG_LAUNCH_CHECK(name)                                                        \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess)                                                             \
      return fvtg::fail(FVTG_ELAUNCH, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    fvtg::count_launch();                                                              \
  } while (0)

// ------------------------------------------------------------- profiler --
// Bench hook (fvtg_prof_enable / fvtg_prof_collect): when enabled every launcher brackets its
// kernel with a CUDA event pair on the launching stream, so bench.py can time each kernel class
// live (the roofline numerator) without a profiler attached.  Off by default: zero cost.
enum ProfClass { PC_GEMM = 0, PC_ATTN = 1, PC_LNCAST = 2, PC_DECODE = 3, PC_OTHER = 4, PC_LAYER = 5, PC_COUNT = 6 };
// Debug build only (libflashvtg_b200_dbg.so, -DFVTG_DEBUG_HOOKS): the device buffer registered with
// fvtg_dbg_set_trace (>= 4096 int64) that kernels stamp with clock64; the product library has no such hook
// and every kernel's trace pointer is a constant null.
#ifdef FVTG_DEBUG_HOOKS
void set_dbg_trace(long long* p);
long long* dbg_trace();
#else
inline long long* dbg_trace() { return nullptr; }
#endif
bool prof_on();
void prof_begin(cudaStream_t st, int cls);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(cudaStream_t s, int cls) : st(s), on(prof_on()) { if (on) prof_begin(st, cls); }
  ~ProfScope() { if (on) prof_end(st); }
};

// Kernel launch with programmatic stream serialisation: the kernel's prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlaps the tail of the previous kernel of the stream; the
// kernel itself calls pdl_wait() before it touches anything a predecessor wrote (ptx.cuh).
bool pdl_enabled();  // env FVTG_PDL (default on)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int check_arch();  // FVTG_OK iff the current device is sm_100
int sm_count();

// 2-D bf16 tensor map: global [rows][cols] with row pitch `pitch_elems`, box
// {box_cols (<=64 -> 128 B, SWIZZLE_128B), box_rows}.  Out-of-bounds elements
// (including negative row coordinates) are zero-filled.
int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                   uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols);

// ----------------------------------------------------- row-space geometry --
// Per-chunk geometry of the pyramid / head row spaces (DESIGN.md "row spaces").
//   chain space, step j : video b owns rows [b*(P0>>j), (b+1)*(P0>>j)), first (len_b>>j) valid
//   H1 (per-level gaps) : video b owns PH1 rows; level l at offset o1[l], then >= pad zero rows
//   H2 (concatenated)   : video b owns PH2 rows; `pad` zeros, then the N_b points, then zeros
struct PyrGeo {
  int nlev;     // levels present at the batch's padded Lv
  int pad;      // zero rows each head conv needs on both sides (max(head_k, coord_k) / 2)
  int P0;       // chain pitch at step 0: Lv rounded up to a multiple of 2^(nlev-1) (and 8)
  int PH1, PH2;
  int n_max;    // sum_l (Lv >> l)
  int o1[FVTG_MAX_LEVELS];
  const int* vlen;  // [Bc] true video lengths of this chunk
};

// fp32 residual streams (video stream Y, dummy/text stream X, the sine table) are kept in a
// tile-blocked layout [row/128][col/4][row%128][4]: the tcgen05 epilogues own one row per thread
// (TMEM lane == row), so consecutive lanes touch consecutive 16-byte groups - every residual load
// and store is fully coalesced without staging through shared memory.
__host__ __device__ inline size_t blk_off(size_t row, int col) {
  return ((row >> 7) * 64 + static_cast<size_t>(col >> 2)) * 512 + (row & 127) * 4 + (col & 3);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

}  // namespace fvtg
