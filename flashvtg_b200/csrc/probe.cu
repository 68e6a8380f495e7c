#ifdef FVTG_DEBUG_HOOKS   // micro-benchmarks: compiled into libflashvtg_b200_dbg.so only, never into the product library
// Debug probes (not on the product path), run by tools/probe_tma.py on a B200:
//  * fvtg_dbg_tma_probe: how fast can every SM stream the SAME weight matrix from L2 into a
//    shared-memory ring of `stages` units with no consumer work?  mode 0: 2-D tensor boxes
//    {64 cols, 128 rows} (16 KB, SWIZZLE_128B) ; mode 1: 1-D bulk copies of contiguous 16 KB
//    (weights pre-packed as shared-memory images) ; mode 2: 2-D boxes {64, 256} (32 KB) ;
//    mode 3: mode 0 issued from two producer threads.
//  * fvtg_dbg_mma_probe: cycles per tcgen05.mma (128 x N x 16, bf16, both operands in shared
//    memory) issued back to back on resident tiles: the tensor-pipe floor for N = 128 / 256.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes),
        "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
tma_stream_probe_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2,
                        const uint8_t* wraw, int mode, int stages, int units, int rows_total,
                        long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 13 * 16384);
  uint64_t* empty = full + 16;
  const int ub = (mode == 2) ? 32768 : 16384;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmW2);
  }
  __syncthreads();
  const long long t0 = clock64();
  const int nprod = (mode == 3) ? 2 : 1;
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nprod) {  // producer(s): warp 0 (and warp 1 in mode 3)
    for (int u = w; u < units; u += nprod) {
      const int s = u % stages;
      const uint32_t ph = (u / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], ub);
      if (mode == 1) {
        bulk_load_1d(smem + s * ub, wraw + static_cast<size_t>(u % (rows_total * 512 / 16384)) * 16384,
                     16384, &full[s]);
      } else if (mode == 2) {
        const int r0 = (u * 256) % rows_total;
        tma_load_2d(smem + s * ub, &tmW2, ((u / (rows_total / 256)) % 4) * 64, r0, &full[s]);
      } else {
        const int r0 = (u * 128) % rows_total;
        tma_load_2d(smem + s * ub, &tmW, ((u / (rows_total / 128)) % 4) * 64, r0, &full[s]);
      }
    }
  } else if (threadIdx.x == 64) {  // consumer: release immediately
    int s = 0; uint32_t ph = 0;
    for (int u = 0; u < units; ++u) {
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
}

// One thread issues `iters` MMAs (each a full K = 64 unit: 4 x tcgen05.mma 128 x N x 16) on tiles
// resident in shared memory, commits, waits.  out[0] = cycles.
__global__ void __launch_bounds__(128, 1)
mma_rate_probe_kernel(int N, int iters, int nbuf, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 12 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&holder, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t base = smem_u32(smem);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int b = i % nbuf;
      const uint64_t da = umma_desc_sw128(base + b * 16384);
      const uint64_t db = umma_desc_sw128(base + 4 * 16384 + (b % 2) * 32768);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + (i & 1) * 256, da + 2 * k, db + 2 * k, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    out_cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}


// ---- v3: weight streaming with an honest producer loop (no integer divisions), optional
// cluster multicast.  Every CTA of a cluster of `csize` loads rows [rank*128/csize, ...) of each
// 16 KB unit and multicasts the slice to all CTAs; nprod producer lanes (different warps) split
// the unit stream.  Consumer releases each stage to every CTA of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, int c0, int c1,
                                               uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(192, 1)
tma_stream_probe3_kernel(const __grid_constant__ CUtensorMap tmW, int csize, int nprod, int stages,
                         int passes, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 12 * 16384);
  uint64_t* empty = full + 16;
  const uint32_t rank = csize > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], nprod > 1 ? 1 : 1); mbar_init(&empty[s], csize); }
    fence_mbar_init();
    prefetch_tmap(&tmW);
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  const long long t0 = clock64();
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slice_rows = 128 / csize;
  const uint32_t slice_bytes = 16384 / csize;
  const uint16_t mask = static_cast<uint16_t>((1u << csize) - 1);
  // unit stream per pass: 18 row blocks x 4 k blocks of the [2304][256] matrix
  if (lane == 0 && w < nprod) {
    // producer w handles stages s with s % nprod == w (stages % nprod == 0 required)
    uint32_t ph = 0;
    int s = 0, sp = 0;  // stage, stage % nprod (kept as counters: no division in the timed loop)
    for (int p = 0; p < passes; ++p)
      for (int rb = 0; rb < 18; ++rb)
        for (int kb = 0; kb < 4; ++kb) {
          if (sp == w) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], 16384);
            if (csize > 1)
              tma_load_2d_mc(smem + s * 16384 + rank * slice_bytes, &tmW, kb * 64,
                             rb * 128 + rank * slice_rows, &full[s], mask);
            else
              tma_load_2d(smem + s * 16384, &tmW, kb * 64, rb * 128, &full[s]);
          }
          if (++sp == nprod) sp = 0;
          if (++s == stages) { s = 0; sp = 0; ph ^= 1; }
        }
  } else if (threadIdx.x == 160) {  // consumer: release immediately to every CTA of the cluster
    int s = 0; uint32_t ph = 0;
    const int units = passes * 72;
    for (int u = 0; u < units; ++u) {
      mbar_wait(&full[s], ph);
      if (csize > 1) {
        for (int c = 0; c < csize; ++c) mbar_arrive_remote(&empty[s], c);
      } else {
        mbar_arrive(&empty[s]);
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
}


// tcgen05.mma issue-rate floor for the four operand flavours the layer kernel uses:
// cg = 1 / 2 (CTA pair, cluster launch), ts = 0 (A in shared memory) / 1 (A in TMEM).
template <int cg>
__global__ void __launch_bounds__(128, 1)
mma_rate_probe2_kernel(int N, int iters, int ts, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 12 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  uint32_t rank = 0;
  if constexpr (cg == 2) rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) {
    if constexpr (cg == 2) { tmem_alloc_cg2(&holder, 512); tmem_relinquish_cg2(); }
    else { tmem_alloc(&holder, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if constexpr (cg == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(cg == 2 ? 256 : 128, N);
    const uint32_t base = smem_u32(smem);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t da = umma_desc_sw128(base + (i & 3) * 16384);
      const uint64_t db = umma_desc_sw128(base + 4 * 16384 + (i & 1) * 32768);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (cg == 2) {
          if (ts) umma_bf16_ts_cg2(tmem, tmem + 256 + 8 * k + 32 * (i & 3), db + 2 * k, idesc, 1u);
          else umma_bf16_cg2(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
        } else {
          if (ts) umma_bf16_ts(tmem, tmem + 256 + 8 * k + 32 * (i & 3), db + 2 * k, idesc, 1u);
          else umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, 1u);
        }
      }
    }
    if constexpr (cg == 2) umma_commit_cg2(&bar, 1); else umma_commit(&bar);
    mbar_wait(&bar, 0);
    out_cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (cg == 2) cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    if constexpr (cg == 2) tmem_dealloc_cg2(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}


// ---------------------------------------------------------------------------------------------
// fvtg_dbg_stream_probe: how fast can 148 CTAs stream an fp32 [rows][dim] matrix with the access
// shape of the first-projection kernel?  16 warps per CTA, warp w owns rows 8w..8w+7 of each
// 128-row tile and copies 2 KB chunks (8 cp.async of 8 B per lane) into private shared-memory
// slots, `slots` deep.  seg = 256-byte segments taken from ONE row per chunk: seg 1 is the k-block
// shape (8 rows x 256 B, every row a different DRAM page), seg 8 is 1 row x 2 KB contiguous.
__global__ void __launch_bounds__(512, 1)
stream_probe_kernel(const float* x, int rows, int dim, int seg, int slots, int cta_rows, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_begin = blockIdx.x * cta_rows;
  const int cta_end = min(cta_begin + cta_rows, rows);
  const int n_tiles = cta_end > cta_begin ? (cta_end - cta_begin) / 128 : 0;
  const int kbs = dim / 64;
  const int rows_per_chunk = 8 / seg;
  const int chunks_per_tile = kbs;   // 8 rows x kbs x 256 B per warp, 2 KB per chunk
  const uint32_t my = smem_u32(smem) + warp * 2048 + lane * 8;
  const size_t pitch = static_cast<size_t>(dim) * 4;
  const int total = n_tiles * chunks_per_tile;
  int pf = 0, pf_slot = 0;
  auto issue = [&]() {
    if (pf < total) {
      const int tile = pf / chunks_per_tile, c = pf - tile * chunks_per_tile;
      // chunk c: super-block sb of `seg` k-blocks, row group rg inside it
      const int groups = 8 / rows_per_chunk;            // == seg
      const int sb = c / groups, rg = c - sb * groups;
      const char* base = reinterpret_cast<const char*>(x) +
                         static_cast<size_t>(cta_begin + tile * 128 + warp * 8 + rg * rows_per_chunk) * pitch +
                         static_cast<size_t>(sb) * seg * 256 + lane * 8;
      const uint32_t dst = my + pf_slot * 32768;
      int i = 0;
      for (int r = 0; r < rows_per_chunk; ++r)
        for (int sgm = 0; sgm < seg; ++sgm, ++i)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + i * 256),
                       "l"(base + r * pitch + sgm * 256) : "memory");
      ++pf;
      if (++pf_slot == slots) pf_slot = 0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = 0; i < slots; ++i) issue();
  float acc = 0.f;
  int slot = 0;
  for (int q = 0; q < total; ++q) {
    switch (slots) {
      case 2: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
      case 3: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
      case 4: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
      case 5: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
      default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    }
    const uint32_t src = my + slot * 32768;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a, b;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(src + j * 256));
      acc += a + b;
    }
    issue();
    if (++slot == slots) slot = 0;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (acc == 123.456f) out[0] = acc;
}

}  // namespace fvtg

extern "C" int32_t fvtg_dbg_tma_probe(const void* w_bf16 /* [rows][256] */, int32_t rows, int32_t stages,
                                      int32_t units, int32_t grid, int32_t mode, void* out_cycles,
                                      void* stream) {
  using namespace fvtg;
  const int ub = mode == 2 ? 32768 : 16384;
  if (stages < 1 || stages * ub > 13 * 16384 || rows % 256) return fail(FVTG_EINVAL, "probe: bad arguments");
  CUtensorMap tw, tw2;
  FVTG_TRY(make_tmap_bf16(&tw, w_bf16, rows, 256, 256, 128, 64));
  FVTG_TRY(make_tmap_bf16(&tw2, w_bf16, rows, 256, 256, 256, 64));
  const int smem = 13 * 16384 + 512 + 1024;
  FVTG_CUDA_OK(cudaFuncSetAttribute(tma_stream_probe_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_stream_probe_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      tw, tw2, static_cast<const uint8_t*>(w_bf16), mode, stages, units, rows,
      static_cast<long long*>(out_cycles));
  FVTG_LAUNCH_CHECK("tma_stream_probe_kernel");
  return FVTG_OK;
}

extern "C" int32_t fvtg_dbg_mma_probe(int32_t N, int32_t iters, int32_t nbuf, int32_t grid,
                                      void* out_cycles, void* stream) {
  using namespace fvtg;
  if ((N != 128 && N != 256 && N != 64) || nbuf < 1 || nbuf > 4) return fail(FVTG_EINVAL, "mma probe: bad arguments");
  const int smem = 12 * 16384 + 1024;
  FVTG_CUDA_OK(cudaFuncSetAttribute(mma_rate_probe_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_probe_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      N, iters, nbuf, static_cast<long long*>(out_cycles));
  FVTG_LAUNCH_CHECK("mma_rate_probe_kernel");
  return FVTG_OK;
}

extern "C" int32_t fvtg_dbg_tma_probe3(const void* w_bf16 /* [2304][256] */, int32_t csize, int32_t nprod,
                                       int32_t stages, int32_t passes, int32_t grid, void* out_cycles,
                                       void* stream) {
  using namespace fvtg;
  if (stages < 1 || stages > 12 || (csize != 1 && csize != 2 && csize != 4) || nprod < 1 || nprod > 4 ||
      stages % nprod || grid % csize)
    return fail(FVTG_EINVAL, "probe3: bad arguments");
  CUtensorMap tw;
  FVTG_TRY(make_tmap_bf16(&tw, w_bf16, 2304, 256, 256, 128 / csize, 64));
  const int smem = 12 * 16384 + 512 + 1024;
  FVTG_CUDA_OK(cudaFuncSetAttribute(tma_stream_probe3_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  long long* oc = static_cast<long long*>(out_cycles);
  FVTG_CUDA_OK(cudaLaunchKernelEx(&cfg, tma_stream_probe3_kernel, tw, csize, nprod, stages, passes, oc));
  count_launch();
  return FVTG_OK;
}

extern "C" int32_t fvtg_dbg_mma_probe2(int32_t N, int32_t iters, int32_t cg, int32_t ts, int32_t grid,
                                       void* out_cycles, void* stream) {
  using namespace fvtg;
  if ((N != 128 && N != 256) || (cg != 1 && cg != 2) || grid % cg) return fail(FVTG_EINVAL, "mma probe2: bad arguments");
  const int smem = 12 * 16384 + 1024;
  FVTG_CUDA_OK(cudaFuncSetAttribute(mma_rate_probe2_kernel<1>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FVTG_CUDA_OK(cudaFuncSetAttribute(mma_rate_probe2_kernel<2>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cg;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cg == 2 ? 1 : 0;
  long long* oc = static_cast<long long*>(out_cycles);
  if (cg == 2) FVTG_CUDA_OK(cudaLaunchKernelEx(&cfg, mma_rate_probe2_kernel<2>, N, iters, ts, oc));
  else FVTG_CUDA_OK(cudaLaunchKernelEx(&cfg, mma_rate_probe2_kernel<1>, N, iters, ts, oc));
  count_launch();
  return FVTG_OK;
}

extern "C" int32_t fvtg_dbg_stream_probe(const float* x, int32_t rows, int32_t dim, int32_t seg, int32_t slots,
                                         float* out, void* stream) {
  using namespace fvtg;
  if ((seg != 1 && seg != 2 && seg != 4 && seg != 8) || slots < 2 || slots > 6 || dim % (64 * seg) || rows % 128)
    return fail(FVTG_EINVAL, "stream probe: bad arguments");
  const int smem = slots * 32768;
  FVTG_CUDA_OK(cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int tiles = rows / 128;
  const int ctas = tiles < sm_count() ? tiles : sm_count();
  const int cta_rows = (tiles + ctas - 1) / ctas * 128;
  const int grid = (rows + cta_rows - 1) / cta_rows;
  stream_probe_kernel<<<grid, 512, smem, static_cast<cudaStream_t>(stream)>>>(x, rows, dim, seg, slots, cta_rows, out);
  FVTG_LAUNCH_CHECK("stream_probe_kernel");
  return FVTG_OK;
}

// fvtg_dbg_mufu_probe: throughput of ex2.approx (MUFU) against fma (the issue-rate reference), per SM sub-partition.
namespace fvtg {
__global__ void mufu_probe_kernel(int iters, float* out) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i + 1);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i];
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = 1.f + 1e-3f * (threadIdx.x + i);
  __syncthreads();
  long long t2 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(0.999f), "f"(1e-4f));
  }
  long long t3 = clock64();
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += y[i];
  // cvt.rn.bf16x2.f32 (F2FP.BF16.F32.PACK_AB): 8 independent chains, the packed result re-enters as a float
  uint32_t z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) z[i] = __float_as_uint(1.f + 1e-3f * (threadIdx.x + i));
  __syncthreads();
  long long t4 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(z[i]) : "f"(__uint_as_float(z[i])), "f"(__uint_as_float(z[(i + 1) & 7])));
  }
  long long t5 = clock64();
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float(z[i]);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[3] = static_cast<float>(t5 - t4) / (8.f * iters * fmaxf(static_cast<float>(blockDim.x) / 128.f, 1.f));
    // warps per sub-partition = blockDim / 128; cycles per warp-instruction on one sub-partition
    const float per_smsp = static_cast<float>(blockDim.x) / 128.f;
    out[0] = static_cast<float>(t1 - t0) / (8.f * iters * fmaxf(per_smsp, 1.f));
    out[1] = static_cast<float>(t3 - t2) / (8.f * iters * fmaxf(per_smsp, 1.f));
    out[2] = acc;
  }
}
}  // namespace fvtg
extern "C" int32_t fvtg_dbg_mufu_probe(int32_t warps, int32_t iters, float* out, void* stream) {
  using namespace fvtg;
  if (warps < 1 || warps > 32 || iters < 1) return fail(FVTG_EINVAL, "mufu_probe: warps 1..32");
  mufu_probe_kernel<<<148, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(iters, out);
  FVTG_LAUNCH_CHECK("mufu_probe_kernel");
  return FVTG_OK;
}

// fvtg_dbg_store_probe: per-SM global store / load throughput with the layer kernel's epilogue access shapes.
// 512 threads; per repetition a CTA touches one 128-row tile: mode 0 = fp32 tile-blocked stores (16 x 16 B per
// thread, a warp writes 512 contiguous bytes), 1 = bf16 row-major rows (8 x 32 B per thread at a 512 B row pitch,
// two outputs), 2 = both (the LayerNorm-2 mix, 256 KB), 3 = fp32 tile-blocked loads (the residual read).
namespace fvtg {
__global__ void __launch_bounds__(512, 1)
store_probe_kernel(int mode, int reps, uint8_t* buf, long long* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, qt = warp >> 2, r = q * 32 + lane;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    uint8_t* tile = buf + (static_cast<size_t>(rep) * gridDim.x + blockIdx.x) * (256u << 10);
    float* yf = reinterpret_cast<float*>(tile) + qt * (16 * 512) + r * 4;
    if (mode == 0 || mode == 2) {
#pragma unroll
      for (int g4 = 0; g4 < 16; ++g4)
        *reinterpret_cast<float4*>(yf + g4 * 512) = make_float4(rep, g4, r, qt);
    }
    if (mode == 1 || mode == 2) {
      uint8_t* b0 = tile + (128u << 10) + static_cast<size_t>(r) * 512 + qt * 128;
#pragma unroll
      for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[8] = {1u, 2u, 3u, 4u, 5u, 6u, 7u, static_cast<uint32_t>(rep)};
          st_global_v8(reinterpret_cast<bf16*>(b0 + o * (64u << 10) + j * 32), w);
        }
    }
    if (mode >= 4) {   // TMA bulk stores from shared memory: 4 = 4 x 32 KB, 5 = 16 x 8 KB, 6 = 64 x 2 KB per tile
      extern __shared__ __align__(1024) uint8_t sm[];
      const int n = mode == 4 ? 4 : (mode == 5 ? 16 : 64);
      const uint32_t sz = (128u << 10) / n;
      if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(tile + i * sz),
                       "r"(smem_u32(sm) + (i * sz) % 32768u), "r"(sz) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (mode == 3) {
#pragma unroll
      for (int g4 = 0; g4 < 16; ++g4) {
        const float4 v = *reinterpret_cast<const float4*>(yf + g4 * 512);
        acc += v.x + v.y + v.z + v.w;
      }
    }
  }
  const long long t1 = clock64();
  if (mode >= 4 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  __syncthreads();
  const long long t2 = clock64();
  if (threadIdx.x == 0) {
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  if (acc == 123.456f) out[0] = 0;
}
}  // namespace fvtg
extern "C" int32_t fvtg_dbg_store_probe(int32_t mode, int32_t reps, int32_t grid, void* buf, void* out_cycles,
                                        void* stream) {
  using namespace fvtg;
  if (mode < 0 || mode > 6 || reps < 1 || grid < 1) return fail(FVTG_EINVAL, "store probe: bad arguments");
  store_probe_kernel<<<grid, 512, 32768, static_cast<cudaStream_t>(stream)>>>(mode, reps, static_cast<uint8_t*>(buf),
                                                                         static_cast<long long*>(out_cycles));
  FVTG_LAUNCH_CHECK("store_probe_kernel");
  return FVTG_OK;
}

#endif  // FVTG_DEBUG_HOOKS
