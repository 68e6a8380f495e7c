// Debug probe (not on the product path): how fast can every SM stream the SAME weight matrix
// from L2 through TMA into a shared-memory ring of `stages` 16 KB units, with no consumer work?
// Answers whether the fused layer kernel's weight ring is latency- or bandwidth-limited.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

__global__ void __launch_bounds__(64, 1)
tma_stream_probe_kernel(const __grid_constant__ CUtensorMap tmW, int stages, int units, int rows_total,
                        long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 13 * 16384);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
    prefetch_tmap(&tmW);
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {  // producer
    int s = 0; uint32_t ph = 0;
    for (int u = 0; u < units; ++u) {
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], 16384);
      const int r0 = (u * 128) % rows_total;
      tma_load_2d(smem + s * 16384, &tmW, ((u / (rows_total / 128)) % 4) * 64, r0, &full[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer: release immediately
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

}  // namespace fvtg

extern "C" int32_t fvtg_dbg_tma_probe(const void* w_bf16 /* [rows][256] */, int32_t rows, int32_t stages,
                                      int32_t units, int32_t grid, void* out_cycles, void* stream) {
  using namespace fvtg;
  if (stages < 1 || stages > 13 || rows % 128) return fail(FVTG_EINVAL, "probe: bad arguments");
  CUtensorMap tw;
  FVTG_TRY(make_tmap_bf16(&tw, w_bf16, rows, 256, 256, 128, 64));
  const int smem = 13 * 16384 + 512 + 1024;
  FVTG_CUDA_OK(cudaFuncSetAttribute(tma_stream_probe_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_stream_probe_kernel<<<grid, 64, smem, static_cast<cudaStream_t>(stream)>>>(
      tw, stages, units, rows, static_cast<long long*>(out_cycles));
  FVTG_LAUNCH_CHECK("tma_stream_probe_kernel");
  return FVTG_OK;
}
