/* Debug / test hooks of flashvtg_b200.  They exist ONLY in libflashvtg_b200_dbg.so - the same sources as
 * libflashvtg_b200.so compiled with -DFVTG_DEBUG_HOOKS plus csrc/probe.cu - which the kernel-level unit tests
 * (tests/test_gpu_parity.py: GEMM and first-projection kernels in isolation) and the tools/ trace and probe
 * scripts load.  The product library exports none of them. */
#ifndef FLASHVTG_B200_DBG_H_
#define FLASHVTG_B200_DBG_H_
#include "flashvtg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug hook: device buffer (>= 4096 int64) that the fused layer kernel's CTA 0 fills with clock64
 * stamps per pipeline phase (tools/trace_layer.py); null (default) disables tracing. */
void fvtg_dbg_set_trace(void* device_buf);

/* Test hook: out = act(A[M][K] * W[N][K]^T + bias) through the production tcgen05 GEMM.
 * A, W bf16 (K multiple of 64, N multiple of 128).  N == 256: out is fp32 [M][256] (full-row
 * epilogue); otherwise out is bf16 [M][N] (tile epilogue).  act: 0 none, 1 relu. */
int32_t fvtg_dbg_gemm(const void* a, const void* w, const float* bias, float* out, int32_t M,
                      int32_t N, int32_t K, int32_t act, void* stream);

/* Test / tuning hook: the fused first input projection alone (FvtgInProj.fc0 semantics):
 * out bf16 [rows][256] = LN_256(ReLU(LN_dim(x) . W^T + b)) with x fp32 [rows][dim] (dim even),
 * wg bf16 [256][dim_pad] = W . diag(gamma) zero padded to dim_pad (multiple of 64), wsum = row sums
 * of wg, cfold = W . beta + b, (g1, b1) the LayerNorm(256) of the next layer. */
int32_t fvtg_dbg_inproj(const float* x, int32_t rows, int32_t dim, int32_t dim_pad, const void* wg,
                        const float* wsum, const float* cfold, const float* g1, const float* b1,
                        void* out, void* stream);

/* Micro-benchmarks of csrc/probe.cu (tools/probe_*.py): TMA streaming rates, tcgen05.mma issue rates,
 * HBM streaming rate of the first projection's access shape.  See the tools for the argument meaning. */
int32_t fvtg_dbg_stream_probe(const float* x, int32_t rows, int32_t dim, int32_t seg, int32_t slots,
                              float* out, void* stream);
/* out[0] = cycles per warp-wide ex2.approx.ftz.f32 with `warps` (1..32) warps per SM issuing 8 independent chains each,
 * out[1] = the same for fma.rn.f32 (the issue-rate reference). */
int32_t fvtg_dbg_mufu_probe(int32_t warps, int32_t iters, float* out, void* stream);
/* Per-SM global store / load rate with the layer kernel's epilogue access shapes: mode 0 fp32 tile-blocked stores,
 * 1 bf16 row-major stores, 2 both (256 KB per tile), 3 fp32 tile-blocked loads; `grid` CTAs x `reps` tiles each over
 * buf (>= grid * reps * 256 KB); out_cycles int64 [grid][2] = (issue loop, until the stores are performed). */
int32_t fvtg_dbg_store_probe(int32_t mode, int32_t reps, int32_t grid, void* buf, void* out_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* FLASHVTG_B200_DBG_H_ */
