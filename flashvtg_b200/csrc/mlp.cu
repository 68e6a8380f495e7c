// Score-head MLP chain (ConfidenceScorer.fc, FlashVTG/model.py:44-71: Linear(256,128) -> ReLU ->
// [Linear(128,128) -> ReLU] x (num_mlp_layers - 2) -> Linear(128,1)) as ONE persistent tcgen05 kernel:
// the 128-wide activations of a 128-row tile never leave the SM.
//
//   layer 0 : D = A[128 x 256] (shared memory, TMA) . W0^T            16 x tcgen05.mma 128x128x16 (SS)
//   layer m : D = act[128 x 128] (TMEM) . Wm^T                         8 x tcgen05.mma 128x128x16 (TS)
//   epilogue of a hidden layer: relu(D + b) -> bf16 pairs written back into the SAME TMEM columns by
//   the thread that read them (they become the next layer's A operand); the last hidden layer's
//   epilogue takes the dot product with the final Linear(128,1) weight and stores the logit.
//
// Two row tiles are in flight per CTA (TMEM columns 0-255 / 256-511, two A buffers): while the 16
// epilogue warps work on one tile's layer, the tensor pipe runs the other tile's; each layer's weight
// units are loaded once per tile pair.  20 warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator,
// 4..19 epilogue (TMEM lane quadrant = warp % 4, 32-column quarter = (warp - 4) / 4).
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int MK_THREADS = 640;
constexpr int MK_UNIT = 128 * 64 * 2;         // 16 KB
constexpr int MK_STAGES = 4;
constexpr int MK_OFF_A = 0;                   // 2 slots x 4 units
constexpr int MK_OFF_W = 8 * MK_UNIT;         // weight ring
constexpr int MK_OFF_BAR = MK_OFF_W + MK_STAGES * MK_UNIT;
constexpr int MK_OFF_PART = MK_OFF_BAR + 256;                   // float [2 slots][4 quarters][128]
constexpr int MK_OFF_PAR = MK_OFF_PART + 2 * 4 * 128 * 4;       // biases [7][128] + last_w [128]
constexpr int MK_SMEM_BYTES = MK_OFF_PAR + 8 * 128 * 4 + 1024;
static_assert(MK_SMEM_BYTES <= 232448, "mlp kernel shared memory over the 227 KB limit");

__device__ __forceinline__ void mk_epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// (video, canonical point index) of a row of the head row spaces H1 / H2 (common.cuh PyrGeo)
__device__ __forceinline__ bool head_row(const PyrGeo& geo, int h2, int row, int M, int* b_out, int* n_out) {
  if (row >= M) return false;
  if (!h2) {
    const int b = row / geo.PH1, q = row - b * geo.PH1;
    int l = 0;
    for (int i = 1; i < geo.nlev; ++i)
      if (q >= geo.o1[i]) l = i;
    const int i = q - geo.o1[l];
    const int vl = geo.vlen[b];
    int off = 0;
    for (int k = 0; k < l; ++k) off += vl >> k;
    *b_out = b;
    *n_out = off + i;
    return i >= 0 && i < (vl >> l);
  }
  const int b = row / geo.PH2, q = row - b * geo.PH2;
  const int vl = geo.vlen[b];
  int nv = 0;
  for (int k = 0; k < geo.nlev; ++k) nv += vl >> k;
  const int n = q - geo.pad;
  *b_out = b;
  *n_out = n;
  return n >= 0 && n < nv;
}

struct MlpArgs {
  int M;            // rows of the head row space
  int nl;           // hidden layers (num_mlp_layers - 1), 1..7
  int h2;           // 0: H1 row space (class head), 1: H2 (conf head)
  float last_b;
  const float* bias[7];
  const float* last_w;
  float* out;       // [B][n_max]
  PyrGeo geo;
};

struct MlpMaps {
  CUtensorMap a;
  CUtensorMap w[7];
};

__global__ void __launch_bounds__(MK_THREADS, 1)
mlp_chain_kernel(const __grid_constant__ MlpMaps tm, const __grid_constant__ MlpArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + MK_OFF_A;
  uint8_t* sW = smem + MK_OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MK_OFF_BAR);
  uint64_t* wfull = bars;              // [4] weight unit landed
  uint64_t* wempty = bars + 4;         // [4] both slots' MMAs on the unit retired
  uint64_t* a_full = bars + 8;         // [2] A tile of the slot landed
  uint64_t* a_empty = bars + 10;       // [2] layer-0 MMAs of the slot retired
  uint64_t* d_full = bars + 12;        // [2] layer accumulated for the slot
  uint64_t* h_ready = bars + 14;       // [2] activations of the slot stored back (16 warps)
  uint64_t* t_free = bars + 16;        // [2] last layer of the slot's tile read out of TMEM (16 warps)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 20);
  float* s_part = reinterpret_cast<float*>(smem + MK_OFF_PART);
  float* s_bias = reinterpret_cast<float*>(smem + MK_OFF_PAR);
  float* s_lastw = s_bias + 7 * 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (g.M + 127) >> 7;
  const int npairs = (ntiles + 1) >> 1;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm.a);
    for (int m = 0; m < g.nl; ++m) prefetch_tmap(&tm.w[m]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MK_STAGES; ++s) {
      mbar_init(&wfull[s], 1);
      mbar_init(&wempty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&d_full[s], 1);
      mbar_init(&h_ready[s], 16);
      mbar_init(&t_free[s], 16);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < g.nl * 128; i += MK_THREADS) s_bias[i] = g.bias[i >> 7][i & 127];
  for (int i = threadIdx.x; i < 128; i += MK_THREADS) s_lastw[i] = g.last_w[i];
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer --
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      int it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        for (int s = 0; s < 2; ++s) {
          const int tile = 2 * pair + s;
          if (tile >= ntiles) break;
          mbar_wait(&a_empty[s], (it & 1) ^ 1);
          mbar_expect_tx(&a_full[s], 4 * MK_UNIT);
          for (int kb = 0; kb < 4; ++kb)
            tma_load_2d(sA + (s * 4 + kb) * MK_UNIT, &tm.a, kb * 64, tile * 128, &a_full[s]);
        }
        for (int m = 0; m < g.nl; ++m) {
          const int units = m == 0 ? 4 : 2;
          for (int u = 0; u < units; ++u) {
            mbar_wait(&wempty[ws], wph ^ 1);
            mbar_expect_tx(&wfull[ws], MK_UNIT);
            tma_load_2d(sW + ws * MK_UNIT, &tm.w[m], u * 64, 0, &wfull[ws]);
            if (++ws == MK_STAGES) { ws = 0; wph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      uint32_t hcnt[2] = {0, 0};   // completed h_ready phases per slot
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
      int it = 0;
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        const int nslots = (2 * pair + 1 < ntiles) ? 2 : 1;
        for (int m = 0; m < g.nl; ++m) {
          const int units = m == 0 ? 4 : 2;
          // the layer's weight units stay in the ring for both slots
          int us[4];
          for (int u = 0; u < units; ++u) {
            mbar_wait(&wfull[ws], wph);
            us[u] = ws;
            if (++ws == MK_STAGES) { ws = 0; wph ^= 1; }
          }
          tc_fence_after();
          for (int s = 0; s < nslots; ++s) {
            const uint32_t d = tmem + s * 256 + (m & 1) * 128;
            if (m == 0) {
              // the epilogue must be done with the slot's previous tile: otherwise d_full[s] could run
              // two phases ahead of its waiter (parity aliasing) and layer 0 could overwrite
              // accumulator columns that are still being read
              mbar_wait(&t_free[s], (it & 1) ^ 1);
              mbar_wait(&a_full[s], it & 1);
              tc_fence_after();
              for (int u = 0; u < 4; ++u) {
                const uint64_t da = umma_desc_sw128(sA_u + (s * 4 + u) * MK_UNIT);
                const uint64_t db = umma_desc_sw128(sW_u + us[u] * MK_UNIT);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (u | k) ? 1u : 0u);
              }
              umma_commit(&a_empty[s]);
            } else {
              mbar_wait(&h_ready[s], hcnt[s] & 1u);
              ++hcnt[s];
              tc_fence_after();
              const uint32_t a = tmem + s * 256 + ((m - 1) & 1) * 128;   // bf16 activations of layer m-1
              for (int u = 0; u < 2; ++u) {
                const uint64_t db = umma_desc_sw128(sW_u + us[u] * MK_UNIT);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int j = u * 4 + k;
                  umma_bf16_ts(d, a + 32 * (j >> 1) + 8 * (j & 1), db + 2 * k, idesc, j ? 1u : 0u);
                }
              }
            }
            umma_commit(&d_full[s]);
          }
          for (int u = 0; u < units; ++u) umma_commit(&wempty[us[u]]);
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue --
    const int ew = warp - 4;
    const int q = ew & 3, qt = ew >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t dcnt[2] = {0, 0};
    uint32_t u[32];
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int nslots = (2 * pair + 1 < ntiles) ? 2 : 1;
      for (int m = 0; m < g.nl; ++m) {
        const bool last = m == g.nl - 1;
        for (int s = 0; s < nslots; ++s) {
          mbar_wait(&d_full[s], dcnt[s] & 1u);
          ++dcnt[s];
          tc_fence_after();
          const uint32_t ta = tmem + lane_addr + s * 256 + (m & 1) * 128 + qt * 32;
          tmem_ld32(ta, u);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + m * 128 + qt * 32);
          if (!last) {
            uint32_t hp[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = b4[i];
              const float x0 = fmaxf(__uint_as_float(u[4 * i + 0]) + b.x, 0.f);
              const float x1 = fmaxf(__uint_as_float(u[4 * i + 1]) + b.y, 0.f);
              const float x2 = fmaxf(__uint_as_float(u[4 * i + 2]) + b.z, 0.f);
              const float x3 = fmaxf(__uint_as_float(u[4 * i + 3]) + b.w, 0.f);
              hp[2 * i + 0] = pack_bf16(x0, x1);
              hp[2 * i + 1] = pack_bf16(x2, x3);
            }
            tmem_st16(ta, hp);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_ready[s]);
          } else {
            // relu -> dot with the final Linear(128, 1) weight; the four column quarters meet in shared memory
            const float4* w4 = reinterpret_cast<const float4*>(s_lastw + qt * 32);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = b4[i], w = w4[i];
              dot += fmaxf(__uint_as_float(u[4 * i + 0]) + b.x, 0.f) * w.x;
              dot += fmaxf(__uint_as_float(u[4 * i + 1]) + b.y, 0.f) * w.y;
              dot += fmaxf(__uint_as_float(u[4 * i + 2]) + b.z, 0.f) * w.z;
              dot += fmaxf(__uint_as_float(u[4 * i + 3]) + b.w, 0.f) * w.w;
            }
            float* part = s_part + s * 512;
            if (qt > 0) part[qt * 128 + r] = dot;
            tc_fence_before();
            mk_epi_bar();
            if (qt == 0) {
              int b_, n_;
              if (head_row(g.geo, g.h2, (2 * pair + s) * 128 + r, g.M, &b_, &n_))
                g.out[static_cast<size_t>(b_) * g.geo.n_max + n_] =
                    ((dot + part[128 + r]) + (part[256 + r] + part[384 + r])) + g.last_b;
            }
            // part[] of this slot is rewritten two layers-of-work later, behind the next barrier of the
            // other slot or pair: a trailing barrier keeps fast warps from overwriting it early
            mk_epi_bar();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_free[s]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// a: bf16 [M][256] (head conv output); w[m]: bf16 [128][256] (m = 0) / [128][128]
int launch_mlp_chain(cudaStream_t st, const bf16* a, const void* const* w, const MlpHostArgs& h) {
  if (h.M <= 0) return FVTG_OK;
  if (h.nl < 1 || h.nl > 7) return fail(FVTG_EINVAL, "mlp_chain: 1..7 hidden layers");
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      MK_SMEM_BYTES));
    attr_set = true;
  }
  MlpMaps tm;
  memset(&tm, 0, sizeof(tm));
  FVTG_TRY(make_tmap_bf16(&tm.a, a, h.M, 256, 256, 128, 64));
  for (int m = 0; m < h.nl; ++m) {
    const int k = m == 0 ? 256 : 128;
    FVTG_TRY(make_tmap_bf16(&tm.w[m], w[m], 128, k, k, 128, 64));
  }
  for (int m = h.nl; m < 7; ++m) tm.w[m] = tm.w[0];
  MlpArgs g;
  memset(&g, 0, sizeof(g));
  g.M = h.M; g.nl = h.nl; g.h2 = h.h2; g.last_b = h.last_b;
  for (int m = 0; m < h.nl; ++m) g.bias[m] = h.bias[m];
  g.last_w = h.last_w; g.out = h.out; g.geo = h.geo;
  const int tiles = (h.M + 127) / 128;
  const int pairs = (tiles + 1) / 2;
  const int grid = pairs < sm_count() ? pairs : sm_count();
  ProfScope prof(st, PC_GEMM);
  FVTG_CUDA_OK(launch_pdl(mlp_chain_kernel, dim3(grid), dim3(MK_THREADS), MK_SMEM_BYTES, st, tm, g));
  FVTG_LAUNCH_CHECK("mlp_chain_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
