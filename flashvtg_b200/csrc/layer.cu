// Fused transformer-layer tail: everything of a post-norm layer that is row-wise, in ONE
// persistent tcgen05 kernel per 128-row tile, with the 1024-wide hidden activations never leaving
// the SM:
//
//   z  = x + att . Wo^T + bo                         (out_proj + residual)
//   T2V layer (transformer.py:359-367):  y = LN2( z  + W2 . PReLU(W1 . LN1(z) + b1) + b2 )
//   SA  layer (transformer.py:416-420):  y = LN2( x1 + W2 . PReLU(W1 . x1 + b1) + b2 ),  x1 = LN1(z)
//
// Data flow per tile (TMEM: D_z = columns 0..255, D_h[2] = columns 256..383 / 384..511):
//   TMA: att tile -> sA ; weight "units" ([128 rows][64 k] bf16, 16 KB) -> 5-stage ring
//   MMA: D_z  = sA . Wo^T                                   (8 units)
//   EPI: z = D_z + bo + x (fp32 residual stream, tile-blocked layout => coalesced 16 B / lane);
//        LayerNorm1 -> bf16 -> sA (SWIZZLE_128B K-major, written by the epilogue threads);
//        D_z <- z (T2V) or LN1(z) (SA) via tcgen05.st: the FFN residual lives in the accumulator
//   for the 8 hidden pieces p of 128:   (ff1(p+1) is issued before ff2(p): MMA never waits on EPI)
//        MMA: D_h[p&1] = sA . W1[p]^T                        (4 units)
//        EPI: h = PReLU(D_h + b1[p]) -> bf16 -> sH[p&1]      (SWIZZLE_128B K-major)
//        MMA: D_z += sH[p&1] . W2[:, p]^T                    (4 units; accumulates onto the residual)
//   EPI: y = LayerNorm2(D_z + b2) -> fp32 residual stream, bf16(y), bf16(y + pos)
//
// 12 warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue (two threads
// per row: TMEM lane quadrant = warp % 4, column half = (warp - 4) / 4).
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int LK_THREADS = 384;
constexpr int LK_STAGES = 5;
constexpr int LK_UNIT = 128 * 64 * 2;       // 16 KB: [128 rows][64 bf16], SWIZZLE_128B
constexpr int LK_OFF_A = 0;                 // 4 units: att tile, then LN1 output
constexpr int LK_OFF_H = 4 * LK_UNIT;       // 2 buffers x 2 units: hidden piece
constexpr int LK_OFF_W = 8 * LK_UNIT;       // weight ring
constexpr int LK_OFF_BAR = LK_OFF_W + LK_STAGES * LK_UNIT;
constexpr int LK_OFF_STAT = LK_OFF_BAR + 256;
constexpr int LK_OFF_PAR = LK_OFF_STAT + 2 * 2 * 128 * 8;
constexpr int LK_PAR_FLOATS = 256 * 3 + 1024 + 256 * 3;
constexpr int LK_SMEM_BYTES = LK_OFF_PAR + LK_PAR_FLOATS * 4 + 1024 /*align slack*/;
static_assert(LK_SMEM_BYTES <= 232448, "layer kernel shared memory over the 227 KB limit");

// trace slots: [role 0 MMA / 1 epilogue][tile it < 8][event < 32]
#define LK_TRACE(role, ev)                                                              \
  do {                                                                                  \
    if (g.trace && blockIdx.x == 0 && it < 8)                                           \
      g.trace[((role) * 8 + it) * 32 + (ev)] = clock64();                               \
  } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(LK_THREADS, 1)
layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWo,
             const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
             const __grid_constant__ LayerArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + LK_OFF_A;
  uint8_t* sH = smem + LK_OFF_H;
  uint8_t* sW = smem + LK_OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LK_OFF_BAR);
  uint64_t* full = bars;                    // [5] TMA -> MMA
  uint64_t* empty = bars + LK_STAGES;       // [5] MMA -> TMA
  uint64_t* a_full = bars + 10;             // att tile landed
  uint64_t* a_empty = bars + 11;            // last ff1 MMA retired: sA reusable
  uint64_t* z1_full = bars + 12;            // out_proj accumulated
  uint64_t* z2_full = bars + 13;            // FFN accumulated
  uint64_t* z_empty = bars + 14;            // final epilogue drained D_z
  uint64_t* ln_ready = bars + 15;           // LN1 tile in sA, residual in D_z
  uint64_t* hacc_full = bars + 16;          // [2] ff1 piece accumulated
  uint64_t* h_ready = bars + 18;            // [2] hidden piece in sH, D_h drained
  uint64_t* h_empty = bars + 20;            // [2] ff2 retired: sH reusable
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 24);
  float2* s_stat = reinterpret_cast<float2*>(smem + LK_OFF_STAT);  // [2 phases][2 halves][128]
  float* s_par = reinterpret_cast<float*>(smem + LK_OFF_PAR);
  float* s_bo = s_par;
  float* s_g1 = s_par + 256;
  float* s_be1 = s_par + 512;
  float* s_b1 = s_par + 768;
  float* s_b2 = s_par + 1792;
  float* s_g2 = s_par + 2048;
  float* s_be2 = s_par + 2304;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = (g.M + 127) >> 7;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmWo);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < LK_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(z1_full, 1);
    mbar_init(z2_full, 1);
    mbar_init(z_empty, 8);
    mbar_init(ln_ready, 8);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hacc_full[b], 1);
      mbar_init(&h_ready[b], 8);
      mbar_init(&h_empty[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += LK_THREADS) {
    s_bo[i] = g.bo[i];
    s_g1[i] = g.g1[i];
    s_be1[i] = g.be1[i];
    s_b2[i] = g.b2[i];
    s_g2[i] = g.g2[i];
    s_be2[i] = g.be2[i];
  }
  for (int i = threadIdx.x; i < 1024; i += LK_THREADS) s_b1[i] = g.b1[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer --
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto unit = [&](const CUtensorMap* tm, int c0, int r0) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], LK_UNIT);
        tma_load_2d(sW + s * LK_UNIT, tm, c0, r0, &full[s]);
        if (++s == LK_STAGES) { s = 0; ph ^= 1; }
      };
      auto ff1 = [&](int p) {
        for (int kb = 0; kb < 4; ++kb) unit(&tmW1, kb * 64, p * 128);
      };
      auto ff2 = [&](int p) {
        for (int kb2 = 0; kb2 < 2; ++kb2)
          for (int nh = 0; nh < 2; ++nh) unit(&tmW2, p * 128 + kb2 * 64, nh * 128);
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        mbar_wait(a_empty, (it & 1) ^ 1);
        mbar_expect_tx(a_full, 4 * LK_UNIT);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sA + kb * LK_UNIT, &tmA, kb * 64, tile * 128, a_full);
        for (int kb = 0; kb < 4; ++kb)
          for (int nh = 0; nh < 2; ++nh) unit(&tmWo, kb * 64, nh * 128);
        ff1(0);
        ff1(1);
        for (int p = 0; p < 8; ++p) {
          ff2(p);
          if (p + 2 < 8) ff1(p + 2);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t sA_u = smem_u32(sA), sH_u = smem_u32(sH), sW_u = smem_u32(sW);
      auto unit = [&](uint32_t d_tmem, uint32_t a_addr, bool acc_first) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(a_addr);
        const uint64_t db = umma_desc_sw128(sW_u + s * LK_UNIT);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (acc_first || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (++s == LK_STAGES) { s = 0; ph ^= 1; }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        auto ff1 = [&](int p) {
          const int buf = p & 1;
          for (int kb = 0; kb < 4; ++kb)
            unit(tmem + 256 + buf * 128, sA_u + kb * LK_UNIT, kb > 0);
          umma_commit(&hacc_full[buf]);
          if (p == 7) umma_commit(a_empty);
        };
        auto ff2 = [&](int p) {
          const int buf = p & 1;
          const uint32_t n = static_cast<uint32_t>(it) * 4u + static_cast<uint32_t>(p >> 1);
          mbar_wait(&h_ready[buf], n & 1u);
          tc_fence_after();
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int nh = 0; nh < 2; ++nh)
              unit(tmem + nh * 128, sH_u + buf * 2 * LK_UNIT + kb2 * LK_UNIT, true);
          umma_commit(&h_empty[buf]);
        };
        LK_TRACE(0, 0);
        mbar_wait(z_empty, (it & 1) ^ 1);
        LK_TRACE(0, 1);
        mbar_wait(a_full, it & 1);
        tc_fence_after();
        LK_TRACE(0, 2);
        for (int kb = 0; kb < 4; ++kb)
          for (int nh = 0; nh < 2; ++nh) unit(tmem + nh * 128, sA_u + kb * LK_UNIT, kb > 0);
        umma_commit(z1_full);
        LK_TRACE(0, 3);
        mbar_wait(ln_ready, it & 1);
        tc_fence_after();
        LK_TRACE(0, 4);
        ff1(0);
        ff1(1);
        LK_TRACE(0, 5);
        for (int p = 0; p < 8; ++p) {
          ff2(p);
          LK_TRACE(0, 6 + 2 * p);
          if (p + 2 < 8) ff1(p + 2);
          LK_TRACE(0, 7 + 2 * p);
        }
        umma_commit(z2_full);
        LK_TRACE(0, 22);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue --
    const int ew = warp - 4;
    const int q = ew & 3;    // TMEM lane quadrant (== warp % 4)
    const int hf = ew >> 2;  // column half
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t u[32];
    float v[32];
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int row = tile * 128 + r;
      const bool inb = row < g.M;
      // fp32 residual stream, tile-blocked: [tile][col/4][row%128][4]
      float* yblk = g.yf + static_cast<size_t>(tile) * (128 * 256) + r * 4;

      // ---- epilogue 1: z = D_z + bo + x ; LayerNorm1 -> sA ; residual back into D_z ----------
      const bool tr = (warp == 4 && lane == 0);
      if (tr) LK_TRACE(1, 0);
      mbar_wait(z1_full, it & 1);
      tc_fence_after();
      if (tr) LK_TRACE(1, 1);
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int c0 = hf * 128 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        float4 rr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rr[i] = inb ? *reinterpret_cast<const float4*>(yblk + ((c0 >> 2) + i) * 512)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[4 * i + 0] = __uint_as_float(u[4 * i + 0]) + s_bo[c0 + 4 * i + 0] + rr[i].x;
          v[4 * i + 1] = __uint_as_float(u[4 * i + 1]) + s_bo[c0 + 4 * i + 1] + rr[i].y;
          v[4 * i + 2] = __uint_as_float(u[4 * i + 2]) + s_bo[c0 + 4 * i + 2] + rr[i].z;
          v[4 * i + 3] = __uint_as_float(u[4 * i + 3]) + s_bo[c0 + 4 * i + 3] + rr[i].w;
        }
        if (c == 0) shift = v[0];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = v[j] - shift;
          s1 += d;
          s2 += d * d;
          u[j] = __float_as_uint(v[j]);
        }
        tmem_st32(tmem + lane_addr + c0, u);
      }
      tmem_st_wait();
      if (tr) LK_TRACE(1, 2);
      s_stat[hf * 128 + r] = make_float2(shift + s1 * (1.f / 128.f), s2 - s1 * s1 * (1.f / 128.f));
      epi_bar_sync();
      float mean, rstd;
      {
        const float2 a = s_stat[r], b = s_stat[128 + r];
        const float dm = a.x - b.x;
        mean = 0.5f * (a.x + b.x);
        const float var = fmaxf((a.y + b.y + dm * dm * 64.f) * (1.f / 256.f), 0.f);
        rstd = rsqrtf(var + 1e-5f);
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int c0 = hf * 128 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = (__uint_as_float(u[j]) - mean) * rstd * s_g1[c0 + j] + s_be1[c0 + j];
        st_shared_bf16x32(sA + (c0 >> 6) * LK_UNIT, r, (c0 & 63) >> 3, v);
        if (g.mode == LAYER_SA) {
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(v[j]);
          tmem_st32(tmem + lane_addr + c0, u);
        }
      }
      if (g.mode == LAYER_SA) tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ln_ready);
      if (tr) LK_TRACE(1, 3);

      // ---- epilogue 2 (x8): hidden piece = PReLU(D_h + b1) -> sH ------------------------------
#pragma unroll 1
      for (int p = 0; p < 8; ++p) {
        const int buf = p & 1;
        const uint32_t n = static_cast<uint32_t>(it) * 4u + static_cast<uint32_t>(p >> 1);
        mbar_wait(&hacc_full[buf], n & 1u);
        mbar_wait(&h_empty[buf], (n & 1u) ^ 1u);
        tc_fence_after();
        if (tr) LK_TRACE(1, 4 + 2 * p);
        uint8_t* hu = sH + buf * 2 * LK_UNIT + hf * LK_UNIT;  // k-block hf of the piece
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int c0 = hf * 64 + c * 32;
          tmem_ld32(tmem + lane_addr + 256 + buf * 128 + c0, u);
          tmem_ld_wait();
          const float* bb = s_b1 + p * 128 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(u[j]) + bb[j];
            v[j] = x > 0.f ? x : g.prelu * x;
          }
          st_shared_bf16x32(hu, r, c * 4, v);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[buf]);
        if (tr) LK_TRACE(1, 5 + 2 * p);
      }

      // ---- final epilogue: y = LayerNorm2(D_z + b2) -> residual stream / bf16 operands --------
      mbar_wait(z2_full, it & 1);
      tc_fence_after();
      if (tr) LK_TRACE(1, 20);
      shift = 0.f; s1 = 0.f; s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int c0 = hf * 128 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        tmem_ld_wait();
        if (c == 0) shift = __uint_as_float(u[0]) + s_b2[c0];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(u[j]) + s_b2[c0 + j] - shift;
          s1 += d;
          s2 += d * d;
        }
      }
      s_stat[256 + hf * 128 + r] = make_float2(shift + s1 * (1.f / 128.f), s2 - s1 * s1 * (1.f / 128.f));
      epi_bar_sync();
      if (tr) LK_TRACE(1, 21);
      {
        const float2 a = s_stat[256 + r], b = s_stat[256 + 128 + r];
        const float dm = a.x - b.x;
        mean = 0.5f * (a.x + b.x);
        const float var = fmaxf((a.y + b.y + dm * dm * 64.f) * (1.f / 256.f), 0.f);
        rstd = rsqrtf(var + 1e-5f);
      }
      int prow = row;
      if (g.pos_mod > 0) prow = row % g.pos_mod;
      const bool st_pos = g.out_pb && inb && (g.pos_rowlim <= 0 || prow < g.pos_rowlim);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int c0 = hf * 128 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        float4 pp[8];
        if (st_pos && g.pos) {
          if (g.pos_mod > 0) {
            const float* ps = g.pos + static_cast<size_t>(prow) * 256 + c0;
#pragma unroll
            for (int i = 0; i < 8; ++i) pp[i] = *reinterpret_cast<const float4*>(ps + 4 * i);
          } else {
            const float* ps = g.pos + static_cast<size_t>(tile) * (128 * 256) + r * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pp[i] = *reinterpret_cast<const float4*>(ps + ((c0 >> 2) + i) * 512);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) pp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = (__uint_as_float(u[j]) + s_b2[c0 + j] - mean) * rstd * s_g2[c0 + j] + s_be2[c0 + j];
        if (inb) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(yblk + ((c0 >> 2) + i) * 512) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (g.out_b) st_global_bf16x32(g.out_b + static_cast<size_t>(row) * 256 + c0, v);
          if (st_pos) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i + 0] += pp[i].x;
              v[4 * i + 1] += pp[i].y;
              v[4 * i + 2] += pp[i].z;
              v[4 * i + 3] += pp[i].w;
            }
            st_global_bf16x32(g.out_pb + static_cast<size_t>(row) * 256 + c0, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(z_empty);
      if (tr) LK_TRACE(1, 22);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int launch_layer(cudaStream_t st, const bf16* att, const bf16* wo, const bf16* w1, const bf16* w2,
                 const LayerArgs& args) {
  if (args.M <= 0) return FVTG_OK;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      LK_SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, two, tw1, tw2;
  FVTG_TRY(make_tmap_bf16(&ta, att, args.M, 256, 256, 128, 64));
  FVTG_TRY(make_tmap_bf16(&two, wo, 256, 256, 256, 128, 64));
  FVTG_TRY(make_tmap_bf16(&tw1, w1, 1024, 256, 256, 128, 64));
  FVTG_TRY(make_tmap_bf16(&tw2, w2, 256, 1024, 1024, 128, 64));
  const int tiles = (args.M + 127) / 128;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(st, PC_LAYER);
  layer_kernel<<<grid, LK_THREADS, LK_SMEM_BYTES, st>>>(ta, two, tw1, tw2, args);
  FVTG_LAUNCH_CHECK("layer_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
