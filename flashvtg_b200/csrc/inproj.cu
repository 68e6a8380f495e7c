// First input-projection layer (LinearLayer 0, FlashVTG/model.py:99-110,782-789) fused with its
// LayerNorm over the RAW feature dim, the ReLU and the LayerNorm(256) that precedes layer 1:
//
//   t = bf16( LN_256( ReLU( LN_D(x) . W^T + b ) ) )                  x: fp32 [rows][D]
//
// in ONE pass over the fp32 features (the only reader of the 0.75 MB/video of inputs), instead of
// a LayerNorm+cast kernel that writes a bf16 copy which a GEMM then re-reads.  LayerNorm is folded
// into the GEMM algebraically:
//
//   LN_D(x) . W^T + b = rstd * ( (x - m0) . Wg^T - ms * colsum(Wg) ) + (W . beta + b)
//       Wg = W * diag(gamma)  (bf16, packed by the host),  ms = mean(x - m0),  rstd = 1/sqrt(var + eps)
//
// where m0 (the mean of the row's first 64 values) is ANY per-row shift: LN is shift invariant, and
// subtracting a cheap estimate of the mean before the bf16 rounding keeps the cancellation in
// "acc - ms * colsum" benign for features with a large common offset.  Row statistics are exact
// fp32 sums over the shifted values, accumulated by the same warps that convert the tile.
//
// 24 warps: 0 TMA producer of the weight tiles, 1 MMA issuer (tcgen05, 128 x 256 x 16), 2 TMEM
// allocator, 4..7 epilogue (one thread per accumulator row), 8..23 converters: warp c owns rows
// 8c..8c+7 of the tile; per 64-column k-block a lane loads 2 columns of each row with 8-byte
// loads (row pitches of the TEF-appended features are only 8-byte aligned, so TMA cannot fetch
// them), keeps the next k-block's loads in flight while it converts the current one, and writes
// bf16 pairs straight into the SWIZZLE_128B A tile.
//
// What bounds it (tools/probe_stream.py, tools/trace_inproj.py, profiles/r01i_inproj_notes.md): the
// k-block shape - 128 rows x 256 B per step, every row another DRAM page - streams at 3.8 TB/s on a
// B200 even with nothing but cp.async copies (512 B per row: 4.7, 1 KB: 5.2, 2 KB: 5.4 TB/s); the
// text projection runs at 3.0 TB/s of that, the short-K video projection is bound by the two-pass
// epilogue of its 4 epilogue warps.  A cp.async-staged variant (3 x 32 KB fp32 slots) and equal
// row ranges per SM were built and measured: no faster, because the limit is the access shape.
#include "gemm.cuh"
#include "kernels.cuh"

namespace fvtg {

constexpr int IP_THREADS = 768;
constexpr int IP_STAGES = 4;
constexpr int IP_A_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int IP_B_BYTES = 256 * 64 * 2;   // 32 KB
constexpr int IP_STAGE_BYTES = IP_A_BYTES + IP_B_BYTES;
constexpr int IP_OFF_BAR = IP_STAGES * IP_STAGE_BYTES;
constexpr int IP_OFF_STAT = IP_OFF_BAR + 256;                 // float2 [2][128]
constexpr int IP_OFF_PAR = IP_OFF_STAT + 2 * 128 * 8;         // wsum, cfold, g1, b1
constexpr int IP_SMEM_BYTES = IP_OFF_PAR + 4 * 256 * 4 + 1024;
static_assert(IP_SMEM_BYTES <= 232448, "inproj kernel shared memory over the 227 KB limit");

struct InprojArgs {
  const float* x;      // fp32 [rows][dim]
  int rows, dim, kbs;  // kbs = dim_pad / 64
  const float* wsum;   // [256] column sums of the bf16 Wg rows
  const float* cfold;  // [256] W . beta + b
  const float* g1;     // LayerNorm(256) of layer 1
  const float* b1;
  bf16* out;           // [rows][256]
};

__global__ void __launch_bounds__(IP_THREADS, 1)
inproj_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ InprojArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + IP_OFF_BAR);
  uint64_t* full = bars;                  // [4] converters (16 warps) + TMA -> MMA
  uint64_t* empty = bars + IP_STAGES;     // [4] MMA -> TMA, converters
  uint64_t* tfull = bars + 8;             // [2] accumulator ready
  uint64_t* tempty = bars + 10;           // [2] accumulator drained (4 epilogue warps)
  uint64_t* sfull = bars + 12;            // [2] row statistics of the tile written (16 warps)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);
  float2* s_stat = reinterpret_cast<float2*>(smem + IP_OFF_STAT);
  float* s_wsum = reinterpret_cast<float*>(smem + IP_OFF_PAR);
  float* s_cf = s_wsum + 256;
  float* s_g1 = s_cf + 256;
  float* s_b1 = s_g1 + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (g.rows + 127) >> 7;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) prefetch_tmap(&tmB);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < IP_STAGES; ++s) {
      mbar_init(&full[s], 17);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
      mbar_init(&sfull[a], 16);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += IP_THREADS) {
    s_wsum[i] = g.wsum[i];
    s_cf[i] = g.cfold[i];
    s_g1[i] = g.g1[i];
    s_b1[i] = g.b1[i];
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ---------------------------------------------------- TMA producer (weights) --
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < g.kbs; ++kb) {
          mbar_wait_sleep(&empty[s], ph ^ 1, 64);
          mbar_expect_tx(&full[s], IP_B_BYTES);
          tma_load_2d(smem + s * IP_STAGE_BYTES + IP_A_BYTES, &tmB, kb * 64, 0, &full[s]);
          if (++s == IP_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer --
    if (lane == 0) {
      int s = 0, it = 0;
      uint32_t ph = 0;
      const uint32_t idesc = umma_idesc_bf16(128, 256);
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * 256;
        for (int kb = 0; kb < g.kbs; ++kb) {
          mbar_wait(&full[s], ph);
          fence_proxy_async_smem();   // the converters' generic-proxy stores -> tensor-core reads
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * IP_STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + IP_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (++s == IP_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 8) {
    // -------------------------------------------------------------- converters --
    const int cw = warp - 8;   // rows 8 cw .. 8 cw + 7 of the tile
    int s = 0, it = 0;
    uint32_t ph = 0;
    const int col_l = 2 * lane;            // this lane's column pair inside a k-block
    const float inv_dim = 1.f / static_cast<float>(g.dim);
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const int row0 = tile * 128 + cw * 8;
      const float* xr[8];
      bool live[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        live[j] = row0 + j < g.rows;
        xr[j] = g.x + static_cast<size_t>(live[j] ? row0 + j : 0) * g.dim + col_l;
      }
      float2 cur[8], nxt[8];
      auto load = [&](float2* dst, int kb) {
        const int col = kb * 64 + col_l;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (live[j] && col + 1 < g.dim) dst[j] = __ldcs(reinterpret_cast<const float2*>(xr[j] + kb * 64));
          else if (live[j] && col < g.dim) dst[j] = make_float2(__ldcs(xr[j] + kb * 64), 0.f);
          else dst[j] = make_float2(0.f, 0.f);
        }
      };
      load(cur, 0);
      // per-row shift: mean of the first k-block's valid columns
      float m0[8], s1[8], s2[8];
      {
        const int n0 = g.dim < 64 ? g.dim : 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          m0[j] = warp_sum(cur[j].x + cur[j].y) / static_cast<float>(n0);
          s1[j] = 0.f;
          s2[j] = 0.f;
        }
      }
      for (int kb = 0; kb < g.kbs; ++kb) {
        if (kb + 1 < g.kbs) load(nxt, kb + 1);
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * IP_STAGE_BYTES;
        const int col = kb * 64 + col_l;
        const bool c0 = col < g.dim, c1 = col + 1 < g.dim;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = c0 ? cur[j].x - m0[j] : 0.f;
          const float b = c1 ? cur[j].y - m0[j] : 0.f;
          s1[j] += a + b;
          s2[j] += a * a + b * b;
          const int r = cw * 8 + j;
          *reinterpret_cast<uint32_t*>(sa + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4) =
              pack_bf16(a, b);
        }
        // no proxy fence here (it is a MEMBAR.ALL.CTA that would wait for the loads in flight): the
        // arrive releases the stores to the MMA thread, which fences generic -> async proxy itself
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (++s == IP_STAGES) { s = 0; ph ^= 1; }
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
      }
      // row statistics of the shifted values -> (mean, rstd) for the epilogue
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);   // the epilogue is done with this slot's previous tile
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t1 = warp_sum(s1[j]), t2 = warp_sum(s2[j]);
        if (lane == 0) {
          const float ms = t1 * inv_dim;
          const float var = fmaxf(t2 * inv_dim - ms * ms, 0.f);
          s_stat[acc * 128 + cw * 8 + j] = make_float2(ms, rsqrtf(var + 1e-5f));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[acc]);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue --
    const int wq = warp - 4;               // TMEM lane quadrant == warp % 4
    const int r = wq * 32 + lane;
    int it = 0;
    uint32_t u[16];
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      const int row = tile * 128 + r;
      mbar_wait_sleep(&tfull[acc], accph, 256);   // long by design: do not poll the issue slots away
      mbar_wait_sleep(&sfull[acc], accph, 64);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * 256 + (static_cast<uint32_t>(wq * 32) << 16);
      const float2 st = s_stat[acc * 128 + r];
      const float ms = st.x, rstd = st.y;
      // pass 1: y = relu(rstd * (acc - ms * wsum) + cfold); LayerNorm(256) statistics; y back into TMEM.
      // Packed fp32x2 arithmetic: the four epilogue warps are the critical path of short-K tiles.
      const uint64_t rstd2 = f2_pack(rstd, rstd), k2 = f2_pack(-rstd * ms, -rstd * ms);
      uint64_t nshift2 = 0ull, t1 = 0ull, t2 = 0ull;
#pragma unroll 1
      for (int c = 0; c < 16; ++c) {
        const int c0 = c * 16;
        tmem_ld16(tacc + c0, u);
        tmem_ld_wait();
        const uint64_t* ws2 = reinterpret_cast<const uint64_t*>(s_wsum + c0);
        const uint64_t* cf2 = reinterpret_cast<const uint64_t*>(s_cf + c0);
        if (c == 0) {
          const float y0 = fmaxf(fmaf(__uint_as_float(u[0]), rstd, fmaf(-rstd * ms, s_wsum[0], s_cf[0])), 0.f);
          nshift2 = f2_pack(-y0, -y0);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t y = f2_fma(f2_pack(__uint_as_float(u[2 * j]), __uint_as_float(u[2 * j + 1])), rstd2,
                                    f2_fma(k2, ws2[j], cf2[j]));
          float ya, yb;
          f2_unpack(y, ya, yb);
          ya = fmaxf(ya, 0.f);
          yb = fmaxf(yb, 0.f);
          const uint64_t d = f2_add(f2_pack(ya, yb), nshift2);
          t1 = f2_add(t1, d);
          t2 = f2_fma(d, d, t2);
          u[2 * j] = __float_as_uint(ya);
          u[2 * j + 1] = __float_as_uint(yb);
        }
        tmem_st16(tacc + c0, u);
      }
      tmem_st_wait();
      float t1a, t1b, t2a, t2b, nsh, nsh_;
      f2_unpack(t1, t1a, t1b);
      f2_unpack(t2, t2a, t2b);
      f2_unpack(nshift2, nsh, nsh_);
      const float s1 = t1a + t1b, s2 = t2a + t2b;
      const float mean = s1 * (1.f / 256.f) - nsh;
      const float var = fmaxf((s2 - s1 * s1 * (1.f / 256.f)) * (1.f / 256.f), 0.f);
      const float rs = rsqrtf(var + 1e-5f);
      const uint64_t rs2 = f2_pack(rs, rs), nmean2 = f2_pack(-mean, -mean);
      // pass 2: normalise -> bf16 -> 32-byte row stores
#pragma unroll 1
      for (int c = 0; c < 16; ++c) {
        const int c0 = c * 16;
        tmem_ld16(tacc + c0, u);
        tmem_ld_wait();
        const uint64_t* g2 = reinterpret_cast<const uint64_t*>(s_g1 + c0);
        const uint64_t* b2 = reinterpret_cast<const uint64_t*>(s_b1 + c0);
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t a2 = f2_mul(rs2, g2[j]);
          const uint64_t o = f2_fma(f2_pack(__uint_as_float(u[2 * j]), __uint_as_float(u[2 * j + 1])), a2,
                                    f2_fma(nmean2, a2, b2[j]));
          float oa, ob;
          f2_unpack(o, oa, ob);
          w[j] = pack_bf16(oa, ob);
        }
        if (row < g.rows) st_global_v8(g.out + static_cast<size_t>(row) * 256 + c0, w);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// x fp32 [rows][dim] -> out bf16 [rows][256]; wg bf16 [256][dim_pad] (gamma folded, zero padded).
int launch_inproj(cudaStream_t st, const float* x, int rows, int dim, int dim_pad, const void* wg,
                  const float* wsum, const float* cfold, const float* g1, const float* b1, bf16* out) {
  if (rows <= 0) return FVTG_OK;
  if (dim < 2 || (dim & 1) || dim_pad % 64 || dim_pad < dim || (reinterpret_cast<uintptr_t>(x) & 7))
    return fail(FVTG_EINVAL, "inproj: feature dim %d must be even, rows 8-byte aligned", dim);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(inproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      IP_SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tb;
  FVTG_TRY(make_tmap_bf16(&tb, wg, 256, dim_pad, dim_pad, 256, 64));
  InprojArgs a;
  a.x = x; a.rows = rows; a.dim = dim; a.kbs = dim_pad / 64;
  a.wsum = wsum; a.cfold = cfold; a.g1 = g1; a.b1 = b1; a.out = out;
  const int m_tiles = (rows + 127) / 128;
  const int grid = m_tiles < sm_count() ? m_tiles : sm_count();
  ProfScope prof(st, PC_LNCAST);
  FVTG_CUDA_OK(launch_pdl(inproj_kernel, dim3(grid), dim3(IP_THREADS), IP_SMEM_BYTES, st, tb, a));
  FVTG_LAUNCH_CHECK("inproj_kernel");
  return FVTG_OK;
}

}  // namespace fvtg

// debug / tuning entry: the fused first projection alone (tools/trace_inproj.py)
#ifdef FVTG_DEBUG_HOOKS   // test / tuning hook: only in libflashvtg_b200_dbg.so
extern "C" int32_t fvtg_dbg_inproj(const float* x, int32_t rows, int32_t dim, int32_t dim_pad, const void* wg,
                                   const float* wsum, const float* cfold, const float* g1, const float* b1,
                                   void* out, void* stream) {
  return fvtg::launch_inproj(static_cast<cudaStream_t>(stream), x, rows, dim, dim_pad, wg, wsum, cfold, g1, b1,
                             static_cast<fvtg::bf16*>(out));
}
#endif
