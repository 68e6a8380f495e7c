// Fused transformer-layer tail: everything of a post-norm layer that is row-wise, in ONE
// persistent tcgen05 kernel per 128-row tile, with the 1024-wide hidden activations never leaving
// the SM (they never even reach shared memory):
//
//   z  = x + att . Wo^T + bo                         (out_proj + residual)
//   T2V layer (transformer.py:359-367):  y = LN2( z  + W2 . PReLU(W1 . LN1(z) + b1) + b2 )
//   SA  layer (transformer.py:416-420):  y = LN2( x1 + W2 . PReLU(W1 . x1 + b1) + b2 ),  x1 = LN1(z)
//
// TMEM (512 columns): Z = columns 0..255 (FFN accumulator, initialised with the residual),
//                     H = columns 256..511 = two 128-column hidden-piece accumulators H0 | H1;
//                     the out_proj of a tile also accumulates in H (free between tiles), so it
//                     overlaps the previous tile's final epilogue, which drains Z.
// Shared memory: sA = 4 x [128 rows][64 k] bf16 SWIZZLE_128B units (att tile, then LN1 output),
//                a 4-stage ring of 32 KB weight stages, LayerNorm statistics, per-layer vectors.
//
// Per tile:
//   TMA : att tile -> sA ; weight stages: Wo[256 n][64 k] x4, then per hidden piece p (128 wide)
//         W1[p] as 2 stages of 2 x [128 n][64 k], W2[:, p] as 2 stages of [256 n][64 k]
//   MMA : H(0..255)  = sA . Wo^T                        16 x (128 x 256 x 16), operands in smem
//   EPI1: z = H + bo + x (fp32 residual stream, tile-blocked layout => coalesced 16 B / lane);
//         LayerNorm1 -> bf16 -> sA (written swizzled by the epilogue threads);
//         Z <- z (T2V) or LN1(z) (SA) via tcgen05.st: the FFN residual lives in the accumulator
//   for the 8 hidden pieces p:     (ff1(p+1), ff1(p+2) are issued before ff2(p): MMA rarely waits)
//         MMA : H[p&1] = sA . W1[p]^T                    16 x (128 x 128 x 16)
//         EPI2: h = PReLU(H[p&1] + b1[p]) -> bf16 pairs -> tcgen05.st back into the SAME columns
//               (each thread overwrites only columns it alone has read)
//         MMA : Z += h . W2[:, p]^T                      8 x (128 x 256 x 16), A operand from TMEM
//   EPI3: LayerNorm2(Z + b2) in two TMEM passes.  Pass 1: row statistics + the fp32 row of the residual stream -
//         the RAW sums plus (rstd, -mean * rstd) per row when the next reader is another layer kernel (it
//         normalises inside its EPI1: stats_in / pg / pbe), so the statistics pass carries half of the tile's
//         stores.  Pass 2: bf16(y), bf16(y + pos) as paired 64-byte row segments (and y itself for the last layer
//         of a stream).  The phase is bound by the SM's ~30 B/cycle store port (tools/probe_store.py).
//
// 20 warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..19 epilogue (four threads per row: TMEM lane quadrant = warp % 4, column
// quarter = (warp - 4) / 4).
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int LK_THREADS = 640;
constexpr int LK_STAGES = 4;
constexpr int LK_UNIT = 128 * 64 * 2;        // 16 KB: [128 rows][64 bf16], SWIZZLE_128B
constexpr int LK_STAGE = 2 * LK_UNIT;        // 32 KB
constexpr int LK_OFF_A = 0;                  // 4 units: att tile, then LN1 output
constexpr int LK_OFF_W = 4 * LK_UNIT;        // weight ring
constexpr int LK_OFF_BAR = LK_OFF_W + LK_STAGES * LK_STAGE;
constexpr int LK_OFF_STAT = LK_OFF_BAR + 256;
constexpr int LK_OFF_PAR = LK_OFF_STAT + 2 * 4 * 128 * 8;   // [2 phases][4 quarters][128 rows] float2
constexpr int LK_PAR_FLOATS = 256 * 3 + 1024 + 256 * 4;
constexpr int LK_SMEM_BYTES = LK_OFF_PAR + LK_PAR_FLOATS * 4 + 1024 /*align slack*/;
static_assert(LK_SMEM_BYTES <= 232448, "layer kernel shared memory over the 227 KB limit");

// trace slots: [role 0 MMA / 1 epilogue][tile it < 8][event < 32]
#define LK_TRACE(role, ev)                                                              \
  do {                                                                                  \
    if (g.trace && static_cast<int>(blockIdx.x) == g.trace_cta && it < 8)                                           \
      g.trace[((role) * 8 + it) * 32 + (ev)] = clock64();                               \
  } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// two packed fp32 pairs (16 bytes) from shared / global memory
__device__ __forceinline__ ulonglong2 ld_p4(const float* p) { return *reinterpret_cast<const ulonglong2*>(p); }
__device__ __forceinline__ uint64_t pk2(const uint32_t* u, int j) {
  return f2_pack(__uint_as_float(u[2 * j]), __uint_as_float(u[2 * j + 1]));
}
__device__ __forceinline__ void unpk2(uint64_t v, uint32_t* u, int j) {
  float x, y;
  f2_unpack(v, x, y);
  u[2 * j] = __float_as_uint(x);
  u[2 * j + 1] = __float_as_uint(y);
}
__device__ __forceinline__ uint32_t bf16x2_of(uint64_t v) {
  float x, y;
  f2_unpack(v, x, y);
  return pack_bf16(x, y);
}

__global__ void __launch_bounds__(LK_THREADS, 1)
layer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWo,
             const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
             const __grid_constant__ LayerArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ symbol (an integer round trip
  // would demote every later access to a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + LK_OFF_A;
  uint8_t* sW = smem + LK_OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LK_OFF_BAR);
  uint64_t* full = bars;                    // [4] TMA -> MMA
  uint64_t* empty = bars + LK_STAGES;       // [4] MMA -> TMA
  uint64_t* a_full = bars + 8;              // att tile landed in sA
  uint64_t* a_empty = bars + 9;             // last ff1 MMA retired: sA reusable
  uint64_t* z1_full = bars + 10;            // out_proj accumulated in H
  uint64_t* z2_full = bars + 11;            // FFN accumulated in Z
  uint64_t* ln_ready = bars + 12;           // LN1 tile in sA, residual in Z, H drained
  uint64_t* hacc_full = bars + 13;          // [2] ff1 piece accumulated
  uint64_t* h_ready = bars + 15;            // [2] bf16 hidden piece stored back into H[buf]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 20);
  float2* s_stat = reinterpret_cast<float2*>(smem + LK_OFF_STAT);  // [2 phases][4 quarters][128]
  float* s_par = reinterpret_cast<float*>(smem + LK_OFF_PAR);
  float* s_bo = s_par;            // bo (+ pbe of the producer's deferred LayerNorm-2)
  float* s_g1 = s_par + 256;
  float* s_be1 = s_par + 512;
  float* s_b1 = s_par + 768;
  float* s_pg = s_par + 1792;     // gamma of the producer's deferred LayerNorm-2 (1 when the residual is final)
  float* s_g2 = s_par + 2048;
  float* s_be2 = s_par + 2304;
  float* s_b2z = s_par + 2560;    // T2V: b2 ; SA: be1 + b2 (the FFN residual LN1(z) enters Z with ff2's bias folded in)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = (g.M + g.tile_rows - 1) / g.tile_rows;

  pdl_launch_dependents();
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
    mbar_init(ln_ready, 16);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hacc_full[b], 1);
      mbar_init(&h_ready[b], 16);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_holder, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 256; i += LK_THREADS) {
    s_bo[i] = g.bo[i] + (g.stats_in ? g.pbe[i] : 0.f);
    s_pg[i] = g.stats_in ? g.pg[i] : 1.f;
    s_g1[i] = g.g1[i];
    s_be1[i] = g.be1[i];
    s_g2[i] = g.g2[i];
    s_be2[i] = g.be2[i];
    s_b2z[i] = g.mode == LAYER_SA ? g.be1[i] + g.b2[i] : g.b2[i];
  }
  for (int i = threadIdx.x; i < 1024; i += LK_THREADS) s_b1[i] = g.b1[i];
  pdl_wait();   // everything above touched only constants / on-chip state
  if (g.trace && static_cast<int>(blockIdx.x) == g.trace_cta && threadIdx.x == 0) {
    long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g.trace[512] = ns;
    g.trace[513] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tmem_h = tmem + 256;
  if (g.stagger_ns > 0) {
    const int my_tiles = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int max_tiles = (ntiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    if (my_tiles < max_tiles) {
      const long long t_end = clock64() + static_cast<long long>(blockIdx.x % 8) * g.stagger_ns * 2;
      while (clock64() < t_end) __nanosleep(200);
    }
  }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer --
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto stage_wait = [&]() -> uint8_t* {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], LK_STAGE);
        return sW + s * LK_STAGE;
      };
      auto stage_next = [&]() {
        if (++s == LK_STAGES) { s = 0; ph ^= 1; }
      };
      auto ff1 = [&](int p) {  // W1 rows p*128.., k-blocks (2h, 2h+1) per stage
        for (int h = 0; h < 2; ++h) {
          uint8_t* d = stage_wait();
          tma_load_2d(d, &tmW1, (2 * h) * 64, p * 128, &full[s]);
          tma_load_2d(d + LK_UNIT, &tmW1, (2 * h + 1) * 64, p * 128, &full[s]);
          stage_next();
        }
      };
      auto ff2 = [&](int p) {  // W2 all 256 rows, k-block p*128 + h*64
        for (int h = 0; h < 2; ++h) {
          uint8_t* d = stage_wait();
          tma_load_2d(d, &tmW2, p * 128 + h * 64, 0, &full[s]);
          stage_next();
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        mbar_wait(a_empty, (it & 1) ^ 1);
        mbar_expect_tx(a_full, 4 * LK_UNIT);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sA + kb * LK_UNIT, &tmA, kb * 64, tile * g.tile_rows, a_full);
        for (int kb = 0; kb < 4; ++kb) {
          uint8_t* d = stage_wait();
          tma_load_2d(d, &tmWo, kb * 64, 0, &full[s]);
          stage_next();
        }
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
      const uint32_t idesc128 = umma_idesc_bf16(128, 128);
      const uint32_t idesc256 = umma_idesc_bf16(128, 256);
      const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
      auto stage_wait = [&]() -> uint32_t {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        return sW_u + s * LK_STAGE;
      };
      auto stage_release = [&]() {
        umma_commit(&empty[s]);
        if (++s == LK_STAGES) { s = 0; ph ^= 1; }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        auto ff1 = [&](int p) {
          const int buf = p & 1;
          const uint32_t d = tmem_h + buf * 128;
          for (int h = 0; h < 2; ++h) {
            const uint32_t w = stage_wait();
#pragma unroll
            for (int kbl = 0; kbl < 2; ++kbl) {
              const uint64_t da = umma_desc_sw128(sA_u + (2 * h + kbl) * LK_UNIT);
              const uint64_t db = umma_desc_sw128(w + kbl * LK_UNIT);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d, da + 2 * k, db + 2 * k, idesc128, (h | kbl | k) ? 1u : 0u);
            }
            stage_release();
          }
          umma_commit(&hacc_full[buf]);
          if (p == 7) umma_commit(a_empty);
        };
        auto ff2 = [&](int p) {
          const int buf = p & 1;
          const uint32_t n = static_cast<uint32_t>(it) * 4u + static_cast<uint32_t>(p >> 1);
          mbar_wait(&h_ready[buf], n & 1u);
          tc_fence_after();
          const uint32_t a = tmem_h + buf * 128;
          for (int h = 0; h < 2; ++h) {
            const uint32_t w = stage_wait();
            const uint64_t db = umma_desc_sw128(w);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = h * 4 + k;  // K = 16 step of the piece: 8 packed columns of its thread quarter
              umma_bf16_ts(tmem, a + 32 * (j >> 1) + 8 * (j & 1), db + 2 * k, idesc256, 1u);
            }
            stage_release();
          }
        };
        LK_TRACE(0, 0);
        mbar_wait(a_full, it & 1);
        tc_fence_after();
        LK_TRACE(0, 2);
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t w = stage_wait();
          const uint64_t da = umma_desc_sw128(sA_u + kb * LK_UNIT);
          const uint64_t db = umma_desc_sw128(w);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_h, da + 2 * k, db + 2 * k, idesc256, (kb | k) ? 1u : 0u);
          stage_release();
        }
        umma_commit(z1_full);
        LK_TRACE(0, 3);
        mbar_wait(ln_ready, it & 1);
        tc_fence_after();
        LK_TRACE(0, 4);
        // the fp32 residual block of this CTA's next tile (128 KB, contiguous in the tile-blocked layout) was
        // written a whole layer ago: pull it back into L2 now, a full FFN ahead of the epilogue that reads it
        if (g.tile_rows == 128 && tile + static_cast<int>(gridDim.x) < ntiles && !(g.dbg & 8)) {
          const float* nxt = g.yf + static_cast<size_t>(tile + gridDim.x) * (128 * 256);
#pragma unroll
          for (int i = 0; i < 8; ++i) bulk_prefetch_l2(nxt + i * 4096, 16384);
        }
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
    const int qt = ew >> 2;  // column quarter
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const bool tr = (warp == 4 && lane == 0);
    uint32_t u[16];
    int it = 0;

    // The residual stream may arrive un-normalised: when the producing layer kernel deferred its LayerNorm-2 to
    // its consumer (stats_in != null), yf holds the raw sums and (rstd, -mean * rstd) per row; x = (raw * rstd
    // - mean * rstd) * pg + pbe is applied here (pbe is folded into s_bo).  Otherwise (1, 0) with pg = 1.
    ulonglong2 rr[4];
    float r_rs = 1.f, r_nm = 0.f;
    // first residual chunk (16 columns) of a tile: it depends on no MMA, so it is fetched a phase early - before
    // the first tile, and for every later tile before the previous tile's drain puts 128 KB of stores in front of
    // it; the other three chunks follow one chunk ahead of their use inside epilogue 1
    auto load_res0 = [&](int t) {
      const int rw_ = t * g.tile_rows + r;
      const bool ok = t < ntiles && r < g.tile_rows && rw_ < g.M && !(g.dbg & 1);
      const float* src = g.yf + static_cast<size_t>(rw_ >> 7) * (128 * 256) + (rw_ & 127) * 4 +
                         static_cast<size_t>(qt) * (16 * 512);
#pragma unroll
      for (int i = 0; i < 4; ++i) rr[i] = ok ? ld_p4(src + i * 512) : make_ulonglong2(0ull, 0ull);
      r_rs = 1.f;
      r_nm = 0.f;
      if (ok && g.stats_in) {
        const float2 st = g.stats_in[rw_];
        r_rs = st.x;
        r_nm = st.y;
      }
    };

    load_res0(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int row = tile * g.tile_rows + r;
      const bool inb = r < g.tile_rows && row < g.M;
      const bool ldres = inb && !(g.dbg & 1);
      // fp32 residual stream, blocked by 128 rows: [row/128][col/4][row%128][4]
      const size_t blk = static_cast<size_t>(row >> 7) * (128 * 256) + (row & 127) * 4;
      float* yblk = g.yf + blk;

      // ---- epilogue 1: z = H + bo + x ; LayerNorm1 -> sA ; FFN residual into Z -----------------
      // (packed fp32 pairs throughout: add.f32x2 / fma.rn.f32x2 halve the FP instruction count)
      if (tr) LK_TRACE(1, 0);
      const float* ysrc = yblk + static_cast<size_t>(qt) * (16 * 512);   // this thread's column quarter
      mbar_wait(z1_full, it & 1);
      tc_fence_after();
      if (tr) LK_TRACE(1, 1);
      float shift = 0.f;
      uint64_t nsh2 = 0ull, a1 = 0ull, a2 = 0ull;
      {
        const uint64_t rrs2 = f2_pack(r_rs, r_rs), rnm2 = f2_pack(r_nm, r_nm);
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 16 columns at a time: half the registers of a 32-column pass
          const int c0 = qt * 64 + c * 16;
          tmem_ld16(tmem_h + lane_addr + c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const ulonglong2 b = ld_p4(s_bo + c0 + 4 * i), pg4 = ld_p4(s_pg + c0 + 4 * i);
            unpk2(f2_add(pk2(u, 2 * i), f2_fma(f2_fma(rr[i].x, rrs2, rnm2), pg4.x, b.x)), u, 2 * i);
            unpk2(f2_add(pk2(u, 2 * i + 1), f2_fma(f2_fma(rr[i].y, rrs2, rnm2), pg4.y, b.y)), u, 2 * i + 1);
            if (c < 3) rr[i] = ldres ? ld_p4(ysrc + ((c + 1) * 4 + i) * 512) : make_ulonglong2(0ull, 0ull);
          }
          if (c == 0) {
            shift = __uint_as_float(u[0]);
            nsh2 = f2_pack(-shift, -shift);
          }
          tmem_st16(tmem + lane_addr + c0, u);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t d = f2_add(pk2(u, j), nsh2);
            a1 = f2_add(a1, d);
            a2 = f2_fma(d, d, a2);
          }
        }
      }
      {
        float s1a, s1b, s2a, s2b;
        f2_unpack(a1, s1a, s1b);
        f2_unpack(a2, s2a, s2b);
        const float s1 = s1a + s1b, s2 = s2a + s2b;
        // per-quarter partial: (mean of 64, centred sum of squares of 64)
        s_stat[qt * 128 + r] = make_float2(shift + s1 * (1.f / 64.f), s2 - s1 * s1 * (1.f / 64.f));
      }
      tmem_st_wait();
      if (tr) LK_TRACE(1, 2);
      epi_bar_sync();
      // every epilogue thread has issued the previous tile's last global stores before this barrier: publish it
      if (g.tile_flags && it > 0 && warp == 4 && lane == 0)
        st_release_gpu(g.tile_flags + (tile - static_cast<int>(gridDim.x)), g.flag_epoch);
      float mean, rstd;
      {
        const float2 a0 = s_stat[r], a1 = s_stat[128 + r], a2 = s_stat[256 + r], a3 = s_stat[384 + r];
        mean = 0.25f * (a0.x + a1.x + a2.x + a3.x);
        const float d0 = a0.x - mean, d1 = a1.x - mean, d2 = a2.x - mean, d3 = a3.x - mean;
        const float m2 = a0.y + a1.y + a2.y + a3.y + 64.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
        rstd = rsqrtf(fmaxf(m2 * (1.f / 256.f), 0.f) + 1e-5f);
      }
      {
        const uint64_t rs2 = f2_pack(rstd, rstd), nm2 = f2_pack(-mean * rstd, -mean * rstd);
        uint8_t* unit = sA + qt * LK_UNIT;   // this thread's column quarter == one 64-column SWIZZLE_128B unit
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int c0 = qt * 64 + c * 16;
          tmem_ld16(tmem + lane_addr + c0, u);
          tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const ulonglong2 gm = ld_p4(s_g1 + c0 + 4 * i), bt = ld_p4(s_be1 + c0 + 4 * i);
            const uint64_t t0 = f2_fma(pk2(u, 2 * i), rs2, nm2), t1 = f2_fma(pk2(u, 2 * i + 1), rs2, nm2);
            w[2 * i] = bf16x2_of(f2_fma(t0, gm.x, bt.x));
            w[2 * i + 1] = bf16x2_of(f2_fma(t1, gm.y, bt.y));
            // s_b2z = b2 (T2V: the FFN residual is the pre-LN1 sum z) or be1 + b2 (SA: it is LN1(z))
            const ulonglong2 bz = ld_p4(s_b2z + c0 + 4 * i);
            if (g.mode == LAYER_SA) {
              unpk2(f2_fma(t0, gm.x, bz.x), u, 2 * i);
              unpk2(f2_fma(t1, gm.y, bz.y), u, 2 * i + 1);
            } else {
              unpk2(f2_add(pk2(u, 2 * i), bz.x), u, 2 * i);
              unpk2(f2_add(pk2(u, 2 * i + 1), bz.y), u, 2 * i + 1);
            }
          }
          tmem_st16(tmem + lane_addr + c0, u);
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4)
            *reinterpret_cast<uint4*>(unit + sw128_off(r, c * 2 + q4)) =
                make_uint4(w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ln_ready);
      if (tr) LK_TRACE(1, 3);

      // ---- epilogue 2 (x8): hidden piece = PReLU(H[buf] + b1) -> bf16 pairs back into H[buf] (two 16-column
      //      halves) ----------------------------------------------------------------------------------------
      const uint64_t al2 = f2_pack(g.prelu, g.prelu);
      const bool prelu_max = g.prelu >= 0.f && g.prelu <= 1.f;   // PReLU(x) = max(x, a x): no compare / select
#pragma unroll 1
      for (int p = 0; p < 8; ++p) {
        const int buf = p & 1;
        const uint32_t n = static_cast<uint32_t>(it) * 4u + static_cast<uint32_t>(p >> 1);
        mbar_wait(&hacc_full[buf], n & 1u);
        tc_fence_after();
        if (tr) LK_TRACE(1, 4 + 2 * p);
        const uint32_t ha = tmem_h + lane_addr + buf * 128 + qt * 32;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld16(ha + hh * 16, u);
          tmem_ld_wait();
          const float* bb = s_b1 + p * 128 + qt * 32 + hh * 16;
          uint32_t hp[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const ulonglong2 b = ld_p4(bb + 4 * i);
            const uint64_t x0 = f2_add(pk2(u, 2 * i), b.x), x1 = f2_add(pk2(u, 2 * i + 1), b.y);
            float a0, a1, a2, a3;
            f2_unpack(x0, a0, a1);
            f2_unpack(x1, a2, a3);
            if (prelu_max) {
              float m0, m1, m2, m3;
              f2_unpack(f2_mul(x0, al2), m0, m1);
              f2_unpack(f2_mul(x1, al2), m2, m3);
              hp[2 * i + 0] = pack_bf16(fmaxf(a0, m0), fmaxf(a1, m1));
              hp[2 * i + 1] = pack_bf16(fmaxf(a2, m2), fmaxf(a3, m3));
            } else {
              a0 = a0 > 0.f ? a0 : g.prelu * a0;
              a1 = a1 > 0.f ? a1 : g.prelu * a1;
              a2 = a2 > 0.f ? a2 : g.prelu * a2;
              a3 = a3 > 0.f ? a3 : g.prelu * a3;
              hp[2 * i + 0] = pack_bf16(a0, a1);
              hp[2 * i + 1] = pack_bf16(a2, a3);
            }
          }
          tmem_st8(ha + hh * 8, hp);   // packed pairs land in columns this thread alone has already read
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[buf]);
        if (tr) LK_TRACE(1, 5 + 2 * p);
      }

      // ---- final epilogue: LayerNorm2(Z) (b2 is already in Z).  An SM stores only ~30 B/cycle (tools/probe_store.py),
      //      so the 192-256 KB a tile writes bound this phase, not arithmetic: the statistics pass, which used to
      //      store nothing, now carries the fp32 row (the RAW sums plus (rstd, -mean * rstd) per row when the next
      //      reader is another layer kernel, which normalises while it reads its residual), and the second pass only
      //      the bf16 operands, as paired 64-byte row segments (full store-port rate instead of half) ------------
      int prow = row;
      if (g.pos_mod > 0) prow = row % g.pos_mod;
      const bool st_ok = inb && !(g.dbg & 4);
      const bool st_pb = g.out_pb && st_ok && (g.pos_rowlim <= 0 || prow < g.pos_rowlim);
      const bool ld_pos = g.out_pb && g.pos && inb && (g.pos_rowlim <= 0 || prow < g.pos_rowlim) && !(g.dbg & 2);
      load_res0(tile + static_cast<int>(gridDim.x));
      mbar_wait(z2_full, it & 1);
      tc_fence_after();
      if (tr) LK_TRACE(1, 20);
      a1 = 0ull;
      a2 = 0ull;
      float* ydst = yblk + static_cast<size_t>(qt) * (16 * 512);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld16(tmem + lane_addr + qt * 64 + c * 16, u);
        tmem_ld_wait();
        if (c == 0) {
          shift = __uint_as_float(u[0]);
          nsh2 = f2_pack(-shift, -shift);
        }
        if (st_ok && g.stats_out) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(ydst + (c * 4 + i) * 512) =
                make_uint4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t d = f2_add(pk2(u, j), nsh2);
          a1 = f2_add(a1, d);
          a2 = f2_fma(d, d, a2);
        }
      }
      {
        float s1a, s1b, s2a, s2b;
        f2_unpack(a1, s1a, s1b);
        f2_unpack(a2, s2a, s2b);
        const float s1 = s1a + s1b, s2 = s2a + s2b;
        s_stat[512 + qt * 128 + r] = make_float2(shift + s1 * (1.f / 64.f), s2 - s1 * s1 * (1.f / 64.f));
      }
      epi_bar_sync();
      if (tr) LK_TRACE(1, 21);
      {
        const float2 a0 = s_stat[512 + r], a1 = s_stat[640 + r], a2 = s_stat[768 + r], a3 = s_stat[896 + r];
        mean = 0.25f * (a0.x + a1.x + a2.x + a3.x);
        const float d0 = a0.x - mean, d1 = a1.x - mean, d2 = a2.x - mean, d3 = a3.x - mean;
        const float m2 = a0.y + a1.y + a2.y + a3.y + 64.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
        rstd = rsqrtf(fmaxf(m2 * (1.f / 256.f), 0.f) + 1e-5f);
      }
      if (g.stats_out && qt == 0 && st_ok) g.stats_out[row] = make_float2(rstd, -mean * rstd);
      if (g.out_b || g.out_pb || !g.stats_out) {
        const uint64_t rs2 = f2_pack(rstd, rstd), nm2 = f2_pack(-mean * rstd, -mean * rstd);
        const float* psrc = nullptr;   // this lane's position row, first column of its quarter
        size_t pstep = 0;              // floats between consecutive 4-column groups
        if (ld_pos) {
          if (g.pos_mod > 0) { psrc = g.pos + static_cast<size_t>(prow) * 256 + qt * 64; pstep = 4; }
          else if (g.pos_cmp_L > 0) {
            psrc = g.pos + (static_cast<size_t>(qt * 16) * g.pos_cmp_L + row % g.pos_cmp_L) * 4;
            pstep = static_cast<size_t>(g.pos_cmp_L) * 4;
          } else { psrc = g.pos + blk + static_cast<size_t>(qt) * (16 * 512); pstep = 512; }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {   // 32 columns: 64 bytes of each bf16 output row
          uint32_t v[32];
          tmem_ld32(tmem + lane_addr + qt * 64 + c * 32, v);
          tmem_ld_wait();
          const int c0 = qt * 64 + c * 32;
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const ulonglong2 gm = ld_p4(s_g2 + c0 + 4 * i), bt = ld_p4(s_be2 + c0 + 4 * i);
            const uint64_t y0 = f2_fma(f2_fma(pk2(v, 2 * i), rs2, nm2), gm.x, bt.x);
            const uint64_t y1 = f2_fma(f2_fma(pk2(v, 2 * i + 1), rs2, nm2), gm.y, bt.y);
            unpk2(y0, v, 2 * i);
            unpk2(y1, v, 2 * i + 1);
            w[2 * i] = bf16x2_of(y0);
            w[2 * i + 1] = bf16x2_of(y1);
          }
          if (!g.stats_out && st_ok) {   // no consumer kernel will normalise: y itself goes to the residual buffer
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<uint4*>(ydst + (c * 8 + i) * 512) =
                  make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (g.out_b) st_global_bf16x32_paired_w(g.out_b + static_cast<size_t>(row) * 256 + c0, 256, st_ok, w);
          if (g.out_pb) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const ulonglong2 pv = psrc ? ld_p4(psrc + (c * 8 + i) * pstep) : make_ulonglong2(0ull, 0ull);
              w[2 * i] = bf16x2_of(f2_add(pk2(v, 2 * i), pv.x));
              w[2 * i + 1] = bf16x2_of(f2_add(pk2(v, 2 * i + 1), pv.y));
            }
            st_global_bf16x32_paired_w(g.out_pb + static_cast<size_t>(row) * 256 + c0, 256, st_pb, w);
          }
        }
      }
      // Z is rewritten next by this same thread (epilogue 1 of the next tile), H by MMAs that
      // are ordered behind ln_ready: no further hand-off is needed here.
      tc_fence_before();
      if (tr) LK_TRACE(1, 22);
    }
    if (g.tile_flags && it > 0) {   // the CTA's last tile
      epi_bar_sync();
      if (warp == 4 && lane == 0)
        st_release_gpu(g.tile_flags + (static_cast<int>(blockIdx.x) + (it - 1) * static_cast<int>(gridDim.x)), g.flag_epoch);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (g.trace && static_cast<int>(blockIdx.x) == g.trace_cta && threadIdx.x == 0) {
    long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g.trace[514] = ns;
    g.trace[515] = clock64();
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Tiles advance by tile_rows rows (default 128).  FVTG_LAYER_TILE_ROWS=0 picks the size that gives every
// SM the same number of tiles (76 800 rows on 148 SMs: 5 x 104 rows instead of 4.05 -> 5 rounds of 128).
// Measured: no gain (1.656 vs 1.638 ms/step for the 11 launches) - a tile's cost does not shrink with its
// live rows (M = 128 MMAs, epilogue threads of dead rows still walk the phases), so it stays off.
int layer_tile_rows(int M) {
  static const int forced = [] { const char* e = getenv("FVTG_LAYER_TILE_ROWS"); return e ? atoi(e) : 128; }();
  int tr = forced;
  if (forced == 0) {
    const int per_sm = (M + sm_count() * 128 - 1) / (sm_count() * 128);
    tr = ((M + sm_count() * per_sm - 1) / (sm_count() * per_sm) + 7) & ~7;
  }
  return tr < 8 ? 8 : (tr > 128 ? 128 : (tr & ~7));
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
  FVTG_TRY(make_tmap_bf16(&two, wo, 256, 256, 256, 256, 64));
  FVTG_TRY(make_tmap_bf16(&tw1, w1, 1024, 256, 256, 128, 64));
  FVTG_TRY(make_tmap_bf16(&tw2, w2, 256, 1024, 1024, 256, 64));
  LayerArgs a2 = args;
  a2.tile_rows = layer_tile_rows(args.M);
  const int tiles = (args.M + a2.tile_rows - 1) / a2.tile_rows;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  {
    const char* d = getenv("FVTG_LAYER_DBG");
    a2.dbg = d ? atoi(d) : 0;
    const char* tc = getenv("FVTG_TRACE_CTA");
    a2.trace_cta = tc ? atoi(tc) : 0;
    const char* sg = getenv("FVTG_LAYER_STAGGER_NS");
    a2.stagger_ns = sg ? atoi(sg) : 3000;
    // debug: trace only the k-th layer launch of each forward (FVTG_TRACE_LAYER_IDX, 11 launches per forward at QVH)
    const char* ti = getenv("FVTG_TRACE_LAYER_IDX");
    if (ti && a2.trace) {
      static thread_local long long launch_no = 0;
      const char* per = getenv("FVTG_TRACE_LAYER_PERIOD");
      const int period = per ? atoi(per) : 11;
      if ((launch_no++ % period) != atoi(ti)) a2.trace = nullptr;
    }
  }
  ProfScope prof(st, PC_LAYER);
  FVTG_CUDA_OK(launch_pdl(layer_kernel, dim3(grid), dim3(LK_THREADS), LK_SMEM_BYTES, st, ta, two, tw1, tw2, a2));
  FVTG_LAUNCH_CHECK("layer_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
