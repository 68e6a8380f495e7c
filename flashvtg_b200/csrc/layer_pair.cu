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
//   EPI3: y = LayerNorm2(Z + b2) -> fp32 residual stream, bf16(y), bf16(y + pos)
//
// 20 warps: 0 and 3 TMA producers (even / odd stages), 1 MMA issuer, 2 TMEM allocator, 4..19 epilogue (four threads per row: TMEM lane quadrant = warp % 4, column
// quarter = (warp - 4) / 4).
//
// CTA pairs (cta_group::2): two CTAs on the two SMs of a TPC form a cluster and work on two
// adjacent 128-row tiles with ONE stream of M = 256 MMAs issued by the leader CTA.  Each CTA
// keeps its own rows (A operand, TMEM accumulators, epilogues) but stages only HALF of every
// weight tile (B's N rows are split between the two shared memories), which halves both the
// L2 -> SM weight traffic and the shared-memory fill bandwidth that bounded the single-CTA
// version.  Loads of both CTAs complete on the leader's "full" barriers;
// tcgen05.commit multicasts every MMA -> {TMA, epilogue} hand-off to both CTAs; the peer's
// epilogue warps arrive remotely on the leader's barriers.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int LP_THREADS = 640;
constexpr int LP_STAGES = 8;
constexpr int LP_UNIT = 128 * 64 * 2;        // 16 KB: [128 rows][64 bf16], SWIZZLE_128B
constexpr int LP_STAGE = LP_UNIT;            // 16 KB per CTA (its half of a 32 KB weight stage)
constexpr int LP_OFF_A = 0;                  // 4 units: att tile, then LN1 output
constexpr int LP_OFF_W = 4 * LP_UNIT;        // weight ring
constexpr int LP_OFF_BAR = LP_OFF_W + LP_STAGES * LP_STAGE;
constexpr int LP_OFF_STAT = LP_OFF_BAR + 256;
constexpr int LP_OFF_PAR = LP_OFF_STAT + 2 * 4 * 128 * 8;   // [2 phases][4 quarters][128 rows] float2
constexpr int LP_PAR_FLOATS = 256 * 3 + 1024 + 256 * 3;
constexpr int LP_SMEM_BYTES = LP_OFF_PAR + LP_PAR_FLOATS * 4 + 1024 /*align slack*/;
static_assert(LP_SMEM_BYTES <= 232448, "layer kernel shared memory over the 227 KB limit");

// trace slots: [role 0 MMA / 1 epilogue][tile it < 8][event < 32]
#define LP_TRACE(role, ev)                                                              \
  do {                                                                                  \
    if (g.trace && blockIdx.x == 0 && it < 8)                                           \
      g.trace[((role) * 8 + it) * 32 + (ev)] = clock64();                               \
  } while (0)

__device__ __forceinline__ void lp_epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ float4 lp_ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__global__ void __launch_bounds__(LP_THREADS, 1)
layer_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWo,
             const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
             const __grid_constant__ LayerArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ symbol (an integer round trip
  // would demote every later access to a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + LP_OFF_A;
  uint8_t* sW = smem + LP_OFF_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LP_OFF_BAR);
  uint64_t* full = bars;                    // [8] TMA (both CTAs) -> MMA   (used in the leader)
  uint64_t* empty = bars + LP_STAGES;       // [8] MMA -> TMA               (multicast to both)
  uint64_t* a_full = bars + 16;             // att tiles of both CTAs landed (leader)
  uint64_t* a_empty = bars + 17;            // last ff1 MMA retired: sA reusable (both)
  uint64_t* z1_full = bars + 18;            // out_proj accumulated in H (both)
  uint64_t* z2_full = bars + 19;            // FFN accumulated in Z (both)
  uint64_t* ln_ready = bars + 20;           // LN1 tile in sA, residual in Z, H drained (leader, 32 warps)
  uint64_t* hacc_full = bars + 21;          // [2] ff1 piece accumulated (both)
  uint64_t* h_ready = bars + 23;            // [2] bf16 hidden piece stored back into H[buf] (leader, 32 warps)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 28);
  float2* s_stat = reinterpret_cast<float2*>(smem + LP_OFF_STAT);  // [2 phases][4 quarters][128]
  float* s_par = reinterpret_cast<float*>(smem + LP_OFF_PAR);
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
  const uint32_t rank = cluster_ctarank();       // 0 = leader of the pair
  const int npairs = (ntiles + 1) >> 1;
  const int nclusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmWo);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < LP_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(z1_full, 1);
    mbar_init(z2_full, 1);
    mbar_init(ln_ready, 32);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hacc_full[b], 1);
      mbar_init(&h_ready[b], 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_holder, 512);
    tmem_relinquish_cg2();
  }
  for (int i = threadIdx.x; i < 256; i += LP_THREADS) {
    s_bo[i] = g.bo[i];
    s_g1[i] = g.g1[i];
    s_be1[i] = g.be1[i];
    s_b2[i] = g.b2[i];
    s_g2[i] = g.g2[i];
    s_be2[i] = g.be2[i];
  }
  for (int i = threadIdx.x; i < 1024; i += LP_THREADS) s_b1[i] = g.b1[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tmem_h = tmem + 256;
  if (g.stagger_ns > 0) {
    // De-phase the clusters: every tile alternates an HBM-bound phase (epilogues 1 / 3) with a
    // tensor-bound one (FFN); started together, all SMs hit HBM at the same time and then leave it
    // idle.  Clusters that own one tile fewer than the critical ones start late by up to ~1 tile.
    const int my_tiles = (npairs - cluster_id + nclusters - 1) / nclusters;
    const int max_tiles = (npairs + nclusters - 1) / nclusters;
    if (my_tiles < max_tiles || g.stagger_ns < 0) {
      const long long t_end = clock64() + static_cast<long long>(cluster_id % 8) * g.stagger_ns * 2;  // ~2 cycles / ns
      while (clock64() < t_end) __nanosleep(200);
    }
  }

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------ TMA producer --
    // Runs in BOTH CTAs: each loads its own att tile and its half of every weight stage into its
    // own shared memory; all completion bytes are posted to the leader's barriers.  Two producer
    // lanes (warps 0 and 3) walk the same stage sequence and issue the even / odd stages: a single
    // lane sustains only one TMA instruction per ~350 cycles (profiles/r01c_probe_tma_mma.log).
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int half = static_cast<int>(rank);
      const int pid = warp == 0 ? 0 : 1;
      auto stage_wait = [&]() -> uint8_t* {   // nullptr: the other producer lane owns this stage
        if ((s & 1) != pid) return nullptr;
        mbar_wait(&empty[s], ph ^ 1);
        if (rank == 0) mbar_expect_tx(&full[s], 2 * LP_STAGE);   // both CTAs' halves
        return sW + s * LP_STAGE;
      };
      auto stage_next = [&]() {
        if (++s == LP_STAGES) { s = 0; ph ^= 1; }
      };
      auto ff1 = [&](int p) {  // this CTA's 64 of the 128 W1 rows of piece p, k-blocks (2h, 2h+1)
        for (int h = 0; h < 2; ++h) {
          uint8_t* d = stage_wait();
          if (d) {
            const uint32_t fb = mapa_u32(&full[s], 0);
            tma_load_2d_cg2(d, &tmW1, (2 * h) * 64, p * 128 + half * 64, fb);
            tma_load_2d_cg2(d + LP_UNIT / 2, &tmW1, (2 * h + 1) * 64, p * 128 + half * 64, fb);
          }
          stage_next();
        }
      };
      auto ff2 = [&](int p) {  // this CTA's 128 of the 256 W2 rows, k-block p*128 + h*64
        for (int h = 0; h < 2; ++h) {
          uint8_t* d = stage_wait();
          if (d) tma_load_2d_cg2(d, &tmW2, p * 128 + h * 64, half * 128, mapa_u32(&full[s], 0));
          stage_next();
        }
      };
      const uint32_t a_full_leader = mapa_u32(a_full, 0);
      int it = 0;
      for (int pair = cluster_id; pair < npairs; pair += nclusters, ++it) {
        const int tile = 2 * pair + half;   // rows past M are zero-filled by TMA
        if (pid == 0) {
          mbar_wait(a_empty, (it & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(a_full, 2 * 4 * LP_UNIT);
          for (int kb = 0; kb < 4; ++kb)
            tma_load_2d_cg2(sA + kb * LP_UNIT, &tmA, kb * 64, tile * 128, a_full_leader);
        }
        for (int kb = 0; kb < 4; ++kb) {   // this CTA's 128 of the 256 Wo rows
          uint8_t* d = stage_wait();
          if (d) tma_load_2d_cg2(d, &tmWo, kb * 64, half * 128, mapa_u32(&full[s], 0));
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
    // Leader CTA only: every instruction is an M = 256 MMA over the pair.
    if (lane == 0 && rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t idesc128 = umma_idesc_bf16(256, 128);
      const uint32_t idesc256 = umma_idesc_bf16(256, 256);
      const uint32_t sA_u = smem_u32(sA), sW_u = smem_u32(sW);
      const uint16_t both = 3;
      auto stage_wait = [&]() -> uint32_t {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        return sW_u + s * LP_STAGE;
      };
      auto stage_release = [&]() {
        umma_commit_cg2(&empty[s], both);
        if (++s == LP_STAGES) { s = 0; ph ^= 1; }
      };
      int it = 0;
      for (int pair = cluster_id; pair < npairs; pair += nclusters, ++it) {
        auto ff1 = [&](int p) {
          const int buf = p & 1;
          const uint32_t d = tmem_h + buf * 128;
          for (int h = 0; h < 2; ++h) {
            const uint32_t w = stage_wait();
#pragma unroll
            for (int kbl = 0; kbl < 2; ++kbl) {
              const uint64_t da = umma_desc_sw128(sA_u + (2 * h + kbl) * LP_UNIT);
              const uint64_t db = umma_desc_sw128(w + kbl * (LP_UNIT / 2));   // 64 rows x 64 k per CTA
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_cg2(d, da + 2 * k, db + 2 * k, idesc128, (h | kbl | k) ? 1u : 0u);
            }
            stage_release();
          }
          umma_commit_cg2(&hacc_full[buf], both);
          if (p == 7) umma_commit_cg2(a_empty, both);
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
              umma_bf16_ts_cg2(tmem, a + 32 * (j >> 1) + 8 * (j & 1), db + 2 * k, idesc256, 1u);
            }
            stage_release();
          }
        };
        LP_TRACE(0, 0);
        mbar_wait(a_full, it & 1);
        tc_fence_after();
        LP_TRACE(0, 2);
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t w = stage_wait();
          const uint64_t da = umma_desc_sw128(sA_u + kb * LP_UNIT);
          const uint64_t db = umma_desc_sw128(w);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_cg2(tmem_h, da + 2 * k, db + 2 * k, idesc256, (kb | k) ? 1u : 0u);
          stage_release();
        }
        umma_commit_cg2(z1_full, both);
        LP_TRACE(0, 3);
        mbar_wait(ln_ready, it & 1);
        tc_fence_after();
        LP_TRACE(0, 4);
        ff1(0);
        ff1(1);
        LP_TRACE(0, 5);
        for (int p = 0; p < 8; ++p) {
          ff2(p);
          LP_TRACE(0, 6 + 2 * p);
          if (p + 2 < 8) ff1(p + 2);
          LP_TRACE(0, 7 + 2 * p);
        }
        umma_commit_cg2(z2_full, both);
        LP_TRACE(0, 22);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue --
    const int ew = warp - 4;
    const int q = ew & 3;    // TMEM lane quadrant (== warp % 4)
    const int qt = ew >> 2;  // column quarter
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t u[32];
    // the leader's MMA thread consumes ln_ready / h_ready of BOTH CTAs
    const uint32_t ln_ready_leader = mapa_u32(ln_ready, 0);
    const uint32_t h_ready_leader[2] = {mapa_u32(&h_ready[0], 0), mapa_u32(&h_ready[1], 0)};
    int it = 0;
    for (int pair = cluster_id; pair < npairs; pair += nclusters, ++it) {
      const int tile = 2 * pair + static_cast<int>(rank);
      const int row = tile * 128 + r;
      const bool inb = row < g.M;
      const bool ldres = inb && !(g.dbg & 1);
      // fp32 residual stream, tile-blocked: [tile][col/4][row%128][4]
      float* yblk = g.yf + static_cast<size_t>(tile) * (128 * 256) + r * 4;
      const bool tr = (warp == 4 && lane == 0);

      // ---- epilogue 1: z = H + bo + x ; LayerNorm1 -> sA ; FFN residual into Z -----------------
      if (tr) LP_TRACE(1, 0);
      float4 rr[8];
      {  // the first residual chunk does not depend on the MMA: fetch it before waiting
        const int c0 = qt * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rr[i] = ldres ? lp_ld_f4(yblk + ((c0 >> 2) + i) * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(z1_full, it & 1);
      tc_fence_after();
      if (tr) LP_TRACE(1, 1);
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = qt * 64 + c * 32;
        tmem_ld32(tmem_h + lane_addr + c0, u);
        if (c == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rr[i] = ldres ? lp_ld_f4(yblk + ((c0 >> 2) + i) * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = lp_ld_f4(s_bo + c0 + 4 * i);
          u[4 * i + 0] = __float_as_uint(__uint_as_float(u[4 * i + 0]) + b.x + rr[i].x);
          u[4 * i + 1] = __float_as_uint(__uint_as_float(u[4 * i + 1]) + b.y + rr[i].y);
          u[4 * i + 2] = __float_as_uint(__uint_as_float(u[4 * i + 2]) + b.z + rr[i].z);
          u[4 * i + 3] = __float_as_uint(__uint_as_float(u[4 * i + 3]) + b.w + rr[i].w);
        }
        if (c == 0) shift = __uint_as_float(u[0]);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = __uint_as_float(u[j]) - shift;
          s1 += d;
          s2 += d * d;
        }
        tmem_st32(tmem + lane_addr + c0, u);
      }
      tmem_st_wait();
      if (tr) LP_TRACE(1, 2);
      // per-quarter partial: (mean of 64, centred sum of squares of 64)
      s_stat[qt * 128 + r] = make_float2(shift + s1 * (1.f / 64.f), s2 - s1 * s1 * (1.f / 64.f));
      lp_epi_bar_sync();
      float mean, rstd;
      {
        const float2 a0 = s_stat[r], a1 = s_stat[128 + r], a2 = s_stat[256 + r], a3 = s_stat[384 + r];
        mean = 0.25f * (a0.x + a1.x + a2.x + a3.x);
        const float d0 = a0.x - mean, d1 = a1.x - mean, d2 = a2.x - mean, d3 = a3.x - mean;
        const float m2 = a0.y + a1.y + a2.y + a3.y + 64.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
        rstd = rsqrtf(fmaxf(m2 * (1.f / 256.f), 0.f) + 1e-5f);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = qt * 64 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 gm = lp_ld_f4(s_g1 + c0 + 4 * i), bt = lp_ld_f4(s_be1 + c0 + 4 * i);
          v[4 * i + 0] = (__uint_as_float(u[4 * i + 0]) - mean) * rstd * gm.x + bt.x;
          v[4 * i + 1] = (__uint_as_float(u[4 * i + 1]) - mean) * rstd * gm.y + bt.y;
          v[4 * i + 2] = (__uint_as_float(u[4 * i + 2]) - mean) * rstd * gm.z + bt.z;
          v[4 * i + 3] = (__uint_as_float(u[4 * i + 3]) - mean) * rstd * gm.w + bt.w;
        }
        st_shared_bf16x32(sA + (c0 >> 6) * LP_UNIT, r, (c0 & 63) >> 3, v);
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
      if (lane == 0) mbar_arrive_cluster(ln_ready_leader);
      if (tr) LP_TRACE(1, 3);

      // ---- epilogue 2 (x8): hidden piece = PReLU(H[buf] + b1) -> bf16 pairs back into H[buf] --
#pragma unroll 1
      for (int p = 0; p < 8; ++p) {
        const int buf = p & 1;
        const uint32_t n = static_cast<uint32_t>(it) * 4u + static_cast<uint32_t>(p >> 1);
        mbar_wait(&hacc_full[buf], n & 1u);
        tc_fence_after();
        if (tr) LP_TRACE(1, 4 + 2 * p);
        const uint32_t ha = tmem_h + lane_addr + buf * 128 + qt * 32;
        tmem_ld32(ha, u);
        tmem_ld_wait();
        const float* bb = s_b1 + p * 128 + qt * 32;
        uint32_t hp[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = lp_ld_f4(bb + 4 * i);
          float x0 = __uint_as_float(u[4 * i + 0]) + b.x;
          float x1 = __uint_as_float(u[4 * i + 1]) + b.y;
          float x2 = __uint_as_float(u[4 * i + 2]) + b.z;
          float x3 = __uint_as_float(u[4 * i + 3]) + b.w;
          x0 = x0 > 0.f ? x0 : g.prelu * x0;
          x1 = x1 > 0.f ? x1 : g.prelu * x1;
          x2 = x2 > 0.f ? x2 : g.prelu * x2;
          x3 = x3 > 0.f ? x3 : g.prelu * x3;
          hp[2 * i + 0] = pack_bf16(x0, x1);
          hp[2 * i + 1] = pack_bf16(x2, x3);
        }
        tmem_st16(ha, hp);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(h_ready_leader[buf]);
        if (tr) LP_TRACE(1, 5 + 2 * p);
      }

      // ---- final epilogue: y = LayerNorm2(Z + b2) -> residual stream / bf16 operands --------
      int prow = row;
      if (g.pos_mod > 0) prow = row % g.pos_mod;
      const bool st_pos = g.out_pb && inb && (g.pos_rowlim <= 0 || prow < g.pos_rowlim);
      const bool ld_pos = st_pos && g.pos && !(g.dbg & 2);
      mbar_wait(z2_full, it & 1);
      tc_fence_after();
      if (tr) LP_TRACE(1, 20);
      shift = 0.f; s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = qt * 64 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        tmem_ld_wait();
        if (c == 0) shift = __uint_as_float(u[0]) + s_b2[c0];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = lp_ld_f4(s_b2 + c0 + 4 * i);
          const float d0 = __uint_as_float(u[4 * i + 0]) + b.x - shift;
          const float d1 = __uint_as_float(u[4 * i + 1]) + b.y - shift;
          const float d2 = __uint_as_float(u[4 * i + 2]) + b.z - shift;
          const float d3 = __uint_as_float(u[4 * i + 3]) + b.w - shift;
          s1 += (d0 + d1) + (d2 + d3);
          s2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      }
      s_stat[512 + qt * 128 + r] = make_float2(shift + s1 * (1.f / 64.f), s2 - s1 * s1 * (1.f / 64.f));
      lp_epi_bar_sync();
      if (tr) LP_TRACE(1, 21);
      {
        const float2 a0 = s_stat[512 + r], a1 = s_stat[640 + r], a2 = s_stat[768 + r], a3 = s_stat[896 + r];
        mean = 0.25f * (a0.x + a1.x + a2.x + a3.x);
        const float d0 = a0.x - mean, d1 = a1.x - mean, d2 = a2.x - mean, d3 = a3.x - mean;
        const float m2 = a0.y + a1.y + a2.y + a3.y + 64.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
        rstd = rsqrtf(fmaxf(m2 * (1.f / 256.f), 0.f) + 1e-5f);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = qt * 64 + c * 32;
        tmem_ld32(tmem + lane_addr + c0, u);
        if (ld_pos) {
          if (g.pos_mod > 0) {
            const float* ps = g.pos + static_cast<size_t>(prow) * 256 + c0;
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = lp_ld_f4(ps + 4 * i);
          } else if (g.pos_cmp_L > 0) {
            const float* ps = g.pos + (static_cast<size_t>(c0 >> 2) * g.pos_cmp_L + row % g.pos_cmp_L) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = lp_ld_f4(ps + static_cast<size_t>(i) * g.pos_cmp_L * 4);
          } else {
            const float* ps = g.pos + static_cast<size_t>(tile) * (128 * 256) + r * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) rr[i] = lp_ld_f4(ps + ((c0 >> 2) + i) * 512);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = lp_ld_f4(s_b2 + c0 + 4 * i), gm = lp_ld_f4(s_g2 + c0 + 4 * i),
                       bt = lp_ld_f4(s_be2 + c0 + 4 * i);
          v[4 * i + 0] = (__uint_as_float(u[4 * i + 0]) + b.x - mean) * rstd * gm.x + bt.x;
          v[4 * i + 1] = (__uint_as_float(u[4 * i + 1]) + b.y - mean) * rstd * gm.y + bt.y;
          v[4 * i + 2] = (__uint_as_float(u[4 * i + 2]) + b.z - mean) * rstd * gm.z + bt.z;
          v[4 * i + 3] = (__uint_as_float(u[4 * i + 3]) + b.w - mean) * rstd * gm.w + bt.w;
        }
        const bool st_ok = inb && !(g.dbg & 4);
        if (st_ok) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(yblk + ((c0 >> 2) + i) * 512) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (g.out_b)
          st_global_bf16x32_paired(g.out_b + static_cast<size_t>(row) * 256 + c0, 256, st_ok, v);
        if (g.out_pb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i + 0] += rr[i].x;
            v[4 * i + 1] += rr[i].y;
            v[4 * i + 2] += rr[i].z;
            v[4 * i + 3] += rr[i].w;
          }
          st_global_bf16x32_paired(g.out_pb + static_cast<size_t>(row) * 256 + c0, 256,
                                   st_pos && !(g.dbg & 4), v);
        }
      }
      // Z is rewritten next by this same thread (epilogue 1 of the next tile), H by MMAs that
      // are ordered behind ln_ready: no further hand-off is needed here.
      tc_fence_before();
      if (tr) LP_TRACE(1, 22);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's half of B / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem, 512);
  }
}

int launch_layer_pair(cudaStream_t st, const bf16* att, const bf16* wo, const bf16* w1, const bf16* w2,
                 const LayerArgs& args) {
  if (args.M <= 0) return FVTG_OK;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(layer_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      LP_SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, two, tw1, tw2;
  FVTG_TRY(make_tmap_bf16(&ta, att, args.M, 256, 256, 128, 64));
  FVTG_TRY(make_tmap_bf16(&two, wo, 256, 256, 256, 128, 64));     // per-CTA half: 128 of 256 rows
  FVTG_TRY(make_tmap_bf16(&tw1, w1, 1024, 256, 256, 64, 64));     // per-CTA half: 64 of a piece's 128 rows
  FVTG_TRY(make_tmap_bf16(&tw2, w2, 256, 1024, 1024, 128, 64));   // per-CTA half: 128 of 256 rows
  const int tiles = (args.M + 127) / 128;
  const int pairs = (tiles + 1) / 2;
  const int max_clusters = sm_count() / 2;
  const int clusters = pairs < max_clusters ? pairs : max_clusters;
  LayerArgs a2 = args;
  {
    const char* d = getenv("FVTG_LAYER_DBG");
    a2.dbg = d ? atoi(d) : 0;
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
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(LP_THREADS);
  cfg.dynamicSmemBytes = LP_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  ProfScope prof(st, PC_LAYER);
  FVTG_CUDA_OK(cudaLaunchKernelEx(&cfg, layer_pair_kernel, ta, two, tw1, tw2, a2));
  FVTG_LAUNCH_CHECK("layer_pair_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
