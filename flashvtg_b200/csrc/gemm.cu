#include "gemm.cuh"

namespace fvtg {

struct RowInfo {
  bool inb;    // row < M: anything may be stored
  bool valid;  // row carries a real (video, position)
  int dst;     // destination row in the primary outputs
  int b, lvl, n;  // video in chunk, pyramid level, canonical point index
  int x1, x2;     // rows in the extra destinations
};

__device__ __forceinline__ int pyr_off(const int vlen, int l) {
  int o = 0;
  for (int i = 0; i < l; ++i) o += vlen >> i;
  return o;
}

__device__ __forceinline__ RowInfo map_row(const GemmArgs& g, int row) {
  RowInfo ri;
  ri.inb = row < g.M;
  ri.valid = ri.inb;
  ri.dst = row;
  ri.b = 0; ri.lvl = 0; ri.n = 0; ri.x1 = row; ri.x2 = row;
  const GemmEpi& e = g.epi;
  if (!ri.inb) return ri;
  switch (e.rowmap) {
    case RM_TXT: {
      const int b = row / e.rm_a, j = row - b * e.rm_a;
      ri.dst = b * e.rm_b + e.rm_c + j;
      ri.x1 = ri.dst;
      break;
    }
    case RM_CHAIN: {
      const int pitch = e.geo.P0 >> e.rm_a;
      const int b = row / pitch, i = row - b * pitch;
      const int vl = e.geo.vlen[b];
      ri.valid = i < (vl >> e.rm_a);
      ri.b = b;
      if (e.rm_c && ri.valid) {
        ri.x1 = b * e.geo.PH1 + e.geo.o1[e.rm_b] + i;
        ri.x2 = b * e.geo.PH2 + e.geo.pad + pyr_off(vl, e.rm_b) + i;
      }
      break;
    }
    case RM_H1: {
      const int b = row / e.geo.PH1, q = row - b * e.geo.PH1;
      int l = 0;
      for (int i = 1; i < e.geo.nlev; ++i)
        if (q >= e.geo.o1[i]) l = i;
      const int i = q - e.geo.o1[l];
      const int vl = e.geo.vlen[b];
      ri.valid = (i >= 0) && (i < (vl >> l));
      ri.b = b; ri.lvl = l;
      ri.n = pyr_off(vl, l) + i;
      break;
    }
    case RM_H2: {
      const int b = row / e.geo.PH2, q = row - b * e.geo.PH2;
      const int vl = e.geo.vlen[b];
      const int n = q - e.geo.pad;
      ri.valid = (n >= 0) && (n < pyr_off(vl, e.geo.nlev));
      ri.b = b; ri.n = n;
      break;
    }
    default: break;
  }
  return ri;
}

// Branch-free activation: slope = 1 (none), 0 (ReLU), a (PReLU) on the negative side.
__device__ __forceinline__ float act_slope(int act, float a) {
  return act == ACT_RELU ? 0.f : (act == ACT_PRELU ? a : 1.f);
}
__device__ __forceinline__ float apply_act(float v, float slope) {
  return fmaxf(v, 0.f) + slope * fminf(v, 0.f);
}

// 32 consecutive columns of one row, starting at column c0: row-major (stride 4 floats between the
// 16-byte groups) or tile-blocked (stride 512 floats).
__device__ __forceinline__ void store_f32x32(float* base, size_t row, int c0, bool blocked,
                                             const float* y) {
  float* p = blocked ? base + blk_off(row, c0) : base + row * 256 + c0;
  const int stride = blocked ? 512 : 4;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(p + q * stride) =
        make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
}
__device__ __forceinline__ void add_f32x32(const float* base, size_t row, int c0, bool blocked,
                                           float* v) {
  const float* p = blocked ? base + blk_off(row, c0) : base + row * 256 + c0;
  const int stride = blocked ? 512 : 4;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 r4 = *reinterpret_cast<const float4*>(p + q * stride);
    v[q * 4 + 0] += r4.x; v[q * 4 + 1] += r4.y;
    v[q * 4 + 2] += r4.z; v[q * 4 + 3] += r4.w;
  }
}

#define GK_TRACE(role, ev)                                                        \
  do {                                                                            \
    if (g.trace && blockIdx.x == 0 && it < 16)                                    \
      g.trace[1024 + ((role) * 16 + it) * 8 + (ev)] = clock64();                  \
  } while (0)

// v[0..31] = u[0..31] + bias[0..31] with 16-byte shared-memory loads
__device__ __forceinline__ void add_bias32(const uint32_t* u, const float* bias, float* v) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 b = b4[i];
    v[4 * i + 0] = __uint_as_float(u[4 * i + 0]) + b.x;
    v[4 * i + 1] = __uint_as_float(u[4 * i + 1]) + b.y;
    v[4 * i + 2] = __uint_as_float(u[4 * i + 2]) + b.z;
    v[4 * i + 3] = __uint_as_float(u[4 * i + 3]) + b.w;
  }
}
// activation with the switch outside the element loop (act is uniform for the launch)
__device__ __forceinline__ void act32(float* v, int act, float slope) {
  if (act == ACT_NONE) return;
  if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], slope);
  }
}

__device__ __forceinline__ void gemm_epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// The whole kernel as a device function of ONE problem (tensor maps + arguments living in kernel
// parameter space): gemm_kernel runs it for its single problem, gemm_group_kernel lets blockIdx.y pick
// one of up to four independent problems that then share the SMs (gridDim.x CTAs each).
// PAIR: two CTAs of a cluster (the two SMs of a TPC) share every MMA (cta_group::2, M = 256): each
// stages its own 128 A rows and HALF of the weight tile, so the weight traffic out of L2 - which bounds
// the long-K GEMMs (7.8 TB/s of L2 reads for the K = 768 head convs) - is halved.
template <bool PAIR>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                                          const CUtensorMap& tmOut, const GemmArgs& g) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ symbol (an integer round trip
  // would demote every later access to a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // a CTA of a pair stages only half of every weight tile: 32 KB stages, and four of them fit where the
  // single CTA keeps three 48 KB stages - one more k-block of prefetch depth on a latency-bound pipeline
  constexpr int NSTAGE = PAIR ? 4 : GEMM_STAGES;
  constexpr int STAGE_BYTES = PAIR ? GEMM_A_BYTES + GEMM_B_BYTES_MAX / 2 : GEMM_STAGE_BYTES;
  static_assert(NSTAGE * STAGE_BYTES <= GEMM_STAGES * GEMM_STAGE_BYTES, "stage ring outgrew its region");
  uint8_t* stage_out = smem + GEMM_STAGES * GEMM_STAGE_BYTES;  // 4 swizzled [128][64] bf16 units
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + GEMM_OUT_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* tfull = bars + 2 * NSTAGE;
  uint64_t* tempty = bars + 2 * NSTAGE + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  float* s_bias = reinterpret_cast<float*>(stage_out + GEMM_OUT_BYTES + 256);
  float* s_gamma = s_bias + 1024;
  float* s_beta = s_gamma + 256;
  float* s_dotw = s_beta + 256;
  float2* s_stat = reinterpret_cast<float2*>(s_dotw + 256);  // [4 column quarters][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int BN = g.BN;
  const int m_tiles = (g.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = g.N / BN;
  const int total_kb = g.ntaps * g.kb_per_tap;
  // work items: (m-tile, n-tile), or (pair of m-tiles, n-tile) for a CTA pair; rank = CTA within the pair
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int num_tiles = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;
  const int first_tile = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], PAIR ? 32 : 16);   // epilogue warps of both CTAs arrive at the leader
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_cg2(tmem_holder, 512);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_holder, 512);
      tmem_relinquish();
    }
  }
  if (warp >= 4) {
    const GemmEpi& e = g.epi;
    const int t = threadIdx.x - 128;
    for (int i = t; i < g.N && i < 1024; i += 512) s_bias[i] = e.bias ? e.bias[i] : 0.f;
    if (e.mode == EPI_ROW && e.gamma) {
      for (int i = t; i < 256; i += 512) {
        s_gamma[i] = e.gamma[i];
        s_beta[i] = e.beta[i];
      }
    }
    if (e.mode == EPI_DOT)
      for (int i = t; i < 128; i += 512) s_dotw[i] = e.dotw[i];
  }
  pdl_wait();   // everything above touched only constants / on-chip state
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------------------------------- TMA producer --
    // PAIR: runs in both CTAs; each loads its own A rows and its half of the weight tile into its own
    // shared memory, all completion bytes are posted to the LEADER's full barrier.
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t b_bytes = (PAIR ? BN / 2 : BN) * GEMM_BK * 2;
      const uint32_t tx = (GEMM_A_BYTES + b_bytes) * (PAIR ? 2 : 1);
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int mp = tile / n_tiles, nt = tile - mp * n_tiles;
        const int mt = PAIR ? 2 * mp + static_cast<int>(rank) : mp;   // rows past M are zero-filled by TMA
        GK_TRACE(0, 0);
        const CUtensorMap* ta = (nt >= g.a_switch_ntile) ? &tmA2 : &tmA;
        for (int tap = 0; tap < g.ntaps; ++tap) {
          const int arow = mt * GEMM_BM + g.tap_shift[tap];
          for (int kb = 0; kb < g.kb_per_tap; ++kb) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + s * STAGE_BYTES;
            if (PAIR) {
              const uint32_t fb = mapa_u32(&full[s], 0);
              if (rank == 0) mbar_expect_tx(&full[s], tx);
              tma_load_2d_cg2(sa, ta, kb * GEMM_BK, arow, fb);
              tma_load_2d_cg2(sa + GEMM_A_BYTES, &tmB, (tap * g.kb_per_tap + kb) * GEMM_BK,
                              nt * BN + static_cast<int>(rank) * (BN / 2), fb);
            } else {
              mbar_expect_tx(&full[s], tx);
              tma_load_2d(sa, ta, kb * GEMM_BK, arow, &full[s]);
              tma_load_2d(sa + GEMM_A_BYTES, &tmB, (tap * g.kb_per_tap + kb) * GEMM_BK, nt * BN,
                          &full[s]);
            }
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        }
        GK_TRACE(0, 1);
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------- MMA issuer --
    // PAIR: the leader CTA alone issues, every instruction an M = 256 MMA over both CTAs' tiles
    if (lane == 0 && (!PAIR || rank == 0)) {
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * GEMM_BM : GEMM_BM, BN);
      const uint16_t both = 3;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t accph = (it >> 1) & 1;
        GK_TRACE(1, 0);
        mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        GK_TRACE(1, 1);
        const uint32_t tacc = tmem_base + acc * 256;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(sa);
          const uint64_t db = umma_desc_sw128(sa + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            if (PAIR) umma_bf16_cg2(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
            else umma_bf16(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          if (PAIR) umma_commit_cg2(&empty[s], both);
          else umma_commit(&empty[s]);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        if (PAIR) umma_commit_cg2(&tfull[acc], both);
        else umma_commit(&tfull[acc]);
        GK_TRACE(1, 2);
      }
    }
  } else if (warp >= 4) {
    // ----------------------------------------------------------- epilogue --
    // 16 warps, four threads per accumulator row: TMEM lane quadrant = warp % 4, column quarter = qt.
    const GemmEpi& e = g.epi;
    const int ew = warp - 4;
    const int wq = ew & 3, qt = ew >> 2;
    const int r = wq * 32 + lane;
    const int QC = BN >> 2;  // columns owned by this thread
    const bool leader = threadIdx.x == 128;
    const float slope = act_slope(e.act, e.prelu);
    // bf16 tile outputs with an identity row map leave through shared memory + TMA (coalesced)
    const bool staged = (e.mode == EPI_TILE) ||
                        (e.mode == EPI_ROW && e.out_bf16 && (e.rowmap == RM_NONE || e.rowmap == RM_CHAIN));
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int mp = tile / n_tiles, nt = tile - mp * n_tiles;
      const int mt = PAIR ? 2 * mp + static_cast<int>(rank) : mp;
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      const int row = mt * GEMM_BM + r;
      const int n0 = nt * BN;
      const RowInfo ri = map_row(g, row);
      if (leader) GK_TRACE(2, 0);
      if (staged && it > 0) {  // the previous tile's TMA store must be done reading the staging tile
        if (leader) tma_store_wait_read();
        gemm_epi_bar();
      }
      if (leader) GK_TRACE(2, 1);
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      if (leader) GK_TRACE(2, 2);
      const uint32_t tacc = tmem_base + acc * 256 + (static_cast<uint32_t>(wq * 32) << 16);
      uint32_t u[32];
      float v[32];

      if (e.mode == EPI_ROW) {
        const bool ln = e.gamma != nullptr;
        const bool has_res = e.res && ri.inb;
        const bool blk = e.f32_blocked != 0;
        float mean = 0.f, rstd = 1.f;
        if (ln) {
          float s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int c0 = qt * 64 + c * 32;
            tmem_ld32(tacc + c0, u);
            tmem_ld_wait();
            add_bias32(u, s_bias + c0, v);
            if (has_res) add_f32x32(e.res, row, c0, blk, v);
            act32(v, e.act, slope);
            if (c == 0) shift = v[0];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = v[j] - shift;
              s1 += d;
              s2 += d * d;
            }
            if (e.out_f32 && e.f32_preln && ri.inb) store_f32x32(e.out_f32, ri.dst, c0, blk, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(v[j]);
            tmem_st32(tacc + c0, u);
          }
          tmem_st_wait();
          // per-quarter partial (mean of 64, centred sum of squares of 64), combined over the 4 quarters
          s_stat[qt * 128 + r] = make_float2(shift + s1 * (1.f / 64.f), s2 - s1 * s1 * (1.f / 64.f));
          gemm_epi_bar();
          const float2 a0 = s_stat[r], a1 = s_stat[128 + r], a2 = s_stat[256 + r], a3 = s_stat[384 + r];
          mean = 0.25f * (a0.x + a1.x + a2.x + a3.x);
          const float d0 = a0.x - mean, d1 = a1.x - mean, d2 = a2.x - mean, d3 = a3.x - mean;
          const float m2 = a0.y + a1.y + a2.y + a3.y + 64.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
          rstd = rsqrtf(fmaxf(m2 * (1.f / 256.f), 0.f) + 1e-5f);
        }
        int prow = row;
        if (e.pos_mod > 0) prow = row % e.pos_mod;
        const bool st_pos = e.out_bf16_pos && ri.inb && (e.pos_rowlim <= 0 || prow < e.pos_rowlim);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int c0 = qt * 64 + c * 32;
          tmem_ld32(tacc + c0, u);
          tmem_ld_wait();
          if (ln) {
            const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c0);
            const float4* be4 = reinterpret_cast<const float4*>(s_beta + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 gm = g4[i], bt = be4[i];
              v[4 * i + 0] = (__uint_as_float(u[4 * i + 0]) - mean) * rstd * gm.x + bt.x;
              v[4 * i + 1] = (__uint_as_float(u[4 * i + 1]) - mean) * rstd * gm.y + bt.y;
              v[4 * i + 2] = (__uint_as_float(u[4 * i + 2]) - mean) * rstd * gm.z + bt.z;
              v[4 * i + 3] = (__uint_as_float(u[4 * i + 3]) - mean) * rstd * gm.w + bt.w;
            }
          } else {
            add_bias32(u, s_bias + c0, v);
            if (has_res) add_f32x32(e.res, row, c0, blk, v);
            act32(v, e.act, slope);
          }
          if (e.post_relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (!ri.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (staged) st_shared_bf16x32(stage_out + (c0 >> 6) * GEMM_A_BYTES, r, (c0 & 63) >> 3, v);
          if (ri.inb) {
            const size_t o = static_cast<size_t>(ri.dst) * 256 + c0;
            if (e.out_f32 && !(ln && e.f32_preln)) store_f32x32(e.out_f32, ri.dst, c0, blk, v);
            if (e.out_bf16 && !staged) st_global_bf16x32(e.out_bf16 + o, v);
            if (e.rowmap == RM_TXT) {
              if (e.out_x1) st_global_bf16x32(e.out_x1 + static_cast<size_t>(ri.x1) * 256 + c0, v);
            } else if (e.rowmap == RM_CHAIN && e.rm_c && ri.valid) {
              if (e.out_x1) st_global_bf16x32(e.out_x1 + static_cast<size_t>(ri.x1) * 256 + c0, v);
              if (e.out_x2) st_global_bf16x32(e.out_x2 + static_cast<size_t>(ri.x2) * 256 + c0, v);
            }
            if (st_pos) {
              if (e.pos && e.pos_cmp_L > 0) {
                const float* pc = e.pos + (static_cast<size_t>(c0 >> 2) * e.pos_cmp_L + row % e.pos_cmp_L) * 4;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 p4 = *reinterpret_cast<const float4*>(pc + static_cast<size_t>(q) * e.pos_cmp_L * 4);
                  v[q * 4 + 0] += p4.x; v[q * 4 + 1] += p4.y; v[q * 4 + 2] += p4.z; v[q * 4 + 3] += p4.w;
                }
              } else if (e.pos) add_f32x32(e.pos, prow, c0, blk && e.pos_mod <= 0, v);
              st_global_bf16x32(e.out_bf16_pos + o, v);
            }
          }
        }
      } else if (e.mode == EPI_TILE) {
        // lean inner loops: 16-byte bias loads, the activation switch hoisted out of the element loop
        for (int c0 = qt * QC; c0 < (qt + 1) * QC; c0 += 32) {
          tmem_ld32(tacc + c0, u);
          tmem_ld_wait();
          add_bias32(u, s_bias + n0 + c0, v);
          act32(v, e.act, slope);
          if (!ri.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          st_shared_bf16x32(stage_out + (c0 >> 6) * GEMM_A_BYTES, r, (c0 & 63) >> 3, v);
        }
      } else if (e.mode == EPI_DOT) {
        float dot = 0.f;
        {
          const int c0 = qt * 32;   // BN == 128: 32 columns per thread
          tmem_ld32(tacc + c0, u);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
          const float4* w4 = reinterpret_cast<const float4*>(s_dotw + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = b4[i], w = w4[i];
            dot += fmaxf(__uint_as_float(u[4 * i + 0]) + b.x, 0.f) * w.x;
            dot += fmaxf(__uint_as_float(u[4 * i + 1]) + b.y, 0.f) * w.y;
            dot += fmaxf(__uint_as_float(u[4 * i + 2]) + b.z, 0.f) * w.z;
            dot += fmaxf(__uint_as_float(u[4 * i + 3]) + b.w, 0.f) * w.w;
          }
        }
        if (qt > 0) s_stat[qt * 128 + r].x = dot;
        gemm_epi_bar();
        if (qt == 0 && ri.valid)
          e.out_dot[static_cast<size_t>(ri.b) * e.geo.n_max + ri.n] =
              ((dot + s_stat[128 + r].x) + (s_stat[256 + r].x + s_stat[384 + r].x)) + e.dotb;
        gemm_epi_bar();  // s_stat is rewritten by the next tile
      } else {  // EPI_COORD
        if (qt == 0) {
          tmem_ld16(tacc, u);
          tmem_ld_wait();
          if (ri.valid) {
            const float c = e.coef[ri.lvl];
            float* o = e.out_coord + (static_cast<size_t>(ri.b) * e.geo.n_max + ri.n) * 2;
            o[0] = expf(__uint_as_float(u[0]) + s_bias[0]) * c;
            o[1] = expf(__uint_as_float(u[1]) + s_bias[1]) * c;
          }
        }
      }
      tc_fence_before();
      if (leader) GK_TRACE(2, 3);
      if (staged) {
        fence_proxy_async_smem();
        gemm_epi_bar();
        if (leader) {
          const int units = (e.mode == EPI_ROW ? 256 : BN) >> 6;
          for (int kb = 0; kb < units; ++kb)
            tma_store_2d(&tmOut, stage_out + kb * GEMM_A_BYTES, n0 + kb * 64, mt * GEMM_BM);
          tma_store_commit();
        }
      }
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
        else mbar_arrive(&tempty[acc]);
      }
      if (leader) GK_TRACE(2, 4);
    }
    if (staged && leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer may still read this CTA's half of B / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

struct GemmProblem {
  CUtensorMap tmA, tmA2, tmB, tmOut;
  GemmArgs g;
};
constexpr int GEMM_MAX_GROUP = 4;
struct GemmGroup {
  GemmProblem p[GEMM_MAX_GROUP];
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
            const __grid_constant__ GemmArgs g) {
  gemm_body<false>(tmA, tmA2, tmB, tmOut, g);
}

// launched as clusters of 2 CTAs; tmB's box is HALF the weight tile (BN / 2 rows)
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ GemmArgs g) {
  gemm_body<true>(tmA, tmA2, tmB, tmOut, g);
}

// Independent small GEMMs in one launch (the pyramid steps of different levels): each problem gets
// gridDim.x persistent CTAs.  The switch keeps every problem's parameter offsets static.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_group_kernel(const __grid_constant__ GemmGroup grp) {
  switch (blockIdx.y) {
    case 0: gemm_body<false>(grp.p[0].tmA, grp.p[0].tmA2, grp.p[0].tmB, grp.p[0].tmOut, grp.p[0].g); break;
    case 1: gemm_body<false>(grp.p[1].tmA, grp.p[1].tmA2, grp.p[1].tmB, grp.p[1].tmOut, grp.p[1].g); break;
    case 2: gemm_body<false>(grp.p[2].tmA, grp.p[2].tmA2, grp.p[2].tmB, grp.p[2].tmOut, grp.p[2].g); break;
    default: gemm_body<false>(grp.p[3].tmA, grp.p[3].tmA2, grp.p[3].tmB, grp.p[3].tmOut, grp.p[3].g); break;
  }
}

static int gemm_check(const GemmArgs& args) {
  if (args.BN != 256 && args.BN != 128 && args.BN != 16)
    return fail(FVTG_EINVAL, "gemm: unsupported BN %d", args.BN);
  if (args.N % args.BN || args.N > 1024) return fail(FVTG_EINVAL, "gemm: bad N %d", args.N);
  if (args.ntaps < 1 || args.ntaps > GEMM_MAX_TAPS || args.kb_per_tap < 1)
    return fail(FVTG_EINVAL, "gemm: bad K tiling");
  if (args.epi.mode == EPI_ROW && (args.BN != 256 || args.N != 256))
    return fail(FVTG_EINVAL, "gemm: EPI_ROW needs N == BN == 256");
  if (args.epi.mode == EPI_DOT && (args.BN != 128 || args.N != 128))
    return fail(FVTG_EINVAL, "gemm: EPI_DOT needs N == BN == 128");
  if (args.epi.mode == EPI_COORD && (args.BN != 16 || args.N != 16))
    return fail(FVTG_EINVAL, "gemm: EPI_COORD needs N == BN == 16");
  return FVTG_OK;
}

static int gemm_build(GemmProblem* P, const void* a, const void* a2, uint64_t a_rows, uint64_t a_cols,
                      uint64_t a_pitch, const void* w, const GemmArgs& args, bool pair = false) {
  FVTG_TRY(make_tmap_bf16(&P->tmA, a, a_rows, a_cols, a_pitch, GEMM_BM, GEMM_BK));
  FVTG_TRY(make_tmap_bf16(&P->tmA2, a2 ? a2 : a, a_rows, a_cols, a_pitch, GEMM_BM, GEMM_BK));
  const uint64_t ktot = static_cast<uint64_t>(args.ntaps) * args.kb_per_tap * GEMM_BK;
  FVTG_TRY(make_tmap_bf16(&P->tmB, w, args.N, ktot, ktot, pair ? args.BN / 2 : args.BN, GEMM_BK));
  {  // bf16 tile output leaving through the staging tile + TMA store (rows >= M are clipped)
    const GemmEpi& e = args.epi;
    const void* optr = w;
    uint64_t ocols = ktot, opitch = ktot, orows = args.N;
    if (e.mode == EPI_TILE) {
      optr = e.out; ocols = e.ld_out; opitch = e.ld_out; orows = args.M;
    } else if (e.mode == EPI_ROW && e.out_bf16 && (e.rowmap == RM_NONE || e.rowmap == RM_CHAIN)) {
      optr = e.out_bf16; ocols = 256; opitch = 256; orows = args.M;
    }
    FVTG_TRY(make_tmap_bf16(&P->tmOut, optr, orows, ocols, opitch, GEMM_BM, GEMM_BK));
  }
  P->g = args;
  return FVTG_OK;
}

int launch_gemm_group(cudaStream_t st, int n, const GemmOperands* ops, const GemmArgs* args) {
  if (n < 1 || n > GEMM_MAX_GROUP) return fail(FVTG_EINVAL, "gemm group: 1..%d problems", GEMM_MAX_GROUP);
  if (n == 1)
    return launch_gemm(st, ops[0].a, ops[0].a2, ops[0].a_rows, ops[0].a_cols, ops[0].a_pitch, ops[0].w, args[0]);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(gemm_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GEMM_SMEM_BYTES));
    attr_set = true;
  }
  GemmGroup grp;
  memset(&grp, 0, sizeof(grp));
  int max_tiles = 0;
  for (int i = 0; i < n; ++i) {
    if (args[i].M <= 0) return fail(FVTG_EINVAL, "gemm group: empty problem");
    FVTG_TRY(gemm_check(args[i]));
    FVTG_TRY(gemm_build(&grp.p[i], ops[i].a, ops[i].a2, ops[i].a_rows, ops[i].a_cols, ops[i].a_pitch, ops[i].w,
                        args[i]));
    const int tiles = ((args[i].M + GEMM_BM - 1) / GEMM_BM) * (args[i].N / args[i].BN);
    if (tiles > max_tiles) max_tiles = tiles;
  }
  int per = sm_count() / n;
  if (per > max_tiles) per = max_tiles;
  ProfScope prof(st, PC_GEMM);
  FVTG_CUDA_OK(launch_pdl(gemm_group_kernel, dim3(per, n), dim3(GEMM_THREADS), GEMM_SMEM_BYTES, st, grp));
  FVTG_LAUNCH_CHECK("gemm_group_kernel");
  return FVTG_OK;
}

int launch_gemm(cudaStream_t st, const void* a, const void* a2, uint64_t a_rows, uint64_t a_cols,
                uint64_t a_pitch, const void* w, const GemmArgs& args) {
  if (args.M <= 0) return FVTG_OK;
  FVTG_TRY(gemm_check(args));
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GEMM_SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (args.M + GEMM_BM - 1) / GEMM_BM;
  {  // CTA pairs for the 256-wide tiles of GEMMs tall enough to fill the machine (FVTG_GEMM_PAIR=0: off)
    static const bool pair_on = [] { const char* e = getenv("FVTG_GEMM_PAIR"); return !e || atoi(e) != 0; }();
    const int pair_tiles = ((m_tiles + 1) / 2) * (args.N / args.BN);
    if (pair_on && args.BN == 256 && pair_tiles >= sm_count() / 2) {
      static thread_local bool pair_attr = false;
      if (!pair_attr) {
        FVTG_CUDA_OK(cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          GEMM_SMEM_BYTES));
        pair_attr = true;
      }
      GemmProblem P;
      FVTG_TRY(gemm_build(&P, a, a2, a_rows, a_cols, a_pitch, w, args, true));
      const int clusters = pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(2 * clusters);
      cfg.blockDim = dim3(GEMM_THREADS);
      cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = pdl_enabled() ? 2 : 1;
      ProfScope prof(st, PC_GEMM);
      FVTG_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_pair_kernel, P.tmA, P.tmA2, P.tmB, P.tmOut, P.g));
      FVTG_LAUNCH_CHECK("gemm_pair_kernel");
      return FVTG_OK;
    }
  }
  GemmProblem P;
  FVTG_TRY(gemm_build(&P, a, a2, a_rows, a_cols, a_pitch, w, args));
  const int tiles = m_tiles * (args.N / args.BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  ProfScope prof(st, PC_GEMM);
  FVTG_CUDA_OK(launch_pdl(gemm_kernel, dim3(grid), dim3(GEMM_THREADS), GEMM_SMEM_BYTES, st, P.tmA, P.tmA2, P.tmB,
                          P.tmOut, P.g));
  FVTG_LAUNCH_CHECK("gemm_kernel");
  return FVTG_OK;
}

}  // namespace fvtg

#ifdef FVTG_DEBUG_HOOKS   // test / tuning hook: only in libflashvtg_b200_dbg.so
extern "C" int32_t fvtg_dbg_gemm(const void* a, const void* w, const float* bias, float* out,
                                 int32_t M, int32_t N, int32_t K, int32_t act, void* stream) {
  using namespace fvtg;
  host_state().launches = 0;
  FVTG_TRY(check_arch());
  if (!a || !w || !out || M <= 0 || K % 64 || N % 128)
    return fail(FVTG_EINVAL, "dbg_gemm: need K %% 64 == 0 and N %% 128 == 0");
  // fp32 result through the EPI_ROW path needs N == 256; other N go through a bf16 tile store,
  // so the hook exposes both: N == 256 -> fp32 `out`; else `out` is reinterpreted as bf16 [M][N].
  GemmArgs g = gemm_args(M, N, N == 256 ? 256 : (N % 256 == 0 ? 256 : 128), K);
  g.trace = dbg_trace();
  g.epi.bias = bias;
  g.epi.act = act;
  if (N == 256) {
    g.epi.mode = EPI_ROW;
    g.epi.out_f32 = out;
  } else {
    g.epi.mode = EPI_TILE;
    g.epi.out = reinterpret_cast<bf16*>(out);
    g.epi.ld_out = N;
  }
  return launch_gemm(static_cast<cudaStream_t>(stream), a, nullptr, M, K, K, w, g);
}
#endif
