// Attention on the 5th-generation tensor cores: softmax(Q K^T / sqrt(32)) V for the three flavours
// of the path (dummy-token encoder / self-attention encoder: transformer.py:413-415; adaptive
// cross-attention: crossattention.py:287-396 - no projections, dummies absorb probability mass but
// carry no value, per-row text mass kept for t2vattnvalues, model.py:215).
//
// One CTA = (video, 128-query block, head pair).  A head pair is one 64-column SWIZZLE_128B unit of
// the Q / K / V rows, so every operand is a plain TMA box:
//   S_h = Q_h . K_h^T      tcgen05.mma 128 x NB x 16 (x2), both operands K-major from shared memory,
//                          accumulator in TMEM columns [0, NB)
//   softmax                one thread per query row (TMEM lane == row): two passes over the row in
//                          32-column tcgen05.ld chunks - valid-key maximum, then exp2 / sum - with
//                          the bf16 probabilities written back over the scores (tcgen05.st, two keys
//                          per 32-bit column): they never touch shared memory
//   O_h = P_h . V          tcgen05.mma 128 x 64 x 16 per 16 keys: A = P from TMEM, B = the V unit
//                          exactly as TMA delivered it (rows = keys) read as an MN-major operand;
//                          the head's own 32 of the 64 output columns are kept
// Keys beyond 128 (TACoS / Charades long videos) run as key blocks of 128 in two sweeps: sweep A
// only tracks the row maximum over all blocks, sweep B recomputes the scores, exponentiates against
// the final maximum and accumulates P.V in TMEM, so no rescaling of a partial O is ever needed.
// TMEM: 128 columns per CTA when the keys fit one block (S/P in [0,NB), O in [64,128) - the dead
// score columns), 256 columns otherwise (S, O_0, O_1) - up to 4 / 2 CTAs per SM cover each other's
// serial S -> softmax -> PV chain.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int ATC_THREADS = 160;  // warp 0: TMA + MMA issue + TMEM allocation; warps 1..4: softmax rows
constexpr int ATC_UNIT = 128 * 128;  // bytes of one [128 rows][64 bf16] unit

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool MULTI>
__global__ void __launch_bounds__(ATC_THREADS, MULTI ? 2 : 4)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnArgs a, const int QR, const int NB,
               const int nkb, const int nqb, const int kv_shared) {
  constexpr uint32_t TM_COLS = MULTI ? 256 : 128;
  extern __shared__ uint8_t atc_raw[];
  uint8_t* smem = atc_raw + ((1024u - (smem_u32(atc_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATC_UNIT;
  uint8_t* sV = kv_shared ? sK : sK + ATC_UNIT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * ATC_UNIT);
  uint64_t* bar_q = bars;
  uint64_t* bar_kv = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_ready = bars + 3;
  uint64_t* o_full = bars + 4;   // [2]
  uint64_t* o_free = bars + 6;
  uint64_t* mma_done = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x & 3;
  const int qb = (blockIdx.x >> 2) % nqb;
  const int b = (blockIdx.x >> 2) / nqb;

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      prefetch_tmap(&tmV);
      mbar_init(bar_q, 1);
      mbar_init(bar_kv, 1);
      mbar_init(s_full, 1);
      mbar_init(p_ready, 4);
      mbar_init(&o_full[0], 1);
      mbar_init(&o_full[1], 1);
      mbar_init(o_free, 4);
      mbar_init(mma_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, TM_COLS);
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  int klen = a.kbase + a.klen_src[b];
  if (klen > a.Lk) klen = a.Lk;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, static_cast<uint32_t>(QR) * 128u);
      tma_load_2d(sQ, &tmQ, pair * 64, b * a.Lq + qb * 128, bar_q);
      const uint32_t idS = umma_idesc_bf16(128, NB);
      const uint32_t idO = umma_idesc_bf16(128, 64) | (1u << 16);   // B operand MN-major (rows = keys)
      const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
      uint32_t kvph = 0, pph = 0, dph = 0;
      int loads = 0;
      for (int sweep = MULTI ? 0 : 1; sweep < 2; ++sweep) {
        for (int kb = 0; kb < nkb; ++kb) {
          if (loads > 0) {   // every MMA that reads the previous K / V block has retired
            mbar_wait(mma_done, dph);
            dph ^= 1;
          }
          const bool ldv = sweep == 1 && !kv_shared;
          mbar_expect_tx(bar_kv, static_cast<uint32_t>(NB) * 128u * (ldv ? 2u : 1u));
          tma_load_2d(sK, &tmK, pair * 64, b * a.Lk + kb * NB, bar_kv);
          if (ldv) tma_load_2d(sV, &tmV, pair * 64, b * a.Lk + kb * NB, bar_kv);
          if (loads == 0) mbar_wait(bar_q, 0);
          ++loads;
          mbar_wait(bar_kv, kvph);
          kvph ^= 1;
          tc_fence_after();
          for (int h = 0; h < 2; ++h) {
            // one-block layout: O of head 0 sits in the dead score columns [64,128) - the next
            // score tile may only land once the softmax warps have read it
            if (!MULTI && h == 1) {
              mbar_wait(o_free, 0);
              tc_fence_after();
            }
            const uint64_t dq = umma_desc_sw128(sQ_u) + 4 * h;   // head h = 64-byte half of the 128-byte row
            const uint64_t dk = umma_desc_sw128(sK_u) + 4 * h;
            umma_bf16(tmem, dq, dk, idS, 0u);
            umma_bf16(tmem, dq + 2, dk + 2, idS, 1u);
            umma_commit(s_full);
            mbar_wait(p_ready, pph);
            pph ^= 1;
            tc_fence_after();
            if (sweep == 1) {
              const uint32_t o = MULTI ? tmem + 128 + 64 * h : tmem + 64;
              const uint64_t dv = umma_desc_sw128(sV_u);
              for (int j = 0; j < NB / 16; ++j)   // 16 keys = 8 packed TMEM columns = 2048 bytes of V rows
                umma_bf16_ts(o, tmem + 8 * j, dv + 128 * j, idO, (kb | j) ? 1u : 0u);
              if (kb == nkb - 1) umma_commit(&o_full[h]);
            }
          }
          umma_commit(mma_done);
        }
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;            // query row within the block
    const int row = qb * 128 + r;
    const bool wactive = qb * 128 + q * 32 < a.Lq;   // warp-uniform: some row of this warp exists
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, ts[2] = {0.f, 0.f};
    uint32_t sph = 0;
    uint32_t u[32];
    for (int sweep = MULTI ? 0 : 1; sweep < 2; ++sweep) {
      for (int kb = 0; kb < nkb; ++kb) {
        for (int h = 0; h < 2; ++h) {
          mbar_wait(s_full, sph);
          sph ^= 1;
          tc_fence_after();
          if (wactive) {
            const int k0 = kb * NB;
            if (sweep == 0 || !MULTI) {   // maximum over the valid keys of this block
              float mx = m[h];
              for (int c0 = 0; c0 < NB; c0 += 32) {
                const int nv = klen - (k0 + c0);
                if (nv <= 0) break;
                tmem_ld32(tmem + lane_addr + c0, u);
                tmem_ld_wait();
                if (nv >= 32) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(u[i]));
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < nv ? __uint_as_float(u[i]) : -INFINITY);
                }
              }
              m[h] = mx;
            }
            if (sweep == 1) {
              const float off = (m[h] == -INFINITY ? 0.f : m[h]) * sc;
              float ls = 0.f, lt = 0.f;
              for (int c0 = 0; c0 < NB; c0 += 32) {
                const int nv = klen - (k0 + c0);          // keys of this chunk below klen
                const int n0 = a.v_first - (k0 + c0);     // keys of this chunk below v_first carry no value
                uint32_t pk[16];
                if (nv <= 0) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) pk[i] = 0u;
                } else {
                  tmem_ld32(tmem + lane_addr + c0, u);
                  tmem_ld_wait();
                  float p[32];
                  if (nv >= 32 && n0 <= 0) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                      p[i] = ex2f(fmaf(__uint_as_float(u[i]), sc, -off));
                      p[i + 1] = ex2f(fmaf(__uint_as_float(u[i + 1]), sc, -off));
                      s0 += p[i];
                      s1 += p[i + 1];
                    }
                    ls += s0 + s1;
                    lt += s0 + s1;
                  } else {
                    float s0 = 0.f, t0 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                      float e = ex2f(fmaf(__uint_as_float(u[i]), sc, -off));
                      e = i < nv ? e : 0.f;
                      s0 += e;
                      e = i >= n0 ? e : 0.f;
                      t0 += e;
                      p[i] = e;
                    }
                    ls += s0;
                    lt += t0;
                  }
#pragma unroll
                  for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(p[2 * i], p[2 * i + 1]);
                }
                tmem_st16(tmem + lane_addr + (c0 >> 1), pk);
              }
              tmem_st_wait();
              l[h] += ls;
              ts[h] += lt;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_ready);

          if (sweep == 1 && kb == nkb - 1) {   // head h is complete: O_h / l_h -> out
            mbar_wait(&o_full[h], 0);
            tc_fence_after();
            if (wactive) {
              const uint32_t o = (MULTI ? tmem + 128 + 64 * h : tmem + 64) + 32 * h;
              tmem_ld32(o + lane_addr, u);
              tmem_ld_wait();
            }
            tc_fence_before();
            if (!MULTI && h == 0) {
              __syncwarp();
              if (lane == 0) mbar_arrive(o_free);
            }
            if (wactive && row < a.Lq) {
              const float inv = l[h] > 0.f ? 1.f / l[h] : 0.f;
              const int hh = pair * 2 + h;
              const size_t grow = static_cast<size_t>(b) * a.Lq + row;
              uint4* dst = reinterpret_cast<uint4*>(a.out + grow * 256 + hh * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 v;
                v.x = pack_bf16(__uint_as_float(u[8 * i + 0]) * inv, __uint_as_float(u[8 * i + 1]) * inv);
                v.y = pack_bf16(__uint_as_float(u[8 * i + 2]) * inv, __uint_as_float(u[8 * i + 3]) * inv);
                v.z = pack_bf16(__uint_as_float(u[8 * i + 4]) * inv, __uint_as_float(u[8 * i + 5]) * inv);
                v.w = pack_bf16(__uint_as_float(u[8 * i + 6]) * inv, __uint_as_float(u[8 * i + 7]) * inv);
                dst[i] = v;
              }
              if (a.tsum) a.tsum[static_cast<size_t>(hh) * a.B * a.Lq + grow] = ts[h] * inv;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------
// Keys that fit one block (every QVHighlights / Charades-STA shape): a PERSISTENT variant of the
// same CTA.  Work items (video, query block, head pair) are dealt round-robin to the resident CTAs;
// Q / K / V of item i+1 stream into the second shared-memory stage while item i computes, barriers
// and the 128 TMEM columns are set up once per CTA, and the softmax arithmetic runs on packed
// fp32 pairs with three-input maxima in 8-key groups whose masking state (below klen / below
// v_first) is warp-uniform, so fully valid groups carry no select instructions at all.

// maximum over the first nv (of W) score columns in u[]
template <int W>
__device__ __forceinline__ float chunk_max(const uint32_t* u, int nv, float mx) {
#pragma unroll
  for (int g = 0; g < W / 8; ++g) {
    const int lo = g * 8;
    if (lo + 8 <= nv) {
#pragma unroll
      for (int j = 0; j < 8; j += 2)
        mx = max3f(mx, __uint_as_float(u[lo + j]), __uint_as_float(u[lo + j + 1]));
    } else if (lo < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mx = fmaxf(mx, lo + j < nv ? __uint_as_float(u[lo + j]) : -INFINITY);
    }
  }
  return mx;
}

// p = 2^(s * sc + noff) for the first nv columns (0 beyond); pk = bf16 pairs of p with the columns
// below n0 zeroed (keys that carry no value); acc_all / acc_val = packed partial sums of p over all
// valid keys / over the value-carrying ones
template <int W>
__device__ __forceinline__ void chunk_softmax(const uint32_t* u, uint32_t* pk, int nv, int n0, float sc,
                                              float noff, uint64_t& acc_all, uint64_t& acc_val) {
  const uint64_t sc2 = f2_pack(sc, sc), off2 = f2_pack(noff, noff);
#pragma unroll
  for (int g = 0; g < W / 8; ++g) {
    const int lo = g * 8;
    if (lo >= nv) {
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[g * 4 + j] = 0u;
    } else if (lo + 8 <= nv && lo >= n0) {          // 8 valid keys, all with a value
      uint64_t s2 = 0ull;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x, y;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(u[lo + 2 * j]), __uint_as_float(u[lo + 2 * j + 1])), sc2, off2), x, y);
        const float e0 = ex2f(x), e1 = ex2f(y);
        s2 = f2_add(s2, f2_pack(e0, e1));
        pk[g * 4 + j] = pack_bf16(e0, e1);
      }
      acc_all = f2_add(acc_all, s2);
      acc_val = f2_add(acc_val, s2);
    } else if (lo + 8 <= nv && lo + 8 <= n0) {      // 8 valid keys, none with a value
      uint64_t s2 = 0ull;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x, y;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(u[lo + 2 * j]), __uint_as_float(u[lo + 2 * j + 1])), sc2, off2), x, y);
        s2 = f2_add(s2, f2_pack(ex2f(x), ex2f(y)));
        pk[g * 4 + j] = 0u;
      }
      acc_all = f2_add(acc_all, s2);
    } else {                                         // the group straddles klen or v_first
      float sa = 0.f, sv = 0.f, e[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = ex2f(fmaf(__uint_as_float(u[lo + j]), sc, noff));
        t = lo + j < nv ? t : 0.f;
        sa += t;
        t = lo + j >= n0 ? t : 0.f;
        sv += t;
        e[j] = t;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[g * 4 + j] = pack_bf16(e[2 * j], e[2 * j + 1]);
      acc_all = f2_add(acc_all, f2_pack(sa, 0.f));
      acc_val = f2_add(acc_val, f2_pack(sv, 0.f));
    }
  }
}

#define ATC_TRACE(role, it, ev)                                                         \
  do {                                                                                  \
    if (a.trace && blockIdx.x == 0 && (it) < 8)                                         \
      a.trace[3072 + (role) * 128 + (it) * 16 + (ev)] = clock64();                      \
  } while (0)

// RES: the whole score row (NB <= 96 keys) is fetched from TMEM in one go and stays in registers
// for both softmax passes (one TMEM round trip per head instead of seven; 3 CTAs per SM at 128
// registers); otherwise the row is walked in 32-column chunks (4 CTAs per SM).
template <bool RES>
__global__ void __launch_bounds__(ATC_THREADS, RES ? 3 : 4)
attn_tc1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnArgs a, const int QR, const int NB,
                const int nqb, const int kv_shared, const int nitems) {
  extern __shared__ uint8_t atc_raw[];
  uint8_t* smem = atc_raw + ((1024u - (smem_u32(atc_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* stage_free = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_full = bars + 6;
  uint64_t* o_free = bars + 7;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t q_bytes = static_cast<uint32_t>(QR) * 128u, k_bytes = static_cast<uint32_t>(NB) * 128u;
  const uint32_t stage_bytes = q_bytes + k_bytes * (kv_shared ? 1u : 2u);
  uint8_t* stage0 = smem + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmQ);
      prefetch_tmap(&tmK);
      prefetch_tmap(&tmV);
      mbar_init(&kv_full[0], 1);
      mbar_init(&kv_full[1], 1);
      mbar_init(&stage_free[0], 1);
      mbar_init(&stage_free[1], 1);
      mbar_init(s_full, 1);
      mbar_init(p_ready, 4);
      mbar_init(o_full, 1);
      mbar_init(o_free, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 128);
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t tmem_o = tmem + 64;
  const int G = static_cast<int>(gridDim.x);

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idS = umma_idesc_bf16(128, NB);
      const uint32_t idO = umma_idesc_bf16(128, 64) | (1u << 16);   // B operand MN-major (rows = keys)
      auto issue_loads = [&](int w, int s) {
        const int pair = w & 3, qb = (w >> 2) % nqb, b = (w >> 2) / nqb;
        uint8_t* st = stage0 + static_cast<size_t>(s) * stage_bytes;
        mbar_expect_tx(&kv_full[s], stage_bytes);
        tma_load_2d(st, &tmQ, pair * 64, b * a.Lq + qb * 128, &kv_full[s]);
        tma_load_2d(st + q_bytes, &tmK, pair * 64, b * a.Lk, &kv_full[s]);
        if (!kv_shared) tma_load_2d(st + q_bytes + k_bytes, &tmV, pair * 64, b * a.Lk, &kv_full[s]);
      };
      if (static_cast<int>(blockIdx.x) < nitems) issue_loads(blockIdx.x, 0);
      uint32_t n = 0;   // score tiles issued so far
      int i = 0;
      for (int w = blockIdx.x; w < nitems; w += G, ++i) {
        const int s = i & 1;
        const uint32_t st_u = smem_u32(stage0 + static_cast<size_t>(s) * stage_bytes);
        const uint32_t sK_u = st_u + q_bytes, sV_u = kv_shared ? sK_u : sK_u + k_bytes;
        ATC_TRACE(0, i, 0);
        mbar_wait(&kv_full[s], (i >> 1) & 1);
        tc_fence_after();
        ATC_TRACE(0, i, 1);
        for (int h = 0; h < 2; ++h) {
          if (n > 0) {   // O of the previous head sits in the dead score columns [64,128): read out?
            mbar_wait(o_free, (n - 1) & 1u);
            tc_fence_after();
          }
          const uint64_t dq = umma_desc_sw128(st_u) + 4 * h;   // head h = 64-byte half of the 128-byte row
          const uint64_t dk = umma_desc_sw128(sK_u) + 4 * h;
          umma_bf16(tmem, dq, dk, idS, 0u);
          umma_bf16(tmem, dq + 2, dk + 2, idS, 1u);
          umma_commit(s_full);
          ATC_TRACE(0, i, 2 + 4 * h);
          if (h == 0 && w + G < nitems) {   // next item's operands stream in behind this item's softmax
            if (i >= 1) mbar_wait(&stage_free[s ^ 1], ((i - 1) >> 1) & 1);
            issue_loads(w + G, s ^ 1);
          }
          ATC_TRACE(0, i, 3 + 4 * h);
          mbar_wait(p_ready, n & 1u);
          tc_fence_after();
          ATC_TRACE(0, i, 4 + 4 * h);
          const uint64_t dv = umma_desc_sw128(sV_u);
          for (int j = 0; j < NB / 16; ++j)   // 16 keys = 8 packed TMEM columns = 2048 bytes of V rows
            umma_bf16_ts(tmem_o, tmem + 8 * j, dv + 128 * j, idO, j ? 1u : 0u);
          umma_commit(o_full);
          ATC_TRACE(0, i, 5 + 4 * h);
          ++n;
        }
        umma_commit(&stage_free[s]);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
    uint32_t n = 0;
    uint32_t u[32];
    int klen_next = 0;
    if (static_cast<int>(blockIdx.x) < nitems) klen_next = a.klen_src[(blockIdx.x >> 2) / nqb];
    int it = 0;
    const bool tr = warp == 1 && lane == 0;
    for (int w = blockIdx.x; w < nitems; w += G, ++it) {
      const int pair = w & 3, qb = (w >> 2) % nqb, b = (w >> 2) / nqb;
      int klen = a.kbase + klen_next;
      if (klen > a.Lk) klen = a.Lk;
      if (w + G < nitems) klen_next = a.klen_src[((w + G) >> 2) / nqb];
      const int row = qb * 128 + r;
      const bool wactive = qb * 128 + q * 32 < a.Lq;   // warp-uniform: some row of this warp exists
      for (int h = 0; h < 2; ++h, ++n) {
        if (tr) ATC_TRACE(1, it, 0 + 8 * h);
        mbar_wait(s_full, n & 1u);
        tc_fence_after();
        if (tr) ATC_TRACE(1, it, 1 + 8 * h);
        float lsum = 0.f, tsum = 0.f;
        if (wactive) {
          float mx = -INFINITY;
          uint64_t acc_all = 0ull, acc_val = 0ull;
          uint32_t pk[16];
          if constexpr (RES) {
            uint32_t u1[32], u2[32];
            tmem_ld32(tmem + lane_addr, u);
            if (NB >= 64) tmem_ld32(tmem + lane_addr + 32, u1);
            else if (NB > 32) tmem_ld16(tmem + lane_addr + 32, u1);
            if (NB >= 96) tmem_ld32(tmem + lane_addr + 64, u2);
            else if (NB > 64) tmem_ld16(tmem + lane_addr + 64, u2);
            tmem_ld_wait();
            mx = chunk_max<32>(u, klen, mx);
            if (NB > 32) mx = chunk_max<32>(u1, klen - 32, mx);
            if (NB > 64) mx = chunk_max<32>(u2, klen - 64, mx);
            const float noff = -(mx == -INFINITY ? 0.f : mx) * sc;
            if (tr) ATC_TRACE(1, it, 2 + 8 * h);
            chunk_softmax<32>(u, pk, klen, a.v_first, sc, noff, acc_all, acc_val);
            tmem_st16(tmem + lane_addr, pk);
            if (NB > 32) {
              chunk_softmax<32>(u1, pk, klen - 32, a.v_first - 32, sc, noff, acc_all, acc_val);
              if (NB >= 64) tmem_st16(tmem + lane_addr + 16, pk);
              else tmem_st8(tmem + lane_addr + 16, pk);
            }
            if (NB > 64) {
              chunk_softmax<32>(u2, pk, klen - 64, a.v_first - 64, sc, noff, acc_all, acc_val);
              if (NB >= 96) tmem_st16(tmem + lane_addr + 32, pk);
              else tmem_st8(tmem + lane_addr + 32, pk);
            }
          } else {
            int c0 = 0;
            for (; c0 + 32 <= NB; c0 += 32) {
              if (klen - c0 <= 0) break;
              tmem_ld32(tmem + lane_addr + c0, u);
              tmem_ld_wait();
              mx = chunk_max<32>(u, klen - c0, mx);
            }
            if ((NB & 16) && klen - (NB - 16) > 0) {
              tmem_ld16(tmem + lane_addr + NB - 16, u);
              tmem_ld_wait();
              mx = chunk_max<16>(u, klen - (NB - 16), mx);
            }
            const float noff = -(mx == -INFINITY ? 0.f : mx) * sc;
            if (tr) ATC_TRACE(1, it, 2 + 8 * h);
            for (c0 = 0; c0 + 32 <= NB; c0 += 32) {
              const int nv = klen - c0;
              if (nv > 0) {
                tmem_ld32(tmem + lane_addr + c0, u);
                tmem_ld_wait();
              }
              chunk_softmax<32>(u, pk, nv, a.v_first - c0, sc, noff, acc_all, acc_val);
              tmem_st16(tmem + lane_addr + (c0 >> 1), pk);
            }
            if (NB & 16) {
              c0 = NB - 16;
              const int nv = klen - c0;
              if (nv > 0) {
                tmem_ld16(tmem + lane_addr + c0, u);
                tmem_ld_wait();
              }
              chunk_softmax<16>(u, pk, nv, a.v_first - c0, sc, noff, acc_all, acc_val);
              tmem_st8(tmem + lane_addr + (c0 >> 1), pk);
            }
          }
          tmem_st_wait();
          float x, y;
          f2_unpack(acc_all, x, y);
          lsum = x + y;
          f2_unpack(acc_val, x, y);
          tsum = x + y;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (tr) ATC_TRACE(1, it, 3 + 8 * h);

        mbar_wait(o_full, n & 1u);
        tc_fence_after();
        if (tr) ATC_TRACE(1, it, 4 + 8 * h);
        if (wactive) {
          tmem_ld32(tmem_o + lane_addr + 32 * h, u);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        if (wactive && row < a.Lq) {
          const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
          const int hh = pair * 2 + h;
          const size_t grow = static_cast<size_t>(b) * a.Lq + row;
          uint4* dst = reinterpret_cast<uint4*>(a.out + grow * 256 + hh * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(u[8 * i + 0]) * inv, __uint_as_float(u[8 * i + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(u[8 * i + 2]) * inv, __uint_as_float(u[8 * i + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(u[8 * i + 4]) * inv, __uint_as_float(u[8 * i + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(u[8 * i + 6]) * inv, __uint_as_float(u[8 * i + 7]) * inv);
            dst[i] = v;
          }
          if (a.tsum) a.tsum[static_cast<size_t>(hh) * a.B * a.Lq + grow] = tsum * inv;
        }
        if (tr) ATC_TRACE(1, it, 5 + 8 * h);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

template <bool RES>
static int launch_attn_tc1(cudaStream_t st, const AttnArgs& a, int QR, int NB, int nqb, bool kv_shared) {
  const int stage = (QR + NB * (kv_shared ? 1 : 2)) * 128;
  // the score MMA always reads 128 A rows from the start of a stage's Q tile: keep them inside the allocation
  const int tail = stage >= ATC_UNIT ? 0 : ATC_UNIT - stage;
  const int smem = 1024 + 2 * stage + tail + 1024;
  static thread_local int smem_set = 0;
  if (smem > smem_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(attn_tc1_kernel<RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    FVTG_CUDA_OK(cudaFuncSetAttribute(attn_tc1_kernel<RES>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
    smem_set = smem;
  }
  // resident CTAs per SM: registers (launch bounds) / 128 TMEM columns each / shared memory (+1 KB reserved per CTA)
  int per_sm = RES ? 3 : 4;
  const int by_smem = (227 * 1024) / (smem + 1024);
  if (by_smem < per_sm) per_sm = by_smem < 1 ? 1 : by_smem;
  const int nitems = a.B * nqb * 4;
  const int grid = nitems < per_sm * sm_count() ? nitems : per_sm * sm_count();
  CUtensorMap tq, tk, tv;
  FVTG_TRY(make_tmap_bf16(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, 256, a.ldq, QR, 64));
  FVTG_TRY(make_tmap_bf16(&tk, a.k, static_cast<uint64_t>(a.B) * a.Lk, 256, a.ldk, NB, 64));
  FVTG_TRY(make_tmap_bf16(&tv, a.v, static_cast<uint64_t>(a.B) * a.Lk, 256, a.ldv, NB, 64));
  ProfScope prof(st, PC_ATTN);
  FVTG_CUDA_OK(launch_pdl(attn_tc1_kernel<RES>, dim3(grid), dim3(ATC_THREADS), smem, st, tq, tk, tv, a, QR, NB, nqb,
                          kv_shared ? 1 : 0, nitems));
  FVTG_LAUNCH_CHECK("attn_tc1_kernel");
  return FVTG_OK;
}

template <bool MULTI>
static int launch_attn_tc_t(cudaStream_t st, const AttnArgs& a, int QR, int NB, int nkb, int nqb,
                            bool kv_shared) {
  const int smem = 3 * ATC_UNIT + 128 + 1024;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  FVTG_TRY(make_tmap_bf16(&tq, a.q, static_cast<uint64_t>(a.B) * a.Lq, 256, a.ldq, QR, 64));
  FVTG_TRY(make_tmap_bf16(&tk, a.k, static_cast<uint64_t>(a.B) * a.Lk, 256, a.ldk, NB, 64));
  FVTG_TRY(make_tmap_bf16(&tv, a.v, static_cast<uint64_t>(a.B) * a.Lk, 256, a.ldv, NB, 64));
  ProfScope prof(st, PC_ATTN);
  FVTG_CUDA_OK(launch_pdl(attn_tc_kernel<MULTI>, dim3(a.B * nqb * 4), dim3(ATC_THREADS), smem, st, tq, tk, tv,
                          a, QR, NB, nkb, nqb, kv_shared ? 1 : 0));
  FVTG_LAUNCH_CHECK("attn_tc_kernel");
  return FVTG_OK;
}

int launch_attention_tc(cudaStream_t st, const AttnArgs& a_in) {
  if (a_in.B <= 0) return FVTG_OK;
  AttnArgs a = a_in;
  {  // debug: FVTG_ATTN_TRACE_T2V=1 keeps the clock64 trace of cross-attention launches only
    static const bool t2v_only = [] { const char* e = getenv("FVTG_ATTN_TRACE_T2V"); return e && atoi(e) != 0; }();
    if (t2v_only && !a.tsum) a.trace = nullptr;
  }
  const int nqb = (a.Lq + 127) / 128;
  const int QR = a.Lq >= 128 ? 128 : round_up(a.Lq, 8);
  const bool kv_shared = (a.k == a.v) && (a.ldk == a.ldv);
  static const bool persistent = [] { const char* e = getenv("FVTG_ATTN_PERSIST"); return !e || atoi(e) != 0; }();
  if (a.Lk <= 96 && persistent) return launch_attn_tc1<true>(st, a, QR, round_up(a.Lk, 16), nqb, kv_shared);
  if (a.Lk <= 128 && persistent) return launch_attn_tc1<false>(st, a, QR, round_up(a.Lk, 16), nqb, kv_shared);
  if (a.Lk <= 128) return launch_attn_tc_t<false>(st, a, QR, round_up(a.Lk, 16), 1, nqb, kv_shared);
  return launch_attn_tc_t<true>(st, a, QR, 128, (a.Lk + 127) / 128, nqb, kv_shared);
}

}  // namespace fvtg
