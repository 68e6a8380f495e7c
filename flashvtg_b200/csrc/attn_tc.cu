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

int launch_attention_tc(cudaStream_t st, const AttnArgs& a) {
  if (a.B <= 0) return FVTG_OK;
  const int nqb = (a.Lq + 127) / 128;
  const int QR = a.Lq >= 128 ? 128 : round_up(a.Lq, 8);
  const bool kv_shared = (a.k == a.v) && (a.ldk == a.ldv);
  if (a.Lk <= 128) return launch_attn_tc_t<false>(st, a, QR, round_up(a.Lk, 16), 1, nqb, kv_shared);
  return launch_attn_tc_t<true>(st, a, QR, 128, (a.Lk + 127) / 128, nqb, kv_shared);
}

}  // namespace fvtg
