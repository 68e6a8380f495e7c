// Per-(video, head) attention for the three attention flavours of the path:
//   dummy-token encoder / self-attention encoder : nn.MultiheadAttention core (softmax(QK^T/sqrt(32)) V,
//       key padding by true length)                      [transformer.py:413-415]
//   adaptive cross-attention : no projections, softmax over [dummies ‖ text] keys, value sum over
//       the text keys only, per-row probability mass on text accumulated for t2vattnvalues
//                                                         [crossattention.py:287-396, model.py:215]
// Sequences of up to 160 keys (head_dim 32, 2-6 % of the FLOPs) run here on warp-level mma.sync
// m16n8k16 with the whole score row in registers; longer ones go to the tcgen05 kernel of attn_tc.cu
// (key blocks of 128 in TMEM).  The dense d_model contractions go through gemm.cu / layer.cu.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

// ---------------------------------------------------------------------------------------------
// Short sequences (everything QVHighlights-sized): one CTA per (video, group of 4 heads), 8 warps =
// 4 heads x 2 interleaved m-tile groups.  The 128-column slices of the video's Q / K / V rows are
// staged once with coalesced 16-byte cp.async (row pitch 272 B keeps ldmatrix conflict-free), the
// whole score row lives in registers (no online softmax), the output tile goes back through
// shared memory so that global stores are 16 bytes per lane.  ~65 KB of shared memory and <= 85
// registers keep 3 CTAs (24 warps) per SM: the kernel is issue-bound on the softmax, not on MMA
// (ncu: ~730 SASS instructions per 16-row unit before the masking was made branch-free, 40 of them HMMA).
// A persistent variant with double-buffered K / V staging, Q fragments read straight from global memory
// and 2 balanced units per warp was built and measured (commit "Attention: branch-free key/value
// masking ..."): 0.70 ms/step against 0.56 for this kernel - its exposed L2 latency on the Q fragments
// and the lower warp count per SM cost more than the hidden staging saves - so it was dropped.
constexpr int AV_PITCH = 136;  // bf16 elements per staged row (4 heads x 32 + 8 pad)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NT, bool KV_SHARED>
__global__ void __launch_bounds__(256, (NT <= 10 ? 3 : 2))
attn_video_kernel(const AttnArgs a, const int LqPad) {
  constexpr int LkPad = NT * 8;
  extern __shared__ __align__(16) uint8_t av_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(av_smem);
  bf16* sK = sQ + static_cast<size_t>(LqPad) * AV_PITCH;
  bf16* sV = KV_SHARED ? sK : sK + static_cast<size_t>(LkPad) * AV_PITCH;
  const int b = blockIdx.x >> 1, hg = blockIdx.x & 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  pdl_launch_dependents();
  if (a.q_flags) {
    // tile-granular dependency on the producing layer kernel (see AttnArgs::q_flags).  The spin is bounded, and a
    // CTA that runs out of patience falls back to the grid-wide wait: never a wrong result, never a hung GPU.
    if (tid == 0) {
      const int t0 = (b * a.Lq) / a.q_tile_rows, t1 = (b * a.Lq + a.Lq - 1) / a.q_tile_rows;
      bool late = false;
      for (int t = t0; t <= t1 && !late; ++t) {
        int spins = 0;
        while (ld_acquire_gpu(a.q_flags + t) - a.q_epoch < 0) {
          if (++spins >= (1 << 22)) { late = true; break; }
          __nanosleep(64);
        }
      }
      if (late) pdl_wait();
    }
    __syncthreads();
  } else {
    pdl_wait();
  }
#define AT_TRACE(ev) do { if (a.trace && tid == 0 && (blockIdx.x == 0 || blockIdx.x == 1000)) a.trace[3072 + (blockIdx.x ? 8 : 0) + (ev)] = clock64(); } while (0)
  AT_TRACE(0);
  int klen = a.kbase + a.klen_src[b];
  if (klen > a.Lk) klen = a.Lk;
  const int col0 = hg * 128;

  // ---- stage: 16 x 16-byte chunks per row ---------------------------------------------------
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int idx = tid; idx < LqPad * 16; idx += 256) {
    const int r = idx >> 4, c = idx & 15;
    bf16* dst = sQ + r * AV_PITCH + c * 8;
    if (r < a.Lq)
      cp_async16(smem_u32(dst), a.q + (static_cast<size_t>(b) * a.Lq + r) * a.ldq + col0 + c * 8);
    else
      *reinterpret_cast<uint4*>(dst) = zero4;
  }
  for (int idx = tid; idx < LkPad * 16; idx += 256) {
    const int r = idx >> 4, c = idx & 15;
    bf16* dk = sK + r * AV_PITCH + c * 8;
    const size_t grow = static_cast<size_t>(b) * a.Lk + r;
    if (r < klen) cp_async16(smem_u32(dk), a.k + grow * a.ldk + col0 + c * 8);
    else *reinterpret_cast<uint4*>(dk) = zero4;
    if (!KV_SHARED) {
      bf16* dv = sV + r * AV_PITCH + c * 8;
      if (r < klen) cp_async16(smem_u32(dv), a.v + grow * a.ldv + col0 + c * 8);
      else *reinterpret_cast<uint4*>(dv) = zero4;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  AT_TRACE(1);

  const int hl = warp & 3;            // head within the group
  const int h = hg * 4 + hl;
  const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
  for (int mt = warp >> 2; mt * 16 < a.Lq; mt += 2) {
    uint32_t aq[2][4];
    {
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t base = smem_u32(sQ + row * AV_PITCH + hl * 32 + (lane >> 4) * 8);
      ldmatrix_x4(aq[0], base);
      ldmatrix_x4(aq[1], base + 32);
    }
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kb[4];
      ldmatrix_x4(kb, smem_u32(sK + (nt * 8 + (lane & 7)) * AV_PITCH + hl * 32 + (lane >> 3) * 8));
      mma16816(s[nt], aq[0], kb[0], kb[1]);
      mma16816(s[nt], aq[1], kb[2], kb[3]);
    }
    // The stabilising offset is the maximum over the VALID keys only (a row whose valid scores all sit
    // far below zero must not be flushed by the zero scores of the padded key rows).  Whether a key tile
    // lies below klen is warp-uniform, so only the tile that straddles klen pays for selects; the
    // probabilities of padded keys are forced to 0 below with two compares per tile against this lane's
    // column limit - branch-free, no per-tile control flow in the exp loop.
    const int lim = klen - 2 * t;          // column c of tile nt is a valid key  <=>  nt * 8 + c < lim
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 + 8 <= klen) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      } else if (nt * 8 < klen) {
        const bool k0 = nt * 8 < lim, k1 = nt * 8 + 1 < lim;
        mx0 = fmaxf(mx0, fmaxf(k0 ? s[nt][0] : -INFINITY, k1 ? s[nt][1] : -INFINITY));
        mx1 = fmaxf(mx1, fmaxf(k0 ? s[nt][2] : -INFINITY, k1 ? s[nt][3] : -INFINITY));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float off0 = (mx0 == -INFINITY ? 0.f : mx0) * sc, off1 = (mx1 == -INFINITY ? 0.f : mx1) * sc;
    const int vlim = a.v_first - 2 * t;    // ... carries a value              <=>  nt * 8 + c >= vlim
    float l0 = 0.f, l1 = 0.f, ts0 = 0.f, ts1 = 0.f;
    if (a.v_first > 0) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const bool k0 = nt * 8 < lim, k1 = nt * 8 + 1 < lim;
        const bool v0 = nt * 8 >= vlim, v1 = nt * 8 + 1 >= vlim;
        float p0 = ex2_approx(fmaf(s[nt][0], sc, -off0));
        float p1 = ex2_approx(fmaf(s[nt][1], sc, -off0));
        float p2 = ex2_approx(fmaf(s[nt][2], sc, -off1));
        float p3 = ex2_approx(fmaf(s[nt][3], sc, -off1));
        p0 = k0 ? p0 : 0.f; p1 = k1 ? p1 : 0.f; p2 = k0 ? p2 : 0.f; p3 = k1 ? p3 : 0.f;
        l0 += p0 + p1;
        l1 += p2 + p3;
        s[nt][0] = v0 ? p0 : 0.f; s[nt][1] = v1 ? p1 : 0.f;     // dummies absorb mass, carry no value
        s[nt][2] = v0 ? p2 : 0.f; s[nt][3] = v1 ? p3 : 0.f;
        ts0 += s[nt][0] + s[nt][1];
        ts1 += s[nt][2] + s[nt][3];
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const bool k0 = nt * 8 < lim, k1 = nt * 8 + 1 < lim;
        const float p0 = ex2_approx(fmaf(s[nt][0], sc, -off0));
        const float p1 = ex2_approx(fmaf(s[nt][1], sc, -off0));
        const float p2 = ex2_approx(fmaf(s[nt][2], sc, -off1));
        const float p3 = ex2_approx(fmaf(s[nt][3], sc, -off1));
        s[nt][0] = k0 ? p0 : 0.f; s[nt][1] = k1 ? p1 : 0.f;
        s[nt][2] = k0 ? p2 : 0.f; s[nt][3] = k1 ? p3 : 0.f;
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      ts0 = l0;
      ts1 = l1;
    }
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const int krow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t vb = smem_u32(sV + krow * AV_PITCH + hl * 32 + (lane >> 4) * 8);
      uint32_t v0[4], v1[4];
      ldmatrix_x4_trans(v0, vb);
      ldmatrix_x4_trans(v1, vb + 32);
      mma16816(o[0], pa, v0[0], v0[1]);
      mma16816(o[1], pa, v0[2], v0[3]);
      mma16816(o[2], pa, v1[0], v1[1]);
      mma16816(o[3], pa, v1[2], v1[3]);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // the Q tile of this (m-tile, head) is dead: reuse its slots for the output
    __syncwarp();
    bf16* d0 = sQ + r0 * AV_PITCH + hl * 32 + 2 * t;
    bf16* d1 = sQ + r1 * AV_PITCH + hl * 32 + 2 * t;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      *reinterpret_cast<uint32_t*>(d0 + n2 * 8) = pack_bf16(o[n2][0] * i0, o[n2][1] * i0);
      *reinterpret_cast<uint32_t*>(d1 + n2 * 8) = pack_bf16(o[n2][2] * i1, o[n2][3] * i1);
    }
    if (a.tsum) {
      ts0 += __shfl_xor_sync(0xffffffffu, ts0, 1); ts0 += __shfl_xor_sync(0xffffffffu, ts0, 2);
      ts1 += __shfl_xor_sync(0xffffffffu, ts1, 1); ts1 += __shfl_xor_sync(0xffffffffu, ts1, 2);
      if (t == 0) {
        // per-layer slot: plain stores, summed by the saliency kernel (no read-modify-write chain)
        float* ts = a.tsum + static_cast<size_t>(h) * a.B * a.Lq + static_cast<size_t>(b) * a.Lq;
        if (r0 < a.Lq) ts[r0] = ts0 * i0;
        if (r1 < a.Lq) ts[r1] = ts1 * i1;
      }
    }
  }
  AT_TRACE(2);
  __syncthreads();
  AT_TRACE(3);
  for (int idx = tid; idx < a.Lq * 16; idx += 256) {
    const int r = idx >> 4, c = idx & 15;
    *reinterpret_cast<uint4*>(a.out + (static_cast<size_t>(b) * a.Lq + r) * 256 + col0 + c * 8) =
        *reinterpret_cast<const uint4*>(sQ + r * AV_PITCH + c * 8);
  }
  AT_TRACE(4);
}

template <int NT, bool KV_SHARED>
static int launch_attn_video(cudaStream_t st, const AttnArgs& a, int LqPad, size_t smem) {
  static thread_local size_t set = 0;
  if (smem > set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(attn_video_kernel<NT, KV_SHARED>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = smem;
  }
  ProfScope prof(st, PC_ATTN);
  FVTG_CUDA_OK(launch_pdl(attn_video_kernel<NT, KV_SHARED>, dim3(a.B * 2), dim3(256), smem, st, a, LqPad));
  FVTG_LAUNCH_CHECK("attn_video_kernel");
  return FVTG_OK;
}

int launch_attention(cudaStream_t st, const AttnArgs& a) {
  if (a.B <= 0) return FVTG_OK;
  {
    // tcgen05 kernel (attn_tc.cu): always for Lk > 160; FVTG_ATTN_TC=1 forces it for the short shapes too
    static const int tc = [] { const char* e = getenv("FVTG_ATTN_TC"); return e ? atoi(e) : -1; }();
    if (tc > 0 || (tc < 0 && a.Lk > 160)) return launch_attention_tc(st, a);
  }
  {
    const int LqPad = round_up(a.Lq, 16);
    const int nt = a.Lk <= 80 ? 10 : 20;
    const bool shared_kv = (a.k == a.v) && (a.ldk == a.ldv);
    const size_t smem_v = static_cast<size_t>(LqPad + nt * 8 * (shared_kv ? 1 : 2)) * AV_PITCH * 2;
    const bool aligned = (a.ldq % 8 == 0) && (a.ldk % 8 == 0) && (a.ldv % 8 == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) |
                           reinterpret_cast<uintptr_t>(a.v)) & 15) == 0;
    if (a.Lk <= 160 && smem_v <= 200 * 1024 && aligned) {
      if (nt == 10) return shared_kv ? launch_attn_video<10, true>(st, a, LqPad, smem_v)
                                     : launch_attn_video<10, false>(st, a, LqPad, smem_v);
      return shared_kv ? launch_attn_video<20, true>(st, a, LqPad, smem_v)
                       : launch_attn_video<20, false>(st, a, LqPad, smem_v);
    }
  }
  // everything else (long videos: TACoS, Charades-STA VGG; odd layouts) runs on the tcgen05 kernel
  return launch_attention_tc(st, a);
}

}  // namespace fvtg
