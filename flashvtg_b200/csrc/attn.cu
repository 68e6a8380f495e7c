// Per-(video, head) attention for the three attention flavours of the path:
//   dummy-token encoder / self-attention encoder : nn.MultiheadAttention core (softmax(QK^T/sqrt(32)) V,
//       key padding by true length)                      [transformer.py:413-415]
//   adaptive cross-attention : no projections, softmax over [dummies ‖ text] keys, value sum over
//       the text keys only, per-row probability mass on text accumulated for t2vattnvalues
//                                                         [crossattention.py:287-396, model.py:215]
// Sequences are 4..~1000 tokens with head_dim 32 (2-6 % of the FLOPs): legacy warp-level
// mma.sync m16n8k16 with an online softmax; the dense d_model contractions go through gemm.cu.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int ATT_THREADS = 128;
constexpr int ATT_KS = 40;  // smem key row stride (bf16): 32 + 8 pad -> conflict-free fragment loads

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const AttnArgs a, const int Lkpad) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* Ks = reinterpret_cast<bf16*>(att_smem);  // [Lkpad][40]
  const int vs = Lkpad + 8;
  bf16* Vt = Ks + static_cast<size_t>(Lkpad) * ATT_KS;  // [32][Lkpad + 8]
  const int b = blockIdx.x >> 3, h = blockIdx.x & 7;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  int klen = a.kbase + a.klen_src[b];
  if (klen > a.Lk) klen = a.Lk;

  for (int idx = threadIdx.x; idx < Lkpad * 4; idx += ATT_THREADS) {
    const int j = idx >> 2, sg = idx & 3;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (j < klen) {
      const size_t r = static_cast<size_t>(b) * a.Lk + j;
      kv = *reinterpret_cast<const uint4*>(a.k + r * a.ldk + h * 32 + sg * 8);
      if (j >= a.v_first) vv = *reinterpret_cast<const uint4*>(a.v + r * a.ldv + h * 32 + sg * 8);
    }
    *reinterpret_cast<uint4*>(Ks + j * ATT_KS + sg * 8) = kv;
    const bf16* ve = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
    for (int e = 0; e < 8; ++e) Vt[(sg * 8 + e) * vs + j] = ve[e];
  }
  __syncthreads();

  const float sc = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
  for (int mt = warp; mt * 16 < a.Lq; mt += ATT_THREADS / 32) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    uint32_t aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const bf16* q0 = a.q + (static_cast<size_t>(b) * a.Lq + r0) * a.ldq + h * 32 + ks * 16 + 2 * t;
      const bf16* q1 = a.q + (static_cast<size_t>(b) * a.Lq + r1) * a.ldq + h * 32 + ks * 16 + 2 * t;
      aq[ks][0] = r0 < a.Lq ? *reinterpret_cast<const uint32_t*>(q0) : 0u;
      aq[ks][1] = r1 < a.Lq ? *reinterpret_cast<const uint32_t*>(q1) : 0u;
      aq[ks][2] = r0 < a.Lq ? *reinterpret_cast<const uint32_t*>(q0 + 8) : 0u;
      aq[ks][3] = r1 < a.Lq ? *reinterpret_cast<const uint32_t*>(q1 + 8) : 0u;
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f, ts0 = 0.f, ts1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[i][e] = 0.f;

    for (int kc = 0; kc < klen; kc += 64) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const bf16* kr = Ks + (kc + nt * 8 + g) * ATT_KS + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
          mma16816(s[nt], aq[ks], b0, b1);
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kc + nt * 8 + 2 * t + (e & 1);
          const float v = key < klen ? s[nt][e] * sc : -INFINITY;
          s[nt][e] = v;
          if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
      const float al0 = exp2f(m0 - n0), al1 = exp2f(m1 - n1);
      m0 = n0; m1 = n1;
      float rs0 = 0.f, rs1 = 0.f, rt0 = 0.f, rt1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kc + nt * 8 + 2 * t + (e & 1);
          const float p = exp2f(s[nt][e] - (e < 2 ? n0 : n1));
          s[nt][e] = p;
          if (e < 2) { rs0 += p; if (key >= a.v_first) rt0 += p; }
          else       { rs1 += p; if (key >= a.v_first) rt1 += p; }
        }
      }
      l0 = l0 * al0 + rs0; l1 = l1 * al1 + rs1;
      ts0 = ts0 * al0 + rt0; ts1 = ts1 * al1 + rt1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int nt2 = 0; nt2 < 4; ++nt2) {
          const bf16* vr = Vt + (nt2 * 8 + g) * vs + kc + kk * 16 + 2 * t;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vr);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vr + 8);
          mma16816(o[nt2], pa, b0, b1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    ts0 += __shfl_xor_sync(0xffffffffu, ts0, 1); ts0 += __shfl_xor_sync(0xffffffffu, ts0, 2);
    ts1 += __shfl_xor_sync(0xffffffffu, ts1, 1); ts1 += __shfl_xor_sync(0xffffffffu, ts1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    if (r0 < a.Lq) {
      bf16* dst = a.out + (static_cast<size_t>(b) * a.Lq + r0) * 256 + h * 32 + 2 * t;
#pragma unroll
      for (int nt2 = 0; nt2 < 4; ++nt2)
        *reinterpret_cast<uint32_t*>(dst + nt2 * 8) = pack_bf16(o[nt2][0] * i0, o[nt2][1] * i0);
      if (a.tsum && t == 0)
        a.tsum[static_cast<size_t>(h) * a.B * a.Lq + static_cast<size_t>(b) * a.Lq + r0] = ts0 * i0;
    }
    if (r1 < a.Lq) {
      bf16* dst = a.out + (static_cast<size_t>(b) * a.Lq + r1) * 256 + h * 32 + 2 * t;
#pragma unroll
      for (int nt2 = 0; nt2 < 4; ++nt2)
        *reinterpret_cast<uint32_t*>(dst + nt2 * 8) = pack_bf16(o[nt2][2] * i1, o[nt2][3] * i1);
      if (a.tsum && t == 0)
        a.tsum[static_cast<size_t>(h) * a.B * a.Lq + static_cast<size_t>(b) * a.Lq + r1] = ts1 * i1;
    }
  }
}

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
  pdl_wait();
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
    // Key rows >= klen are zero in shared memory, so their scores are exactly 0: the stabilising
    // offset may include them (softmax is shift invariant) and needs no masking; their probabilities
    // are forced to 0 below with two compares per tile against this lane's column limit - branch-free,
    // no per-tile control flow for the register allocator to patch up with moves.
    float mx0 = fmaxf(fmaxf(s[0][0], s[0][1]), 0.f), mx1 = fmaxf(fmaxf(s[0][2], s[0][3]), 0.f);
#pragma unroll
    for (int nt = 1; nt < NT; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float off0 = mx0 * sc, off1 = mx1 * sc;
    const int lim = klen - 2 * t;          // column c of tile nt is a valid key  <=>  nt * 8 + c < lim
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
    static const bool tc = [] { const char* e = getenv("FVTG_ATTN_TC"); return !e || atoi(e) != 0; }();
    if (tc) return launch_attention_tc(st, a);
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
  const int Lkpad = round_up(a.Lk, 64);
  const size_t smem = static_cast<size_t>(Lkpad) * ATT_KS * 2 + 32 * static_cast<size_t>(Lkpad + 8) * 2;
  if (smem > 220 * 1024) return fail(FVTG_EINVAL, "attention: %d keys exceed shared memory", a.Lk);
  static thread_local size_t set = 0;
  if (smem > set) {
    FVTG_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    set = smem;
  }
  ProfScope prof(st, PC_ATTN);
  attention_kernel<<<a.B * 8, ATT_THREADS, smem, st>>>(a, Lkpad);
  FVTG_LAUNCH_CHECK("attention_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
