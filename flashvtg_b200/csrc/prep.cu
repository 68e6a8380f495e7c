// HBM-bound staging kernels: coalesced vector loads, warp-shuffle reductions, one pass over the
// big fp32 inputs (the only place the 0.75-0.9 MB/video of features is read).
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

// One warp per row.  The row is staged in shared memory (one HBM read), then mean / variance are
// taken two-pass in fp32 exactly like torch's LayerNorm, then normalised and written as bf16.
constexpr int LNC_WARPS = 4;

template <int VEC>
__global__ void __launch_bounds__(LNC_WARPS * 32)
ln_cast_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
               const float* __restrict__ beta, bf16* __restrict__ out, int rows, int dim,
               int dim_pad) {
  extern __shared__ float s_row[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LNC_WARPS + warp;
  if (row >= rows) return;
  float* buf = s_row + static_cast<size_t>(warp) * dim_pad;
  const float* src = in + static_cast<size_t>(row) * dim;
  float sum = 0.f;
  if (VEC == 2) {
    const float2* s2 = reinterpret_cast<const float2*>(src);
    const int n2 = dim >> 1;
#pragma unroll 4
    for (int i = lane; i < n2; i += 32) {
      const float2 v = __ldg(s2 + i);
      reinterpret_cast<float2*>(buf)[i] = v;
      sum += v.x + v.y;
    }
  } else {
    for (int i = lane; i < dim; i += 32) {
      const float v = __ldg(src + i);
      buf[i] = v;
      sum += v;
    }
  }
  sum = warp_sum(sum);
  const float mean = sum / static_cast<float>(dim);
  __syncwarp();
  float sq = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float d = buf[i] - mean;
    sq += d * d;
  }
  sq = warp_sum(sq);
  const float rstd = rsqrtf(sq / static_cast<float>(dim) + 1e-5f);
  bf16* dst = out + static_cast<size_t>(row) * dim_pad;
  if (VEC == 2) {
    const int n2 = dim >> 1;
    for (int i = lane; i < (dim_pad >> 1); i += 32) {
      float2 y = make_float2(0.f, 0.f);
      if (i < n2) {
        const float2 x = reinterpret_cast<const float2*>(buf)[i];
        const float2 g = __ldg(reinterpret_cast<const float2*>(gamma) + i);
        const float2 b = __ldg(reinterpret_cast<const float2*>(beta) + i);
        y.x = (x.x - mean) * rstd * g.x + b.x;
        y.y = (x.y - mean) * rstd * g.y + b.y;
      }
      reinterpret_cast<__nv_bfloat162*>(dst)[i] = __floats2bfloat162_rn(y.x, y.y);
    }
  } else {
    for (int i = lane; i < dim_pad; i += 32) {
      float y = 0.f;
      if (i < dim) y = (buf[i] - mean) * rstd * __ldg(gamma + i) + __ldg(beta + i);
      dst[i] = __float2bfloat16(y);
    }
  }
}

int launch_ln_cast(cudaStream_t st, const float* in, const float* gamma, const float* beta,
                   bf16* out, int rows, int dim, int dim_pad) {
  if (rows <= 0) return FVTG_OK;
  const size_t smem = static_cast<size_t>(LNC_WARPS) * dim_pad * sizeof(float);
  if (smem > 200 * 1024) return fail(FVTG_EINVAL, "ln_cast: feature dim %d too large", dim);
  const int grid = (rows + LNC_WARPS - 1) / LNC_WARPS;
  ProfScope prof(st, PC_LNCAST);
  const bool vec2 = (dim % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & 7) == 0) &&
                    ((reinterpret_cast<uintptr_t>(gamma) & 7) == 0) &&
                    ((reinterpret_cast<uintptr_t>(beta) & 7) == 0);
  if (vec2) {
    static thread_local size_t set2 = 0;
    if (smem > set2) {
      FVTG_CUDA_OK(cudaFuncSetAttribute(ln_cast_kernel<2>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set2 = smem;
    }
    ln_cast_kernel<2><<<grid, LNC_WARPS * 32, smem, st>>>(in, gamma, beta, out, rows, dim, dim_pad);
  } else {
    static thread_local size_t set1 = 0;
    if (smem > set1) {
      FVTG_CUDA_OK(cudaFuncSetAttribute(ln_cast_kernel<1>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set1 = smem;
    }
    ln_cast_kernel<1><<<grid, LNC_WARPS * 32, smem, st>>>(in, gamma, beta, out, rows, dim, dim_pad);
  }
  FVTG_LAUNCH_CHECK("ln_cast_kernel");
  return FVTG_OK;
}

// pos[b*Lv + i][c]: e = (i+1) / (len + 1e-6) * 2pi ; even c: sin(e / w_c), odd c: cos(e / w_c),
// w_c = 10000^(2*floor(c/2)/256).
__global__ void posenc_kernel(float* __restrict__ pos, const int* __restrict__ vlen, int B, int Lv) {
  const int row = blockIdx.x;
  const int c = threadIdx.x;
  const int b = row / Lv, i = row - b * Lv;
  const int len = vlen[b];
  float v = 0.f;
  if (i < len) {
    const float e = static_cast<float>(i + 1) / (static_cast<float>(len) + 1e-6f) * 6.283185307179586f;
    const float w = powf(10000.f, static_cast<float>(2 * (c >> 1)) / 256.f);
    const float a = e / w;
    v = (c & 1) ? cosf(a) : sinf(a);
  }
  pos[blk_off(row, c)] = v;
}

int launch_posenc(cudaStream_t st, float* pos, const int* vlen, int B, int Lv) {
  if (B * Lv <= 0) return FVTG_OK;
  ProfScope prof(st, PC_OTHER);
  posenc_kernel<<<B * Lv, 256, 0, st>>>(pos, vlen, B, Lv);
  FVTG_LAUNCH_CHECK("posenc_kernel");
  return FVTG_OK;
}

__global__ void fill_dummy_kernel(const float* __restrict__ dtok, const float* __restrict__ dpos,
                                  float* __restrict__ X, bf16* __restrict__ Xb,
                                  bf16* __restrict__ XPb, float* __restrict__ pos_d, int B, int S,
                                  int nd) {
  const int c = threadIdx.x;
  const int j = blockIdx.x;  // 0..S-1
  if (blockIdx.y == 0) pos_d[j * 256 + c] = j < nd ? dpos[j * 256 + c] : 0.f;
  if (j >= nd) return;
  const float t = dtok[j * 256 + c], p = dpos[j * 256 + c];
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const size_t o = (static_cast<size_t>(b) * S + j) * 256 + c;
    X[blk_off(static_cast<size_t>(b) * S + j, c)] = t;
    Xb[o] = __float2bfloat16(t);
    XPb[o] = __float2bfloat16(t + p);
  }
}

int launch_fill_dummy(cudaStream_t st, const float* dtok, const float* dpos, float* X, bf16* Xb,
                      bf16* XPb, float* pos_d, int B, int S, int nd) {
  dim3 grid(S, B < 64 ? B : 64);
  ProfScope prof(st, PC_OTHER);
  fill_dummy_kernel<<<grid, 256, 0, st>>>(dtok, dpos, X, Xb, XPb, pos_d, B, S, nd);
  FVTG_LAUNCH_CHECK("fill_dummy_kernel");
  return FVTG_OK;
}

// thread = (chain row, 8 channels)
__global__ void level0_kernel(const float* __restrict__ F, bf16* __restrict__ chain0,
                              bf16* __restrict__ H1, bf16* __restrict__ H2, int B, int Lv,
                              PyrGeo geo, int blocked) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t >> 5, c8 = (t & 31) * 8;
  if (row >= B * geo.P0) return;
  const int b = row / geo.P0, i = row - b * geo.P0;
  const int vl = geo.vlen[b];
  uint4 o = make_uint4(0, 0, 0, 0);
  const bool valid = i < vl;
  if (valid) {
    const size_t frow = static_cast<size_t>(b) * Lv + i;
    float4 a, c;
    if (blocked) {
      a = __ldg(reinterpret_cast<const float4*>(F + blk_off(frow, c8)));
      c = __ldg(reinterpret_cast<const float4*>(F + blk_off(frow, c8 + 4)));
    } else {
      const float4* src = reinterpret_cast<const float4*>(F + frow * 256 + c8);
      a = __ldg(src);
      c = __ldg(src + 1);
    }
    o.x = pack_bf16(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
    o.y = pack_bf16(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
    o.z = pack_bf16(fmaxf(c.x, 0.f), fmaxf(c.y, 0.f));
    o.w = pack_bf16(fmaxf(c.z, 0.f), fmaxf(c.w, 0.f));
  }
  *reinterpret_cast<uint4*>(chain0 + static_cast<size_t>(row) * 256 + c8) = o;
  if (valid) {
    const size_t r1 = static_cast<size_t>(b) * geo.PH1 + geo.o1[0] + i;
    const size_t r2 = static_cast<size_t>(b) * geo.PH2 + geo.pad + i;
    *reinterpret_cast<uint4*>(H1 + r1 * 256 + c8) = o;
    *reinterpret_cast<uint4*>(H2 + r2 * 256 + c8) = o;
  }
}

int launch_level0(cudaStream_t st, const float* F, bf16* chain0, bf16* H1, bf16* H2, int B, int Lv,
                  const PyrGeo& geo, bool blocked) {
  const long long threads = static_cast<long long>(B) * geo.P0 * 32;
  if (threads <= 0) return FVTG_OK;
  const int grid = static_cast<int>((threads + 255) / 256);
  ProfScope prof(st, PC_OTHER);
  level0_kernel<<<grid, 256, 0, st>>>(F, chain0, H1, H2, B, Lv, geo, blocked ? 1 : 0);
  FVTG_LAUNCH_CHECK("level0_kernel");
  return FVTG_OK;
}

// thread = (output row, 4 columns)
__global__ void unblock_kernel(const float* __restrict__ src, float* __restrict__ dst, int B,
                               int rows_in, int rows_out) {
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long orow = t >> 6;
  const int c4 = static_cast<int>(t & 63) * 4;
  if (orow >= static_cast<long long>(B) * rows_out) return;
  const long long b = orow / rows_out, j = orow - b * rows_out;
  const float4 v = *reinterpret_cast<const float4*>(src + blk_off(static_cast<size_t>(b * rows_in + j), c4));
  *reinterpret_cast<float4*>(dst + orow * 256 + c4) = v;
}

int launch_unblock(cudaStream_t st, const float* src_blk, float* dst, int B, int rows_in, int rows_out) {
  const long long threads = static_cast<long long>(B) * rows_out * 64;
  if (threads <= 0) return FVTG_OK;
  ProfScope prof(st, PC_OTHER);
  unblock_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(src_blk, dst, B, rows_in, rows_out);
  FVTG_LAUNCH_CHECK("unblock_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
