// Saliency head (transformer.py:106-113) and the t2vattnvalues finalisation (model.py:215-216).
//   g = mean_i F_i ; u = W2 g + b2 ; sal_i = <W1 F_i + b1, u> / 16
// evaluated as sal_i = (F_i . (W1^T u) + b1 . u) / 16: two 256x256 mat-vecs per video instead of a
// 256x256 GEMM per clip.  Three small kernels, all reading the tile-blocked fp32 stream with lanes
// running over consecutive rows (16 contiguous bytes per lane):
//   colmean : G[b][c]  = mean over the video's rows
//   matvec  : WV[b][:] = W1^T (W2 G[b] + b2),  C[b] = b1 . (W2 G[b] + b2)     (8 videos per CTA)
//   rowdot  : sal[row] = (F_row . WV[b] + C[b]) / 16 ; t2v[row] = clamp(mean over layers x heads)
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

// one CTA (256 threads = 8 warps) per video; warp w owns column groups w, w+8, ...
__global__ void __launch_bounds__(256)
sal_colmean_kernel(const float* __restrict__ F, const int* __restrict__ vlen, float* __restrict__ G,
                   int Lv) {
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = vlen[b];
  const size_t r0 = static_cast<size_t>(b) * Lv;
  const float inv = 1.f / static_cast<float>(len > 0 ? len : 1);
  for (int c4 = warp; c4 < 64; c4 += 8) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < len; i += 32) {
      const float4 v = *reinterpret_cast<const float4*>(F + blk_off(r0 + i, c4 * 4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
    if (lane == 0)
      *reinterpret_cast<float4*>(G + static_cast<size_t>(b) * 256 + c4 * 4) =
          make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

constexpr int SAL_VPB = 2;  // few videos per CTA: the weight-load latency chain is hidden by 4+ co-resident CTAs per SM

__global__ void __launch_bounds__(256)
sal_matvec_kernel(const float* __restrict__ G, const float* __restrict__ w1,
                  const float* __restrict__ b1, const float* __restrict__ w2t,
                  const float* __restrict__ b2, float* __restrict__ WV, float* __restrict__ Cc, int B) {
  __shared__ __align__(16) float s_g[SAL_VPB][256];
  __shared__ __align__(16) float s_u[SAL_VPB][256];
  __shared__ float s_red[SAL_VPB][8];
  const int c = threadIdx.x, warp = c >> 5, lane = c & 31;
  const int b0 = blockIdx.x * SAL_VPB;
#pragma unroll
  for (int v = 0; v < SAL_VPB; ++v) s_g[v][c] = (b0 + v < B) ? G[static_cast<size_t>(b0 + v) * 256 + c] : 0.f;
  __syncthreads();
  float u[SAL_VPB];
#pragma unroll
  for (int v = 0; v < SAL_VPB; ++v) u[v] = 0.f;
  // u[n] = b2[n] + sum_k W2t[k][n] g[k]: 8 independent weight loads in flight per thread, the
  // videos' g values come as 16-byte broadcast loads (4 k per LDS)
  for (int k0 = 0; k0 < 256; k0 += 8) {
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = __ldg(w2t + (k0 + j) * 256 + c);
#pragma unroll
    for (int v = 0; v < SAL_VPB; ++v) {
      const float4 ga = *reinterpret_cast<const float4*>(&s_g[v][k0]);
      const float4 gb = *reinterpret_cast<const float4*>(&s_g[v][k0 + 4]);
      // same summation order as a plain k loop (bit-identical to the scalar version)
      u[v] += w[0] * ga.x; u[v] += w[1] * ga.y; u[v] += w[2] * ga.z; u[v] += w[3] * ga.w;
      u[v] += w[4] * gb.x; u[v] += w[5] * gb.y; u[v] += w[6] * gb.z; u[v] += w[7] * gb.w;
    }
  }
  const float bb2 = b2[c], bb1 = b1[c];
#pragma unroll
  for (int v = 0; v < SAL_VPB; ++v) {
    u[v] += bb2;
    s_u[v][c] = u[v];
    const float p = warp_sum(bb1 * u[v]);
    if (lane == 0) s_red[v][warp] = p;
  }
  __syncthreads();
  float wv[SAL_VPB];
#pragma unroll
  for (int v = 0; v < SAL_VPB; ++v) wv[v] = 0.f;
  for (int n0 = 0; n0 < 256; n0 += 8) {  // wv[k] = sum_n W1[n][k] u[n]
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __ldg(w1 + (n0 + j) * 256 + c);
#pragma unroll
    for (int v = 0; v < SAL_VPB; ++v) {
      const float4 ua = *reinterpret_cast<const float4*>(&s_u[v][n0]);
      const float4 ub = *reinterpret_cast<const float4*>(&s_u[v][n0 + 4]);
      wv[v] += x[0] * ua.x; wv[v] += x[1] * ua.y; wv[v] += x[2] * ua.z; wv[v] += x[3] * ua.w;
      wv[v] += x[4] * ub.x; wv[v] += x[5] * ub.y; wv[v] += x[6] * ub.z; wv[v] += x[7] * ub.w;
    }
  }
#pragma unroll
  for (int v = 0; v < SAL_VPB; ++v) {
    if (b0 + v < B) {
      WV[static_cast<size_t>(b0 + v) * 256 + c] = wv[v];
      if (c == 0) {
        float tot = 0.f;
        for (int i = 0; i < 8; ++i) tot += s_red[v][i];
        Cc[b0 + v] = tot;
      }
    }
  }
}

// thread per row
__global__ void __launch_bounds__(128)
sal_rowdot_kernel(const float* __restrict__ F, const int* __restrict__ vlen,
                  const float* __restrict__ WV, const float* __restrict__ Cc,
                  const float* __restrict__ tsum, int t2v_layers, float* __restrict__ sal_out,
                  float* __restrict__ t2v_out, int B, int Lv) {
  const long long row = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(B) * Lv;
  if (row >= rows) return;
  const int b = static_cast<int>(row / Lv), i = static_cast<int>(row - static_cast<long long>(b) * Lv);
  const bool valid = i < vlen[b];
  float acc = 0.f;
  if (valid) {
    const float* f = F + blk_off(static_cast<size_t>(row), 0);
    const float4* wv = reinterpret_cast<const float4*>(WV + static_cast<size_t>(b) * 256);
#pragma unroll 8
    for (int c4 = 0; c4 < 64; ++c4) {
      const float4 x = *reinterpret_cast<const float4*>(f + c4 * 512);
      const float4 w = __ldg(wv + c4);
      acc += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
    }
    acc = (acc + Cc[b]) * 0.0625f;
  }
  sal_out[row] = acc;
  if (t2v_out) {
    float tv = 0.f;
    if (valid && tsum && t2v_layers > 0) {
      for (int hh = 0; hh < 8 * t2v_layers; ++hh) tv += tsum[static_cast<size_t>(hh) * rows + row];
      tv = fminf(fmaxf(tv / static_cast<float>(8 * t2v_layers), 0.f), 1.f);
    }
    t2v_out[row] = tv;
  }
}

int launch_saliency(cudaStream_t st, const float* F, const int* vlen, const float* w1,
                    const float* b1, const float* w2t, const float* b2, const float* tsum,
                    int t2v_layers, float* scratch, float* sal_out, float* t2v_out, int B, int Lv) {
  if (B <= 0) return FVTG_OK;
  float* G = scratch;                                   // [B][256]
  float* WV = scratch + static_cast<size_t>(B) * 256;   // [B][256]
  float* Cc = scratch + static_cast<size_t>(B) * 512;   // [B]
  ProfScope prof(st, PC_OTHER);
  sal_colmean_kernel<<<B, 256, 0, st>>>(F, vlen, G, Lv);
  FVTG_LAUNCH_CHECK("sal_colmean_kernel");
  sal_matvec_kernel<<<(B + SAL_VPB - 1) / SAL_VPB, 256, 0, st>>>(G, w1, b1, w2t, b2, WV, Cc, B);
  FVTG_LAUNCH_CHECK("sal_matvec_kernel");
  const long long rows = static_cast<long long>(B) * Lv;
  sal_rowdot_kernel<<<static_cast<int>((rows + 127) / 128), 128, 0, st>>>(F, vlen, WV, Cc, tsum,
                                                                          t2v_layers, sal_out,
                                                                          t2v_out, B, Lv);
  FVTG_LAUNCH_CHECK("sal_rowdot_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
