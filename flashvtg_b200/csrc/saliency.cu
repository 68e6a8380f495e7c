// Saliency head (transformer.py:106-113) and the t2vattnvalues finalisation (model.py:215-216).
//   g = mean_i F_i ; u = W2 g + b2 ; sal_i = <W1 F_i + b1, u> / 16
// evaluated as sal_i = (F_i . (W1^T u) + b1 . u) / 16: two 256x256 mat-vecs per video instead of a
// 256x256 GEMM per clip.  SAL_VPB videos share one pass over the two weight matrices.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int SAL_VPB = 4;

__global__ void __launch_bounds__(256)
saliency_kernel(const float* __restrict__ F, const int* __restrict__ vlen,
                const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2t, const float* __restrict__ b2,
                const float* __restrict__ tsum, int t2v_layers, float* __restrict__ sal_out,
                float* __restrict__ t2v_out, int B, int Lv) {
  __shared__ float s_g[SAL_VPB][256];
  __shared__ float s_u[SAL_VPB][256];
  __shared__ float s_w[SAL_VPB][256];
  __shared__ float s_c[SAL_VPB];
  __shared__ float s_red[8];
  const int c = threadIdx.x;
  const int b0 = blockIdx.x * SAL_VPB;
  const int warp = c >> 5, lane = c & 31;
#pragma unroll
  for (int v = 0; v < SAL_VPB; ++v) {
    const int b = b0 + v;
    float acc = 0.f;
    if (b < B) {
      const int len = vlen[b];
      const size_t r0 = static_cast<size_t>(b) * Lv;
      for (int i = 0; i < len; ++i) acc += F[blk_off(r0 + i, c)];
      acc /= static_cast<float>(len > 0 ? len : 1);
    }
    s_g[v][c] = acc;
  }
  __syncthreads();
  {  // u[n] = b2[n] + sum_k W2t[k][n] g[k]
    float u[SAL_VPB];
#pragma unroll
    for (int v = 0; v < SAL_VPB; ++v) u[v] = 0.f;
    for (int k = 0; k < 256; ++k) {
      const float w = __ldg(w2t + k * 256 + c);
#pragma unroll
      for (int v = 0; v < SAL_VPB; ++v) u[v] += w * s_g[v][k];
    }
    const float bb = b2[c];
#pragma unroll
    for (int v = 0; v < SAL_VPB; ++v) s_u[v][c] = u[v] + bb;
  }
  __syncthreads();
  {  // w[k] = sum_n W1[n][k] u[n] ; c0 = sum_n b1[n] u[n]
    float w[SAL_VPB];
#pragma unroll
    for (int v = 0; v < SAL_VPB; ++v) w[v] = 0.f;
    for (int n = 0; n < 256; ++n) {
      const float x = __ldg(w1 + n * 256 + c);
#pragma unroll
      for (int v = 0; v < SAL_VPB; ++v) w[v] += x * s_u[v][n];
    }
#pragma unroll
    for (int v = 0; v < SAL_VPB; ++v) s_w[v][c] = w[v];
    const float bb = b1[c];
    for (int v = 0; v < SAL_VPB; ++v) {
      float p = warp_sum(bb * s_u[v][c]);
      if (lane == 0) s_red[warp] = p;
      __syncthreads();
      if (c == 0) {
        float tot = 0.f;
        for (int i = 0; i < 8; ++i) tot += s_red[i];
        s_c[v] = tot;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int v = 0; v < SAL_VPB; ++v) {
    const int b = b0 + v;
    if (b >= B) break;
    const int len = vlen[b];
    for (int i = warp; i < Lv; i += 8) {
      const size_t row = static_cast<size_t>(b) * Lv + i;
      float acc = 0.f;
      if (i < len) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += F[blk_off(row, lane + 32 * q)] * s_w[v][lane + 32 * q];
        acc = warp_sum(acc);
      }
      if (lane == 0) {
        sal_out[row] = i < len ? (acc + s_c[v]) * 0.0625f : 0.f;
        if (t2v_out) {
          float tv = 0.f;
          if (i < len && tsum) {
            const size_t hs = static_cast<size_t>(B) * Lv;
            for (int hh = 0; hh < 8; ++hh) tv += tsum[hh * hs + row];
            tv = tv / static_cast<float>(8 * t2v_layers);
            tv = fminf(fmaxf(tv, 0.f), 1.f);
          }
          t2v_out[row] = tv;
        }
      }
    }
  }
}

int launch_saliency(cudaStream_t st, const float* F, const int* vlen, const float* w1,
                    const float* b1, const float* w2t, const float* b2, const float* tsum,
                    int t2v_layers, float* sal_out, float* t2v_out, int B, int Lv) {
  if (B <= 0) return FVTG_OK;
  ProfScope prof(st, PC_OTHER);
  saliency_kernel<<<(B + SAL_VPB - 1) / SAL_VPB, 256, 0, st>>>(F, vlen, w1, b1, w2t, b2, tsum,
                                                              t2v_layers, sal_out, t2v_out, B, Lv);
  FVTG_LAUNCH_CHECK("saliency_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
