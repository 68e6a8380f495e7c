// Device-resident input pipeline: what the reference's loader does
// per item on the host - l2_normalize_np_array per feature directory (start_end_dataset.py:524-530,
// utils/basic_utils.py:84-86), concatenation of the directories, temporal-endpoint features
// [i / L, (i + 1) / L] (start_end_dataset.py:174-180), zero padding + 0/1 masks
// (start_end_collate -> pad_sequences_1d, utils/tensor_utils.py:5-53) - as ONE pass over raw feature
// rows already on the device.  Raw rows may be fp32, fp16 or bf16 (the loader casts to fp32 first:
// "astype(np.float32)"), so feature stores kept in half precision cross PCIe at half the bytes.
//
// One warp per output row; a row's group is reduced with warp shuffles; all arithmetic fp32:
//   x / (sqrt(sum x^2) + 1e-5)        (division, like numpy; the sum order differs from numpy's
//                                      pairwise sum -> parity 1e-6 relative, not bit-exact)
//   tef_st = float(i) / float(L) ; tef_ed = tef_st + 1.0f / float(L)      (bit-exact, torch fp32)
#include "kernels.cuh"
#include "ptx.cuh"

#include <cuda_fp16.h>

namespace fvtg {

template <typename T>
__device__ __forceinline__ float ld_raw(const void* p, size_t i);
template <>
__device__ __forceinline__ float ld_raw<float>(const void* p, size_t i) {
  return __ldcs(static_cast<const float*>(p) + i);
}
template <>
__device__ __forceinline__ float ld_raw<__half>(const void* p, size_t i) {
  return __half2float(static_cast<const __half*>(p)[i]);
}
template <>
__device__ __forceinline__ float ld_raw<bf16>(const void* p, size_t i) {
  return __bfloat162float(static_cast<const bf16*>(p)[i]);
}

struct PrepArgs {
  int B, L;            // rows per item (padded length)
  int n_groups;
  int dim[FVTG_RAW_MAX_GROUPS];
  const void* src[FVTG_RAW_MAX_GROUPS];  // [B][L][dim[g]]
  const int* len;      // [B] true lengths
  int normalize, tef;
  int out_dim;         // sum(dim) + 2 * tef
  float* out;          // [B][L][out_dim]
  float* mask;         // [B][L] or null
};

template <typename T>
__global__ void __launch_bounds__(256)
prepare_rows_kernel(const PrepArgs a) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(a.B) * a.L) return;
  const int b = static_cast<int>(row / a.L), i = static_cast<int>(row - static_cast<long long>(b) * a.L);
  const int len = a.len[b];
  const bool valid = i < len;
  float* o = a.out + static_cast<size_t>(row) * a.out_dim;
  if (a.mask && lane == 0) a.mask[row] = valid ? 1.f : 0.f;
  int off = 0;
  for (int g = 0; g < a.n_groups; ++g) {
    const int d = a.dim[g];
    if (!valid) {
      for (int j = lane; j < d; j += 32) o[off + j] = 0.f;
    } else {
      const size_t base = static_cast<size_t>(row) * d;
      float div = 1.f;
      if (a.normalize) {
        float ss = 0.f;
        for (int j = lane; j < d; j += 32) {
          const float x = ld_raw<T>(a.src[g], base + j);
          ss += x * x;
        }
        ss = warp_sum(ss);
        div = __fadd_rn(sqrtf(ss), 1e-5f);
      }
      for (int j = lane; j < d; j += 32) {
        const float x = ld_raw<T>(a.src[g], base + j);
        o[off + j] = a.normalize ? __fdiv_rn(x, div) : x;
      }
    }
    off += d;
  }
  if (a.tef && lane == 0) {
    float st = 0.f, ed = 0.f;
    if (valid) {
      const float fl = static_cast<float>(len);
      st = __fdiv_rn(static_cast<float>(i), fl);
      ed = __fadd_rn(st, __fdiv_rn(1.f, fl));
    }
    o[off] = st;
    o[off + 1] = ed;
  }
}

static int launch_prepare(cudaStream_t st, const PrepArgs& a, int dtype) {
  const long long rows = static_cast<long long>(a.B) * a.L;
  if (rows <= 0) return FVTG_OK;
  const int grid = static_cast<int>((rows * 32 + 255) / 256);
  ProfScope prof(st, PC_OTHER);
  if (dtype == FVTG_RAW_F32) prepare_rows_kernel<float><<<grid, 256, 0, st>>>(a);
  else if (dtype == FVTG_RAW_F16) prepare_rows_kernel<__half><<<grid, 256, 0, st>>>(a);
  else if (dtype == FVTG_RAW_BF16) prepare_rows_kernel<bf16><<<grid, 256, 0, st>>>(a);
  else return fail(FVTG_EINVAL, "prepare_inputs: unknown raw dtype %d", dtype);
  FVTG_LAUNCH_CHECK("prepare_rows_kernel");
  return FVTG_OK;
}

}  // namespace fvtg

using namespace fvtg;

extern "C" int32_t fvtg_prepare_inputs(const FvtgRawBatch* raw, float* src_vid, float* src_vid_mask,
                                       float* src_txt, float* src_txt_mask, void* stream) {
  host_state().launches = 0;
  if (!raw || !src_vid || !src_txt || !raw->vid_len || !raw->txt_len || !raw->txt)
    return fail(FVTG_EINVAL, "prepare_inputs: null argument");
  if (raw->B < 1 || raw->Lv < 1 || raw->Lt < 1 || raw->n_groups < 1 || raw->n_groups > FVTG_RAW_MAX_GROUPS ||
      raw->t_dim < 1)
    return fail(FVTG_EINVAL, "prepare_inputs: bad shapes");
  FVTG_TRY(check_arch());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PrepArgs v;
  memset(&v, 0, sizeof(v));
  v.B = raw->B; v.L = raw->Lv; v.n_groups = raw->n_groups;
  v.out_dim = raw->use_tef ? 2 : 0;
  for (int g = 0; g < raw->n_groups; ++g) {
    if (raw->group_dim[g] < 1 || !raw->vid[g]) return fail(FVTG_EINVAL, "prepare_inputs: group %d", g);
    v.dim[g] = raw->group_dim[g];
    v.src[g] = raw->vid[g];
    v.out_dim += raw->group_dim[g];
  }
  v.len = raw->vid_len; v.normalize = raw->normalize_v; v.tef = raw->use_tef;
  v.out = src_vid; v.mask = src_vid_mask;
  FVTG_TRY(launch_prepare(st, v, raw->dtype));
  PrepArgs t;
  memset(&t, 0, sizeof(t));
  t.B = raw->B; t.L = raw->Lt; t.n_groups = 1; t.dim[0] = raw->t_dim; t.src[0] = raw->txt;
  t.out_dim = raw->t_dim; t.len = raw->txt_len; t.normalize = raw->normalize_t; t.tef = 0;
  t.out = src_txt; t.mask = src_txt_mask;
  return launch_prepare(st, t, raw->dtype);
}
