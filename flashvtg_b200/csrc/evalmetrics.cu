// QVHighlights evaluation per query on the device (include/flashvtg_b200.h: fvtg_eval_submission).
//
// The reference computes these metrics with Python loops over queries spread over an 8-process pool
// (standalone_eval/eval.py:24-69 compute_mr_ap, :72-102 compute_mr_r1, :173-236 highlight Hit1 / AP;
// standalone_eval/utils.py:65-159 detection AP, :162-209 get_ap over sklearn's precision_recall_curve).
// Here one thread owns one (length range, query) pair of the moment-retrieval metrics and one warp
// owns one query of the highlight metrics.  Everything is fp64 and follows numpy's operation order -
// including np.sum's 8-accumulator pairwise reduction - so the per-query numbers are bit-identical to
// the oracle restatement; the (tiny) means over queries are left to the host.
//
// This is index / compare work on a few KB per query: no tensor cores, no staging; the grid is sized
// by the query count and the kernels are bound by their serial per-query loops.
#include "common.cuh"

namespace fvtg {

constexpr int EV_MAX_GT = 32;
constexpr int EV_MAX_PRED = 32;
constexpr int EV_THDS = 10;

struct EvalThds {
  double v[EV_THDS];
};

// numpy's pairwise summation of a contiguous fp64 array (numpy/_core/src/umath/loops_utils.h.src):
// one block of <= 128 values ...
__device__ double np_sum_block(const double* a, int n) {
  if (n < 8) {
    double res = 0.;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  }
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}
// ... and its recursive halving above 128 values (n2 = n / 2 rounded down to a multiple of 8), unrolled
// into an explicit stack: device recursion would need a run-time stack-size limit.
__device__ double np_pairwise_sum(const double* a, int n) {
  if (n <= 128) return np_sum_block(a, n);
  struct Frame {
    const double* a;
    int n, state;
    double left;
  } st[24];
  int sp = 0;
  st[0] = {a, n, 0, 0.};
  double ret = 0.;
  while (sp >= 0) {
    Frame& f = st[sp];
    if (f.state == 0) {
      if (f.n <= 128) {
        ret = np_sum_block(f.a, f.n);
        --sp;
      } else {
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        f.state = 1;
        st[sp + 1] = {f.a, n2, 0, 0.};
        ++sp;
      }
    } else if (f.state == 1) {
      int n2 = f.n / 2;
      n2 -= n2 % 8;
      f.left = ret;
      f.state = 2;
      st[sp + 1] = {f.a + n2, f.n - n2, 0, 0.};
      ++sp;
    } else {
      ret = f.left + ret;
      --sp;
    }
  }
  return ret;
}

// ------------------------------------------------------------- moment retrieval --
__global__ void __launch_bounds__(128)
eval_mr_kernel(const FvtgEvalBatch b, const int max_pred_windows, const EvalThds thds, double* __restrict__ mr_ap,
               double* __restrict__ mr_iou, uint8_t* __restrict__ mr_valid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int Q = b.n_queries;
  if (t >= 4 * Q) return;
  const int r = t / Q, q = t - r * Q;
  double* ap = mr_ap + (static_cast<size_t>(r) * Q + q) * EV_THDS;
  for (int k = 0; k < EV_THDS; ++k) ap[k] = 0.;
  mr_iou[t] = 0.;
  mr_valid[t] = 0;

  // ground-truth windows of this length range (eval.py:109-137): lo < end - start <= hi, "full" keeps all
  const double lo = r == 0 ? 0. : (r == 1 ? 10. : (r == 2 ? 30. : 0.));
  const double hi = r == 0 ? 10. : (r == 1 ? 30. : 150.);
  double gs[EV_MAX_GT], ge[EV_MAX_GT];
  int G = 0;
  {
    const double* gw = b.gt_win + static_cast<size_t>(q) * b.max_gt * 2;
    const int n = min(b.gt_cnt[q], EV_MAX_GT);
    for (int j = 0; j < n; ++j) {
      const double s = gw[2 * j], e = gw[2 * j + 1], len = e - s;
      if (r == 3 || (lo < len && len <= hi)) {
        gs[G] = s;
        ge[G] = e;
        ++G;
      }
    }
  }
  if (G == 0) return;
  mr_valid[t] = 1;
  const double* pw = b.pred_win + static_cast<size_t>(q) * b.max_pred * 3;
  const int n_pred = b.pred_cnt[q];
  if (n_pred <= 0) return;

  // R1 / mIoU (eval.py:72-102): the FIRST listed window against the GT window with the highest IoU
  {
    const double ps = pw[0], pe = pw[1];
    int best = 0;
    double best_iou = 0.;
    for (int j = 0; j < G; ++j) {
      const double inter = fmax(fmin(pe, ge[j]) - fmax(ps, gs[j]), 0.);
      const double iou = inter / ((pe - ps) + (ge[j] - gs[j]) - inter);
      if (j == 0 || iou > best_iou) {
        best_iou = iou;
        best = j;
      }
    }
    const double inter = fmax(0., fmin(pe, ge[best]) - fmax(ps, gs[best]));
    const double uni = fmax(pe, ge[best]) - fmin(ps, gs[best]);
    mr_iou[t] = uni != 0. ? inter / uni : 0.;
  }

  // detection AP (utils.py:83-159) over the first max_pred_windows listed windows
  const int P = min(min(n_pred, max_pred_windows), EV_MAX_PRED);
  int order[EV_MAX_PRED];   // stable sort by decreasing score (list.sort(key=-score))
  for (int i = 0; i < P; ++i) {
    const double sc = pw[3 * i + 2];
    int k = i;
    while (k > 0 && pw[3 * order[k - 1] + 2] < sc) {
      order[k] = order[k - 1];
      --k;
    }
    order[k] = i;
  }
  uint32_t lock[EV_THDS], tp[EV_THDS];
  for (int k = 0; k < EV_THDS; ++k) lock[k] = tp[k] = 0u;
  for (int idx = 0; idx < P; ++idx) {
    const double ps = pw[3 * order[idx]], pe = pw[3 * order[idx] + 1];
    double iou[EV_MAX_GT];
    int by[EV_MAX_GT];   // argsort(iou)[::-1] of a stable ascending sort: ties -> higher index first
    for (int j = 0; j < G; ++j) {
      const double inter = fmax(fmin(pe, ge[j]) - fmax(ps, gs[j]), 0.);
      iou[j] = inter / ((pe - ps) + (ge[j] - gs[j]) - inter);
      int k = j;
      while (k > 0 && iou[by[k - 1]] <= iou[j]) {
        by[k] = by[k - 1];
        --k;
      }
      by[k] = j;
    }
    for (int k = 0; k < EV_THDS; ++k) {
      for (int jj = 0; jj < G; ++jj) {
        const int j = by[jj];
        if (iou[j] < thds.v[k]) break;                 // false positive
        if (lock[k] >> j & 1u) continue;
        tp[k] |= 1u << idx;
        lock[k] |= 1u << j;
        break;
      }
    }
  }
  for (int k = 0; k < EV_THDS; ++k) {
    // interpolated_precision_recall (utils.py:65-80) on cumulative precision / recall
    double mp[EV_MAX_PRED + 2], mr[EV_MAX_PRED + 2], term[EV_MAX_PRED + 2];
    mp[0] = 0.;
    mr[0] = 0.;
    double tpc = 0.;
    for (int i = 0; i < P; ++i) {
      tpc += (tp[k] >> i & 1u) ? 1. : 0.;
      mp[i + 1] = tpc / static_cast<double>(i + 1);    // tp + fp == i + 1 exactly
      mr[i + 1] = tpc / static_cast<double>(G);
    }
    mp[P + 1] = 0.;
    mr[P + 1] = 1.;
    for (int i = P; i >= 0; --i) mp[i] = fmax(mp[i], mp[i + 1]);
    int n = 0;
    for (int i = 1; i <= P + 1; ++i)
      if (mr[i] != mr[i - 1]) term[n++] = (mr[i] - mr[i - 1]) * mp[i];
    ap[k] = np_sum_block(term, n);   // n <= 34
  }
}

// ------------------------------------------------------------------- highlights --
// one warp per query; dynamic shared memory per warp: y[C] f64, ys[C] f64 (sorted), pos[C] u8 x 3
// (annotator scores in sorted order), sel[9][C + 1] f64
__global__ void __launch_bounds__(128)
eval_hl_kernel(const FvtgEvalBatch b, double* __restrict__ hl_ap, uint8_t* __restrict__ hl_hit) {
  extern __shared__ __align__(16) uint8_t ev_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  const int Q = b.n_queries, C = b.max_clips;
  if (q >= Q) return;
  const size_t per_warp = (static_cast<size_t>(C) * 2 + 9 * (static_cast<size_t>(C) + 1)) * 8 +
                          ((static_cast<size_t>(C) * 3 + 15) & ~static_cast<size_t>(15));
  uint8_t* base = ev_smem + warp * per_warp;
  double* y = reinterpret_cast<double*>(base);
  double* ys = y + C;
  double* sel = ys + C;                                   // [9][C + 1]
  uint8_t* sc = reinterpret_cast<uint8_t*>(sel + 9 * (C + 1));   // [C][3] scores in sorted order

  const int n = min(b.gt_clips[q], C);
  const int lp = min(b.pred_sal_len[q], b.max_sal);
  const double* ps = b.pred_sal + static_cast<size_t>(q) * b.max_sal;
  const uint8_t* gsal = b.gt_sal + static_cast<size_t>(q) * C * 3;

  // Hit1 (eval.py:173-186): first arg-max of ALL predicted scores, looked up in the GT if inside
  {
    double bv = 0.;
    int bi = 0x7fffffff;
    for (int i = lane; i < lp; i += 32) {
      const double v = ps[i];
      if (bi == 0x7fffffff || v > bv) {
        bv = v;
        bi = i;
      }
    }
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane < 3) {
      const int smin = 2 + lane;
      uint8_t h = 0;
      if (bi < n) h = (gsal[bi * 3] >= smin) | (gsal[bi * 3 + 1] >= smin) | (gsal[bi * 3 + 2] >= smin);
      hl_hit[static_cast<size_t>(lane) * Q + q] = h;
    }
  }

  // scores cut / zero-padded to the GT length (eval.py:221-229), then sorted by decreasing score the
  // way argsort(kind="mergesort")[::-1] orders them: ties -> higher index first
  for (int i = lane; i < n; i += 32) y[i] = i < lp ? ps[i] : 0.;
  __syncwarp();
  for (int e = lane; e < n; e += 32) {
    const double v = y[e];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double w = y[j];
      rank += (w > v) || (w == v && j > e);
    }
    ys[rank] = v;
    sc[rank * 3] = gsal[e * 3];
    sc[rank * 3 + 1] = gsal[e * 3 + 1];
    sc[rank * 3 + 2] = gsal[e * 3 + 2];
  }
  __syncwarp();

  if (lane < 9) {
    // get_ap (utils.py:162-209) for (minimum score m, annotator w)
    const int m = lane / 3, w = lane - 3 * m, smin = 2 + m;
    int T = 0;
    for (int i = 0; i < n; ++i) T += sc[i * 3 + w] >= smin;
    double ap;
    if (n == 0 || T == 0) {
      ap = 0.;
    } else if (T == n) {
      ap = 1.;
    } else {
      // precision_recall_curve walks the distinct-score groups from the lowest threshold (all clips)
      // to the highest; precision is made non-decreasing along that walk and averaged over the
      // points after which the (float32) recall changes
      double* s = sel + lane * (C + 1);
      int cnt = 0, above = 0;          // positives at sorted positions > i
      bool have = false;
      double pmax = 0.;
      float prev_rec = 0.f;
      for (int i = n - 1; i >= 0; --i) {
        if (i == n - 1 || ys[i] != ys[i + 1]) {
          const double tps = static_cast<double>(T - above);
          const double prec = tps / static_cast<double>(i + 1);
          const float rec = static_cast<float>(tps / static_cast<double>(T));
          if (have && rec != prev_rec) s[cnt++] = pmax;
          pmax = have ? fmax(pmax, prec) : prec;
          prev_rec = rec;
          have = true;
        }
        above += sc[i * 3 + w] >= smin;
      }
      if (have && 0.f != prev_rec) s[cnt++] = pmax;   // the appended (precision 1, recall 0) end point
      ap = np_pairwise_sum(s, cnt) / static_cast<double>(cnt);
    }
    hl_ap[(static_cast<size_t>(m) * Q + q) * 3 + w] = ap;
  }
}

}  // namespace fvtg

extern "C" int32_t fvtg_eval_submission(const FvtgEvalBatch* batch, int32_t max_pred_windows, double* mr_ap,
                                        double* mr_iou, uint8_t* mr_valid, double* hl_ap, uint8_t* hl_hit,
                                        void* stream) {
  using namespace fvtg;
  host_state().launches = 0;
  if (!batch) return fail(FVTG_EINVAL, "eval: null batch");
  const FvtgEvalBatch& b = *batch;
  if (b.n_queries <= 0) return FVTG_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mr_ap) {
    if (!mr_iou || !mr_valid || !b.pred_win || !b.pred_cnt || !b.gt_win || !b.gt_cnt)
      return fail(FVTG_EINVAL, "eval: moment-retrieval metrics need windows, counts and all three outputs");
    if (b.max_gt < 1 || b.max_gt > EV_MAX_GT || max_pred_windows < 1 || max_pred_windows > EV_MAX_PRED ||
        b.max_pred < 1)
      return fail(FVTG_EINVAL, "eval: at most %d GT windows per query and %d predicted windows are scored",
                  EV_MAX_GT, EV_MAX_PRED);
    EvalThds th;
    // [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]: the decimal literals themselves
    const double lit[EV_THDS] = {0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95};
    for (int k = 0; k < EV_THDS; ++k) th.v[k] = lit[k];
    const int threads = 128, total = 4 * b.n_queries;
    ProfScope prof(st, PC_OTHER);
    eval_mr_kernel<<<(total + threads - 1) / threads, threads, 0, st>>>(b, max_pred_windows, th, mr_ap, mr_iou,
                                                                        mr_valid);
    FVTG_LAUNCH_CHECK("eval_mr_kernel");
  }
  if (hl_ap) {
    if (!hl_hit || !b.pred_sal || !b.pred_sal_len || !b.gt_sal || !b.gt_clips)
      return fail(FVTG_EINVAL, "eval: highlight metrics need saliency scores, GT scores and both outputs");
    const size_t C = static_cast<size_t>(b.max_clips);
    if (b.max_clips < 1 || b.max_sal < 1) return fail(FVTG_EINVAL, "eval: empty saliency rows");
    const size_t per_warp = (C * 2 + 9 * (C + 1)) * 8 + ((C * 3 + 15) & ~static_cast<size_t>(15));
    int warps = 4;
    while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
    if (per_warp > 200 * 1024) return fail(FVTG_EINVAL, "eval: %d clips per video exceed shared memory", b.max_clips);
    const size_t smem = warps * per_warp;
    FVTG_CUDA_OK(cudaFuncSetAttribute(eval_hl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    ProfScope prof(st, PC_OTHER);
    eval_hl_kernel<<<(b.n_queries + warps - 1) / warps, warps * 32, smem, st>>>(b, hl_ap, hl_hit);
    FVTG_LAUNCH_CHECK("eval_hl_kernel");
  }
  return FVTG_OK;
}
