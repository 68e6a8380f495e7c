/* TEST INFRASTRUCTURE ONLY - plain-C restatement of the reference's temporal NMS arithmetic.
 *
 *   oracle_nms_f32   : post_processing_mr_nms, FlashVTG/inference.py:36-57 (fp32 tensors;
 *                      IoU = nncore.ops.temporal_iou == FlashVTG/span_utils.py:61-70)
 *   oracle_nms_hull  : temporal_nms, utils/temporal_nms.py:25-74 (python floats = fp64)
 *
 * Built by oracle/Makefile with -O2 -ffp-contract=off so that every fp32 operation rounds
 * separately, as torch's CPU kernels do.  Checked against the unmodified reference functions in
 * tests/test_oracle_vs_reference.py and against tests/golden/nms_*.json.  Never linked into the
 * product.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* rows: n x 3 fp32 (st, ed, score), modified in place into the final order.
 * order[i]: source row of final row i.  sel[i]: source row selected at step i.
 * mode 0 = normal (score <- 0 if iou >= thd), 1 = linear (score *= 1 - iou). */
void oracle_nms_f32(float* rows, int n, float thd, int mode, int32_t* order, int32_t* sel) {
  int32_t src[1024];
  float tmp[3];
  if (n > 1024) n = 1024;
  for (int i = 0; i < n; ++i) src[i] = i;
  for (int i = 0; i < n; ++i) {
    int m = i;
    float best = rows[i * 3 + 2];
    if (!isnan(best)) {
      for (int r = i + 1; r < n; ++r) {
        float s = rows[r * 3 + 2];
        if (isnan(s)) { m = r; break; }         /* torch.argmax: first NaN wins */
        if (s > best) { best = s; m = r; }      /* strict >: first index among ties */
      }
    }
    if (m != i) {
      memcpy(tmp, rows + i * 3, sizeof tmp);
      memcpy(rows + i * 3, rows + m * 3, sizeof tmp);
      memcpy(rows + m * 3, tmp, sizeof tmp);
      int32_t t = src[i]; src[i] = src[m]; src[m] = t;
    }
    const float s0 = rows[i * 3], e0 = rows[i * 3 + 1];
    const float a0 = e0 - s0;
    for (int r = i + 1; r < n; ++r) {
      const float s1 = rows[r * 3], e1 = rows[r * 3 + 1];
      const float a1 = e1 - s1;
      float inter = fminf(e0, e1) - fmaxf(s0, s1);
      if (inter < 0.0f) inter = 0.0f;           /* clamp(min=0); NaN stays NaN */
      const float uni = (a0 + a1) - inter;
      const float iou = inter / uni;            /* 0/0 = NaN */
      if (mode == 0) {
        if (iou >= thd) rows[r * 3 + 2] = 0.0f; /* NaN compares false: kept */
      } else {
        rows[r * 3 + 2] = rows[r * 3 + 2] * (1.0f - iou);
      }
    }
  }
  if (sel) for (int i = 0; i < n; ++i) sel[i] = src[i];
  /* stable descending insertion sort, NaN first (torch.sort(descending=True)) */
  for (int i = 1; i < n; ++i) {
    float key[3];
    memcpy(key, rows + i * 3, sizeof key);
    int32_t ks = src[i];
    int j = i - 1;
    while (j >= 0) {
      const float sj = rows[j * 3 + 2];
      int before;                                /* does key sort strictly before rows[j]? */
      if (isnan(key[2])) before = !isnan(sj);
      else if (isnan(sj)) before = 0;
      else before = key[2] > sj;
      if (!before) break;
      memcpy(rows + (j + 1) * 3, rows + j * 3, sizeof key);
      src[j + 1] = src[j];
      --j;
    }
    memcpy(rows + (j + 1) * 3, key, sizeof key);
    src[j + 1] = ks;
  }
  for (int i = 0; i < n; ++i) order[i] = src[i];
}

/* rows: n x 3 fp64.  out: kept rows (<= n), out_src: their source indices.  Returns count. */
int oracle_nms_hull(const double* rows, int n, double thd, int max_after_nms, double* out,
                    int32_t* out_src) {
  int32_t idx[1024];
  uint8_t dead[1024];
  if (n > 1024) n = 1024;
  if (n == 1) {
    memcpy(out, rows, 3 * sizeof(double));
    out_src[0] = 0;
    return 1;
  }
  for (int i = 0; i < n; ++i) { idx[i] = i; dead[i] = 0; }
  /* stable descending sort by score (python sorted(reverse=True) keeps ties in input order) */
  for (int i = 1; i < n; ++i) {
    int32_t k = idx[i];
    int j = i - 1;
    while (j >= 0 && rows[idx[j] * 3 + 2] < rows[k * 3 + 2]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = k;
  }
  int cnt = 0, alive = n;
  for (int a = 0; a < n; ++a) {
    if (dead[a]) continue;
    /* the python loop runs while len(t) > 1 and kept < max; the trailing element is appended
       afterwards if kept < max: together "take heads in order until max_after_nms" */
    if (cnt >= max_after_nms) break;
    const double s0 = rows[idx[a] * 3], e0 = rows[idx[a] * 3 + 1];
    if (alive > 1) {
      for (int b = a + 1; b < n; ++b) {
        if (dead[b]) continue;
        const double s1 = rows[idx[b] * 3], e1 = rows[idx[b] * 3 + 1];
        double inter = fmin(e0, e1) - fmax(s0, s1);
        if (inter < 0) inter = 0;
        const double uni = fmax(e0, e1) - fmin(s0, s1);
        const double iou = uni == 0 ? 0.0 : inter / uni;
        if (iou > thd) { dead[b] = 1; --alive; }
      }
    }
    memcpy(out + cnt * 3, rows + idx[a] * 3, 3 * sizeof(double));
    out_src[cnt++] = idx[a];
    dead[a] = 1;
    --alive;
  }
  return cnt;
}
