// Kernel group C: Adaptive Score Refinement mix, sigmoid, span decode, top-k, the host
// post-processing chain (clamp / 4-decimal rounding / PostProcessorDETR) and temporal NMS,
// one CTA per video.  Reference: FlashVTG/model.py:201,247-266; FlashVTG/inference.py:286-290;
// FlashVTG/postprocessing.py:25-50; FlashVTG/inference.py:36-57; utils/temporal_nms.py:25-74.
//
// Everything that decides an index (sort keys, IoU, threshold compares) uses explicitly rounded
// fp32 / fp64 operations (__fadd_rn ...; no FMA contraction) in the reference's operation order,
// so NMS selection order and keep masks are bit-exact for identical inputs.
#include "kernels.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int DEC_THREADS = 128;  // 54 regs x 128 threads: 9 CTAs per SM, so 1024 videos fit one wave
constexpr int NMS_MAX = FVTG_MAX_TOPK;  // 64 rows

struct NmsSmem {
  float st[NMS_MAX], ed[NMS_MAX], sc[NMS_MAX];
  int src[NMS_MAX];
  float o_st[NMS_MAX], o_ed[NMS_MAX], o_sc[NMS_MAX];
  int o_src[NMS_MAX];
  int o_cnt;
};

__device__ __forceinline__ float round4_f32(float x) {
  // float(f"{x:.4f}") then back to fp32 (torch.tensor(list)): x * 1e4 is exact in fp64.
  return static_cast<float>(__ddiv_rn(rint(__dmul_rn(static_cast<double>(x), 1e4)), 1e4));
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  // torch.clamp: NaN propagates; min then max like at::clamp(min, max)
  if (x != x) return x;
  return fminf(fmaxf(x, lo), hi);
}

// a sorts strictly before b in a descending sort with NaN first and position as tie-break
__device__ __forceinline__ bool sorts_before(float sa, int pa, float sb, int pb) {
  const bool na = sa != sa, nb = sb != sb;
  if (na != nb) return na;
  if (!na && sa != sb) return sa > sb;
  return pa < pb;
}

// post_processing_mr_nms on n <= 64 rows held in shared memory; executed by one full warp.
// mode 0 normal / 1 linear.  Results in o_* (final order) and o_src (source row of each).
// Order-preserving key of a score for torch.argmax semantics: NaN is the maximum, -0.0 == +0.0.
__device__ __forceinline__ unsigned score_key(float v) {
  if (v != v) return 0xFFFFFFFFu;
  if (v == 0.f) return 0x80000000u;
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ void nms_f32_warp(NmsSmem& S, int n, float thd, int mode, int lane) {
  // Rows stay in registers (lane l owns rows l and l + 32); the reference's row swaps
  // (nncore.swap_element) are tracked as a position per row, so one step of its serial loop -
  // first argmax over positions >= i, swap, IoU suppression of the positions behind - is four
  // warp reductions (redux.sync) plus two IoU evaluations per lane, with no shared-memory traffic.
  const unsigned FULL = 0xffffffffu;
  float st[2] = {0.f, 0.f}, ed[2] = {0.f, 0.f}, sc[2] = {0.f, 0.f};
  int sr[2] = {0, 0};
  unsigned pos[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};   // current position of the row; 0xFFFFFFFF = no row
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = lane + 32 * h;
    if (r < n) { st[h] = S.st[r]; ed[h] = S.ed[r]; sc[h] = S.sc[r]; sr[h] = S.src[r]; pos[h] = r; }
  }
  // Everything that does not change inside the serial loop is hoisted: the order-preserving score keys
  // (recomputed only when a score changes), the window lengths, the guard band of the IoU test.
  unsigned key[2];
  float len[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    key[h] = score_key(sc[h]);
    len[h] = __fsub_rn(ed[h], st[h]);
  }
  const float thd_hi = thd * (1.f + 1e-6f), thd_lo = thd * (1.f - 1e-6f);
  for (int i = 0; i < n; ++i) {
    const unsigned ui = static_cast<unsigned>(i);
    // NORMAL mode: once every row still in play has score +0 the remaining steps are no-ops (equal keys:
    // the arg-max is position i itself, the swap is the identity, suppression writes 0 over 0)
    if (mode == FVTG_NMS_NORMAL) {
      const bool idle0 = pos[0] == 0xFFFFFFFFu || pos[0] < ui || __float_as_uint(sc[0]) == 0u;
      const bool idle1 = pos[1] == 0xFFFFFFFFu || pos[1] < ui || __float_as_uint(sc[1]) == 0u;
      if (__all_sync(FULL, idle0 && idle1)) break;
    }
    // first argmax over positions [i, n): largest key, then smallest position
    unsigned k[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) k[h] = (pos[h] != 0xFFFFFFFFu && pos[h] >= ui) ? key[h] : 0u;
    const unsigned kmax = __reduce_max_sync(FULL, k[0] > k[1] ? k[0] : k[1]);
    unsigned cand = 0xFFFFFFFFu;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (pos[h] != 0xFFFFFFFFu && pos[h] >= ui && k[h] == kmax && pos[h] < cand) cand = pos[h];
    const unsigned bpos = __reduce_min_sync(FULL, cand);   // position of the selected row
    // the selected row's window, broadcast from its owner; then the swap of positions i <-> bpos
    unsigned sb = 0u, eb = 0u;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (pos[h] == bpos) { sb = __float_as_uint(st[h]); eb = __float_as_uint(ed[h]); }
    const float s0 = __uint_as_float(__reduce_or_sync(FULL, sb));
    const float e0 = __uint_as_float(__reduce_or_sync(FULL, eb));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (pos[h] == bpos) pos[h] = ui;
      else if (pos[h] == ui) pos[h] = bpos;
    }
    // suppression of the rows behind position i against the selected row
    const float a0 = __fsub_rn(e0, s0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (pos[h] != 0xFFFFFFFFu && pos[h] > ui) {
        float inter = __fsub_rn(fminf(e0, ed[h]), fmaxf(s0, st[h]));
        if (inter < 0.f) inter = 0.f;  // clamp(min=0), NaN stays NaN
        const float uni = __fsub_rn(__fadd_rn(a0, len[h]), inter);
        if (mode == FVTG_NMS_NORMAL) {
          // iou >= thd with the IEEE division only inside a +-1e-6 band around the threshold (and for
          // NaN): the fast quotient is within 2 ulp, so outside the band both agree
          const float qa = __fdividef(inter, uni);
          bool sup;
          if (qa > thd_hi) sup = true;
          else if (qa < thd_lo) sup = false;
          else sup = __fdiv_rn(inter, uni) >= thd;   // NaN compares false: kept
          if (sup) { sc[h] = 0.f; key[h] = 0x80000000u; }
        } else {
          const float iou = __fdiv_rn(inter, uni);
          sc[h] = __fmul_rn(sc[h], __fsub_rn(1.f, iou));
          key[h] = score_key(sc[h]);
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (pos[h] != 0xFFFFFFFFu) { S.st[pos[h]] = st[h]; S.ed[pos[h]] = ed[h]; S.sc[pos[h]] = sc[h]; S.src[pos[h]] = sr[h]; }
  __syncwarp();
  // final descending sort, stable by current position (rank by counting)
  for (int r = lane; r < n; r += 32) {
    const float s = S.sc[r];
    int rank = 0;
    for (int q = 0; q < n; ++q)
      if (q != r && sorts_before(S.sc[q], q, s, r)) ++rank;
    S.o_st[rank] = S.st[r]; S.o_ed[rank] = S.ed[r]; S.o_sc[rank] = s; S.o_src[rank] = S.src[r];
  }
  if (lane == 0) S.o_cnt = n;
  __syncwarp();
}

// utils/temporal_nms.py on n <= 64 rows, fp64 arithmetic (python floats), one full warp.
// Input rows in H.st / H.ed / H.sc; result: H.out[0..H.cnt) = source rows kept, in output order.
struct HullSmem {
  double st[NMS_MAX], ed[NMS_MAX], sc[NMS_MAX];
  int idx[NMS_MAX];   // rows in stable descending score order
  int out[NMS_MAX];
  int cnt;
};

__device__ void nms_hull_warp(HullSmem& H, int n, double thd, int max_after, int lane) {
  if (n == 1) {  // temporal_nms.py:37-38: a single prediction is returned as is
    if (lane == 0) { H.out[0] = 0; H.cnt = 1; }
    __syncwarp();
    return;
  }
  // stable descending sort by score (python sorted(reverse=True) keeps input order on ties)
  for (int r = lane; r < n; r += 32) {
    const double s = H.sc[r];
    int rank = 0;
    for (int q = 0; q < n; ++q) {
      const double sq = H.sc[q];
      if (q != r && (sq > s || (sq == s && q < r))) ++rank;
    }
    H.idx[rank] = r;
  }
  __syncwarp();
  unsigned long long dead = 0ull;
  int alive = n, cnt = 0;
  for (int a = 0; a < n; ++a) {
    if ((dead >> a) & 1ull) continue;
    if (cnt >= max_after) break;
    if (alive > 1) {
      const double s0 = H.st[H.idx[a]], e0 = H.ed[H.idx[a]];
      for (int half = 0; half < 2; ++half) {
        const int bidx = lane + 32 * half;
        bool kill = false;
        if (bidx > a && bidx < n && !((dead >> bidx) & 1ull)) {
          const double s1 = H.st[H.idx[bidx]], e1 = H.ed[H.idx[bidx]];
          double inter = __dsub_rn(fmin(e0, e1), fmax(s0, s1));
          if (inter < 0.0) inter = 0.0;
          const double uni = __dsub_rn(fmax(e0, e1), fmin(s0, s1));
          const double iou = (uni == 0.0) ? 0.0 : __ddiv_rn(inter, uni);
          kill = iou > thd;
        }
        const unsigned m = __ballot_sync(0xffffffffu, kill);
        dead |= static_cast<unsigned long long>(m) << (32 * half);
        alive -= __popc(m);
      }
    }
    if (lane == 0) H.out[cnt] = H.idx[a];
    dead |= 1ull << a;
    --alive;
    ++cnt;
  }
  if (lane == 0) H.cnt = cnt;
  __syncwarp();
}

__device__ void run_nms_and_store(NmsSmem& S, HullSmem& H, int n, int mode, double thd,
                                  int max_after, int M, float* out_w, int* out_order,
                                  int* out_count, int lane) {
  if (mode == FVTG_NMS_HULL) {
    for (int r = lane; r < n; r += 32) { H.st[r] = S.st[r]; H.ed[r] = S.ed[r]; H.sc[r] = S.sc[r]; }
    __syncwarp();
    nms_hull_warp(H, n, thd, max_after, lane);
    for (int r = lane; r < H.cnt; r += 32) {
      const int j = H.out[r];
      S.o_st[r] = S.st[j]; S.o_ed[r] = S.ed[j]; S.o_sc[r] = S.sc[j]; S.o_src[r] = S.src[j];
    }
    if (lane == 0) S.o_cnt = H.cnt;
    __syncwarp();
  } else {
    nms_f32_warp(S, n, static_cast<float>(thd), mode, lane);
  }
  const int cnt = S.o_cnt;
  for (int r = lane; r < M; r += 32) {
    const bool v = r < cnt;
    if (out_w) {
      out_w[r * 3 + 0] = v ? S.o_st[r] : 0.f;
      out_w[r * 3 + 1] = v ? S.o_ed[r] : 0.f;
      out_w[r * 3 + 2] = v ? S.o_sc[r] : 0.f;
    }
    if (out_order) out_order[r] = v ? S.o_src[r] : -1;
  }
  if (lane == 0 && out_count) *out_count = cnt;
}

constexpr size_t DEC_NMS_BYTES = (sizeof(NmsSmem) + 15) / 16 * 16;

__global__ void __launch_bounds__(DEC_THREADS)
decode_nms_kernel(const FvtgDecodeParams p, const int Lv, const int n_max, const int npow2,
                  const float* __restrict__ cls, const float* __restrict__ conf,
                  const float* __restrict__ coord, const int* __restrict__ vlen,
                  const float* __restrict__ duration, const FvtgDecodeOut out, long long* trace) {
#define DEC_TRACE(ev) do { if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[2048 + (ev)] = clock64(); } while (0)
  extern __shared__ __align__(16) uint8_t dec_smem[];
  DEC_TRACE(0);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dec_smem);
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(dec_smem + static_cast<size_t>(npow2) * 8);
  HullSmem& H = *reinterpret_cast<HullSmem*>(dec_smem + static_cast<size_t>(npow2) * 8 + DEC_NMS_BYTES);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int vl = vlen[b];
  int offs[FVTG_MAX_LEVELS + 1];
  offs[0] = 0;
#pragma unroll
  for (int l = 0; l < FVTG_MAX_LEVELS; ++l)
    offs[l + 1] = offs[l] + (l < p.num_levels ? (vl >> l) : 0);
  const int N = offs[FVTG_MAX_LEVELS] < n_max ? offs[FVTG_MAX_LEVELS] : n_max;
  const float omx = __fsub_rn(1.f, p.x);
  const float* cl = cls + static_cast<size_t>(b) * n_max;
  const float* cf = conf + static_cast<size_t>(b) * n_max;
  // ASR mix + sigmoid -> sort keys (score bits, then lower point index first)
  for (int n = tid; n < npow2; n += DEC_THREADS) {
    unsigned long long key = 0ull;
    if (n < N) {
      const float logit = __fadd_rn(__fmul_rn(p.x, cl[n]), __fmul_rn(omx, cf[n]));
      const float score = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-logit)));
      key = (static_cast<unsigned long long>(__float_as_uint(score)) << 32) |
            static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(n));
    }
    keys[n] = key;
  }
  __syncthreads();
  DEC_TRACE(1);
  const int topk = p.topk;
  const int top = N < topk ? N : topk;
  const unsigned long long* top_keys = keys;
  if (N <= 512) {
    // Few candidates (155 points for a 75-clip video): rank by counting instead of a bitonic network -
    // one pass of N broadcast reads per key and ONE barrier instead of 36 barrier-separated stages.
    // Keys are unique (the point index is part of the key), so ranks are a permutation.
    unsigned long long* tk = reinterpret_cast<unsigned long long*>(
        dec_smem + static_cast<size_t>(npow2) * 8 + DEC_NMS_BYTES + sizeof(HullSmem));
    for (int n = tid; n < N; n += DEC_THREADS) {
      const unsigned long long key = keys[n];
      int rank = 0;
      for (int m = 0; m < N; ++m) rank += keys[m] > key;
      if (rank < top) tk[rank] = key;
    }
    __syncthreads();
    top_keys = tk;
  } else {
    // bitonic sort, descending
    for (int k = 2; k <= npow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < npow2; i += DEC_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = keys[i], c = keys[ixj];
            const bool desc = (i & k) == 0;
            if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
  }
  DEC_TRACE(2);
  const float dur = duration ? duration[b] : 3.0e38f;
  if (tid < NMS_MAX) {
    float st = 0.f, ed = 0.f, sc = 0.f;
    if (tid < top) {
      const unsigned long long key = top_keys[tid];
      const int n = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFull));
      sc = __uint_as_float(static_cast<unsigned>(key >> 32));
      int l = 0;
#pragma unroll
      for (int i = 1; i < FVTG_MAX_LEVELS; ++i)
        if (i < p.num_levels && n >= offs[i] && offs[i + 1] > offs[i]) l = i;
      const float stride = static_cast<float>(1 << l);
      const float t = static_cast<float>((n - offs[l]) << l);
      const float d0 = coord[(static_cast<size_t>(b) * n_max + n) * 2 + 0];
      const float d1 = coord[(static_cast<size_t>(b) * n_max + n) * 2 + 1];
      // model.py:257-260: b[:,0] *= -1 ; b *= stride ; b += t ; b /= (1/clip_len)
      st = __fdiv_rn(__fadd_rn(__fmul_rn(-d0, stride), t), p.inv_clip_len);
      ed = __fdiv_rn(__fadd_rn(__fmul_rn(d1, stride), t), p.inv_clip_len);
    }
    if (tid < topk) {
      if (out.boundary) {
        float* o = out.boundary + (static_cast<size_t>(b) * topk + tid) * 3;
        o[0] = st; o[1] = ed; o[2] = sc;
      }
      // inference.py:286-290: clamp all three columns to [0, duration], 4-dp
      float w0 = round4_f32(clampf(st, 0.f, dur));
      float w1 = round4_f32(clampf(ed, 0.f, dur));
      float w2 = round4_f32(clampf(sc, 0.f, dur));
      // postprocessing.py:38-50
      if (p.clip_ts) { w0 = clampf(w0, p.min_ts, p.max_ts); w1 = clampf(w1, p.min_ts, p.max_ts); }
      if (p.round_multiple) {
        w0 = __fmul_rn(rintf(__fdiv_rn(w0, p.clip_len)), p.clip_len);
        w1 = __fmul_rn(rintf(__fdiv_rn(w1, p.clip_len)), p.clip_len);
      }
      w2 = round4_f32(w2);
      if (tid >= top) { w0 = w1 = w2 = 0.f; }
      if (out.windows) {
        float* o = out.windows + (static_cast<size_t>(b) * topk + tid) * 3;
        o[0] = w0; o[1] = w1; o[2] = w2;
      }
      S.st[tid] = w0; S.ed[tid] = w1; S.sc[tid] = w2; S.src[tid] = tid;
    }
  }
  if (tid == 0 && out.count) out.count[b] = top;
  __syncthreads();
  DEC_TRACE(3);
  if (p.nms_mode != FVTG_NMS_NONE && tid < 32) {
    run_nms_and_store(S, H, top, p.nms_mode, p.nms_thd, p.max_after_nms, topk,
                      out.nms_windows ? out.nms_windows + static_cast<size_t>(b) * topk * 3 : nullptr,
                      out.nms_order ? out.nms_order + static_cast<size_t>(b) * topk : nullptr,
                      out.nms_count ? out.nms_count + b : nullptr, tid);
  }
  DEC_TRACE(4);
}

__global__ void __launch_bounds__(32)
temporal_nms_kernel(const float* __restrict__ windows, const int* __restrict__ count, int M,
                    double thd, int mode, int max_after, float* __restrict__ out_w,
                    int* __restrict__ order, int* __restrict__ out_count) {
  __shared__ NmsSmem S;
  __shared__ HullSmem H;
  const int b = blockIdx.x, lane = threadIdx.x;
  int n = count ? count[b] : M;
  if (n > M) n = M;
  if (n < 0) n = 0;
  for (int r = lane; r < n; r += 32) {
    const float* w = windows + (static_cast<size_t>(b) * M + r) * 3;
    S.st[r] = w[0]; S.ed[r] = w[1]; S.sc[r] = w[2]; S.src[r] = r;
  }
  __syncwarp();
  if (n == 0) {
    for (int r = lane; r < M; r += 32) {
      if (out_w) { float* o = out_w + (static_cast<size_t>(b) * M + r) * 3; o[0] = o[1] = o[2] = 0.f; }
      if (order) order[static_cast<size_t>(b) * M + r] = -1;
    }
    if (lane == 0 && out_count) out_count[b] = 0;
    return;
  }
  run_nms_and_store(S, H, n, mode, thd, max_after, M,
                    out_w ? out_w + static_cast<size_t>(b) * M * 3 : nullptr,
                    order ? order + static_cast<size_t>(b) * M : nullptr,
                    out_count ? out_count + b : nullptr, lane);
}

// utils/temporal_nms.py on caller fp64 rows (python floats): order / count only.
__global__ void __launch_bounds__(32)
temporal_nms_hull_f64_kernel(const double* __restrict__ windows, const int* __restrict__ count,
                             int M, double thd, int max_after, int* __restrict__ order,
                             int* __restrict__ out_count) {
  __shared__ HullSmem H;
  const int b = blockIdx.x, lane = threadIdx.x;
  int n = count ? count[b] : M;
  if (n > M) n = M;
  if (n < 0) n = 0;
  for (int r = lane; r < n; r += 32) {
    const double* w = windows + (static_cast<size_t>(b) * M + r) * 3;
    H.st[r] = w[0]; H.ed[r] = w[1]; H.sc[r] = w[2];
  }
  if (lane == 0) H.cnt = 0;
  __syncwarp();
  if (n > 0) nms_hull_warp(H, n, thd, max_after, lane);
  const int cnt = H.cnt;
  for (int r = lane; r < M; r += 32) order[static_cast<size_t>(b) * M + r] = r < cnt ? H.out[r] : -1;
  if (lane == 0 && out_count) out_count[b] = cnt;
}

int launch_temporal_nms_hull_f64(cudaStream_t st, const double* windows, const int* count, int B,
                                 int M, double thd, int max_after, int* order, int* out_count) {
  if (B <= 0) return FVTG_OK;
  if (M < 1 || M > NMS_MAX) return fail(FVTG_EINVAL, "temporal_nms: M must be 1..64");
  ProfScope prof(st, PC_DECODE);
  temporal_nms_hull_f64_kernel<<<B, 32, 0, st>>>(windows, count, M, thd, max_after, order, out_count);
  FVTG_LAUNCH_CHECK("temporal_nms_hull_f64_kernel");
  return FVTG_OK;
}

int launch_decode_nms(cudaStream_t st, const FvtgDecodeParams& p, int B, int Lv, int n_max,
                      const float* cls, const float* conf, const float* coord, const int* vlen,
                      const float* duration, const FvtgDecodeOut& out) {
  if (B <= 0) return FVTG_OK;
  if (p.topk < 1 || p.topk > NMS_MAX) return fail(FVTG_EINVAL, "decode: topk must be 1..64");
  if (n_max < 1 || n_max > 4096) return fail(FVTG_EINVAL, "decode: n_max %d out of range", n_max);
  if (p.nms_mode < FVTG_NMS_NONE || p.nms_mode > FVTG_NMS_HULL)
    return fail(FVTG_EINVAL, "decode: unknown nms_mode %d", p.nms_mode);
  int npow2 = 64;
  while (npow2 < n_max) npow2 <<= 1;
  const size_t smem = static_cast<size_t>(npow2) * 8 + DEC_NMS_BYTES + sizeof(HullSmem) + NMS_MAX * 8 /*top keys*/;
  ProfScope prof(st, PC_DECODE);
  decode_nms_kernel<<<B, DEC_THREADS, smem, st>>>(p, Lv, n_max, npow2, cls, conf, coord, vlen,
                                                   duration, out, dbg_trace());
  FVTG_LAUNCH_CHECK("decode_nms_kernel");
  return FVTG_OK;
}

int launch_temporal_nms(cudaStream_t st, const float* windows, const int* count, int B, int M,
                        double thd, int mode, int max_after, float* out_windows, int* order,
                        int* out_count) {
  if (B <= 0) return FVTG_OK;
  if (M < 1 || M > NMS_MAX) return fail(FVTG_EINVAL, "temporal_nms: M must be 1..64");
  if (mode < FVTG_NMS_NORMAL || mode > FVTG_NMS_HULL)
    return fail(FVTG_EINVAL, "temporal_nms: unknown mode %d", mode);
  ProfScope prof(st, PC_DECODE);
  temporal_nms_kernel<<<B, 32, 0, st>>>(windows, count, M, thd, mode, max_after, out_windows,
                                         order, out_count);
  FVTG_LAUNCH_CHECK("temporal_nms_kernel");
  return FVTG_OK;
}

}  // namespace fvtg
