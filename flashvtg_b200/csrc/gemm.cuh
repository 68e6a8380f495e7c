// The dense contraction engine of the hot path: one persistent, warp-specialised
// tcgen05 GEMM  C[M][N] = A[M][K] * W[N][K]^T  (bf16 in, fp32 accumulate in TMEM)
// with the row-wise work of the reference fused into its epilogue.
//
//   warp 0      TMA producer   (A / W tiles -> 3-stage SWIZZLE_128B smem ring)
//   warp 1      MMA issuer     (one elected lane issues tcgen05.mma, 128 x BN x 16)
//   warp 2      TMEM allocator (512 columns = two BN<=256 accumulators, double buffered)
//   warps 4..19 epilogue       (four threads per row; tcgen05.ld -> bias / activation / residual /
//                               LayerNorm / row-dot / exp; bf16 tiles leave through a swizzled
//                               staging tile + TMA store), overlapping the next tile's MMAs
//
// "Taps": K is a sequence of ntaps slabs; slab t reads A rows shifted by tap_shift[t]
// (TMA zero-fills out-of-range rows).  That turns the reference's Conv2d(1,k) /
// Conv1d(k=3) heads (model.py:53-66, blocks.py:96-101) into the same GEMM with no
// im2col, and the strided Conv1d(k=2,s=2) pyramid (blocks.py:41) is a plain GEMM on a
// [rows/2][512] view of its input.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace fvtg {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 640;       // 4 control warps + 16 epilogue warps
constexpr int GEMM_MAX_TAPS = 8;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;        // 16 KB
constexpr int GEMM_B_BYTES_MAX = 256 * GEMM_BK * 2;        // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES_MAX;
constexpr int GEMM_OUT_BYTES = 4 * GEMM_A_BYTES;           // staging tile for TMA stores: 128 x 256 bf16
constexpr int GEMM_PARAM_FLOATS = 1024 + 256 + 256 + 256;  // bias, gamma, beta, dotw
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_OUT_BYTES + 256 +
                                GEMM_PARAM_FLOATS * 4 + 4 * 128 * 8 /*LN stats*/ + 1024 /*align slack*/;
static_assert(GEMM_SMEM_BYTES <= 232448, "gemm kernel shared memory over the 227 KB limit");

enum EpiMode { EPI_ROW = 0, EPI_TILE = 1, EPI_DOT = 2, EPI_COORD = 3 };
enum RowMap { RM_NONE = 0, RM_TXT = 1, RM_CHAIN = 2, RM_H1 = 3, RM_H2 = 4 };
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_PRELU = 2 };

struct GemmEpi {
  int mode;    // EpiMode
  int rowmap;  // RowMap
  int act;     // Act applied to (acc + bias [+ res]) before LayerNorm / store
  float prelu;
  const float* bias;  // [N]
  // ---- EPI_ROW (BN == N == 256: a thread owns a full output row) ----
  const float* res;    // fp32 residual [M][256] or null
  const float* gamma;  // LayerNorm weight or null (no LN)
  const float* beta;
  int post_relu;       // ReLU after LayerNorm (pyramid)
  int f32_preln;       // out_f32 receives the pre-LayerNorm sum
  int f32_blocked;     // res / out_f32 / per-row pos use the tile-blocked layout (blk_off)
  float* out_f32;      // [rows][256] or null
  bf16* out_bf16;      // [rows][256] or null
  bf16* out_bf16_pos;  // [rows][256] or null: bf16(y + pos[row])
  const float* pos;    // fp32 [*][256] or null (treated as zero)
  int pos_mod;         // pos row = row % pos_mod when > 0
  int pos_cmp_L;       // > 0: pos is the compact table [64][pos_cmp_L][4] indexed by row % pos_cmp_L
  int pos_rowlim;      // when > 0: out_bf16_pos only for rows with row % pos_mod < pos_rowlim
  bf16* out_x1;        // extra bf16 destinations (RM_TXT: Kc; RM_CHAIN final: H1 / H2)
  bf16* out_x2;
  // ---- EPI_TILE ----
  bf16* out;  // [rows][ld_out], column offset n0
  int ld_out;
  // ---- EPI_DOT (BN == 128): out_dot[b][n] = sum_c relu(acc+bias)[c] * dotw[c] + dotb ----
  const float* dotw;
  float dotb;
  float* out_dot;
  // ---- EPI_COORD (BN == 16): out_coord[b][n][0..1] = exp(acc + bias) * coef[level] ----
  float* out_coord;
  float coef[FVTG_MAX_LEVELS];
  // ---- row map parameters ----
  int rm_a, rm_b, rm_c;  // RM_TXT: Lt, S, nd.  RM_CHAIN: step j, level l, is_final
  PyrGeo geo;
};

struct GemmArgs {
  int M, N, BN;
  int ntaps, kb_per_tap;
  int tap_shift[GEMM_MAX_TAPS];
  int a_switch_ntile;  // n-tiles >= this read A from the second tensor map
  long long* trace;    // debug (fvtg_dbg_gemm only): clock64 stamps of CTA 0, [role 3][tile 16][event 8] at +1024
  GemmEpi epi;
};

// Host launcher.  A: bf16 [a_rows][a_cols] with row pitch a_pitch (elements); a2 (optional)
// has the same geometry.  W: bf16 [N][ntaps*kb_per_tap*64].
int launch_gemm(cudaStream_t st, const void* a, const void* a2, uint64_t a_rows, uint64_t a_cols,
                uint64_t a_pitch, const void* w, const GemmArgs& args);

// Up to 4 independent GEMMs in ONE launch; problem i runs on its own sm_count / n persistent CTAs.
struct GemmOperands {
  const void* a;
  const void* a2;
  uint64_t a_rows, a_cols, a_pitch;
  const void* w;
};
int launch_gemm_group(cudaStream_t st, int n, const GemmOperands* ops, const GemmArgs* args);

inline GemmArgs gemm_args(int M, int N, int BN, int K) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.BN = BN;
  g.ntaps = 1;
  g.kb_per_tap = K / GEMM_BK;
  g.a_switch_ntile = 1 << 30;
  return g;
}

}  // namespace fvtg
