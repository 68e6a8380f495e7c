// Launchers of the non-GEMM kernels of the hot path (definitions in prep.cu, attn.cu,
// saliency.cu, decode_nms.cu).  All take device pointers + a stream, allocate nothing, never sync.
#pragma once
#include "common.cuh"

namespace fvtg {

// ---- prep.cu : HBM-bound staging ------------------------------------------------------------
// inproj.cu: LayerNorm(raw dim) folded into the first projection GEMM, fp32 features read once:
// out bf16 [rows][256] = bf16(LN_256(relu(LN_dim(x) . W^T + b))); wg = bf16(W * diag(gamma)) [256][dim_pad].
int launch_inproj(cudaStream_t st, const float* x, int rows, int dim, int dim_pad, const void* wg,
                  const float* wsum, const float* cfold, const float* g1, const float* b1, bf16* out);
// Sine position table (position_encoding.py:61-72) for every video row of the chunk:
// pos fp32 [B*Lv][256] in the tile-blocked layout; rows >= vlen[b] are zero.
// compact: all videos of the chunk share vlen[0] -> pos is the [64][Lv][4] table (see prep.cu)
int launch_posenc(cudaStream_t st, float* pos, const int* vlen, int B, int Lv, bool compact);
// Initial dummy-encoder stream: rows [b*S + j], j < nd, of X (fp32, tile-blocked), Xb = bf16(X),
// XPb = bf16(X + dummy_pos); and the [dummy_pos ‖ 0] position table pos_d fp32 [S][256].
int launch_fill_dummy(cudaStream_t st, const float* dtok, const float* dpos, float* X, bf16* Xb,
                      bf16* XPb, float* pos_d, int B, int S, int nd);
// Pyramid level 0 (blocks.py:35, in-place ReLU): F fp32 [B*Lv][256] -> relu -> bf16 into the chain
// buffer (pitch P0, pad rows zero) and the two head layouts H1 / H2.
int launch_level0(cudaStream_t st, const float* F, bf16* chain0, bf16* H1, bf16* H2, int B, int Lv,
                  const PyrGeo& geo, bool blocked);

// ---- layer.cu : fused out_proj + LN1 + FFN + LN2 (tcgen05, hidden activations stay on-chip) ------
enum LayerMode { LAYER_T2V = 0, LAYER_SA = 1 };
struct LayerArgs {
  int M;       // rows of the stream (B * L)
  int mode;    // LAYER_T2V: FFN residual is the pre-LN1 sum ; LAYER_SA: it is the LN1 output
  float prelu;
  const float *bo, *g1, *be1, *b1, *b2, *g2, *be2;
  float* yf;         // fp32 residual stream, tile-blocked (blk_off), read then overwritten
  // deferred LayerNorm-2 across kernels: a layer whose only fp32 consumer is the next layer kernel leaves the RAW
  // sums in yf plus (rstd, -mean * rstd) per row in stats_out, and the consumer applies ((raw * rstd - mean * rstd)
  // * pg + pbe) while it reads its residual (stats_in / pg / pbe = the producer's stats_out / norm2 gamma / beta).
  const float2* stats_in;   // null: yf holds final values
  const float *pg, *pbe;
  float2* stats_out;        // null: this kernel writes y itself
  // tile-granular hand-off to the next kernel of the stream (attn.cu, AttnArgs::q_flags): tile_flags[t] = flag_epoch
  // is released once every global store of tile t has been issued behind a CTA barrier; null = not published
  int* tile_flags;
  int flag_epoch;
  bf16* out_b;       // bf16(y) [M][256] row-major or null
  bf16* out_pb;      // bf16(y + pos) [M][256] or null
  const float* pos;  // pos_mod == 0: tile-blocked per-row table ; > 0: row-major [pos_mod][256] ; null = 0
  int pos_mod;
  int pos_cmp_L;     // > 0: pos is the compact table [64][pos_cmp_L][4], row -> row % pos_cmp_L (uniform-length chunk)
  int pos_rowlim;    // > 0: out_pb only for rows with row % pos_mod < pos_rowlim
  long long* trace;  // debug: per-phase clock64 stamps of CTA 0 (fvtg_dbg_set_trace), null in production
  int trace_cta;     // debug: the CTA whose stamps are recorded (env FVTG_TRACE_CTA)
  int stagger_ns;    // start-up delay step: cluster c sleeps (c % 8) * stagger_ns before its first tile
  int dbg;           // debug (env FVTG_LAYER_DBG): bit 0 skip residual loads, bit 1 skip pos loads, bit 2 skip global stores
  int tile_rows;     // rows a tile advances by (<= 128, multiple of 8; set by launch_layer): the MMAs always run
                     // M = 128, rows past tile_rows belong to the next tile and are computed but not stored
};
int launch_layer(cudaStream_t st, const bf16* att, const bf16* wo, const bf16* w1, const bf16* w2,
                 const LayerArgs& args);
int layer_tile_rows(int M);   // rows a layer-kernel tile advances by for a stream of M rows

// fp32 tile-blocked rows -> row-major fp32: dst[(b * rows_out + j)][256] = src row (b * rows_in + j), j < rows_out
int launch_unblock(cudaStream_t st, const float* src_blk, float* dst, int B, int rows_in, int rows_out);

// ---- mlp.cu : score-head MLP chain (Linear 256->128, ReLU, [Linear 128->128, ReLU]*, Linear 128->1) fused,
// activations in TMEM; a = head conv output bf16 [M][256], w[m] bf16 [128][256 | 128], logits -> out[b][n]
struct MlpHostArgs {
  int M, nl, h2;          // rows of the head row space, hidden layers, 0 = H1 (class) / 1 = H2 (conf)
  float last_b;
  const float* bias[7];
  const float* last_w;
  float* out;
  PyrGeo geo;
};
int launch_mlp_chain(cudaStream_t st, const bf16* a, const void* const* w, const MlpHostArgs& h);

// ---- attn.cu : per-(video, head) attention on legacy warp MMA ----------------------------------
struct AttnArgs {
  const bf16* q; int ldq;   // query rows  [b*Lq + i][ldq], head h at columns h*32
  const bf16* k; int ldk;   // key rows    [b*Lk + j][ldk]
  const bf16* v; int ldv;   // value rows  [b*Lk + j][ldv]
  bf16* out;                // [b*Lq + i][256]
  int B, Lq, Lk;
  const int* klen_src;      // per-video length added to kbase: valid keys = kbase + klen_src[b]
  int kbase;
  int v_first;              // keys < v_first contribute no value (the dummy tokens, crossattention.py:385-386)
  float* tsum;              // optional fp32 [8][B*Lq] (this layer's slot): probability mass on keys >= v_first
  long long* trace;         // debug: clock64 stamps of CTA 0 / CTA 1000 at +3072 (fvtg_dbg_set_trace)
  // q (and nothing else this kernel reads) comes from the layer kernel launched just before, which publishes its
  // 128-row tiles one by one (LayerArgs::tile_flags): a CTA then waits only for the tiles that hold its video's
  // query rows instead of the whole grid, so it runs on the SMs the producer's last wave leaves idle.  null = off.
  const int* q_flags;
  int q_epoch;
  int q_tile_rows;
};
int launch_attention(cudaStream_t st, const AttnArgs& a);
// attn_tc.cu: the same contract on tcgen05 / TMEM (one CTA per video x 128-query block x head pair)
int launch_attention_tc(cudaStream_t st, const AttnArgs& a);

// ---- saliency.cu ---------------------------------------------------------------------------
// saliency (transformer.py:106-113) + t2vattnvalues finalisation (model.py:215-216).  F is the
// tile-blocked fp32 stream.
int launch_saliency(cudaStream_t st, const float* F, const int* vlen, const float* w1,
                    const float* b1, const float* w2t, const float* b2, const float* tsum,
                    int t2v_layers, float* scratch /* B * 513 floats */, float* sal_out,
                    float* t2v_out, int B, int Lv);

// ---- decode_nms.cu -------------------------------------------------------------------------
int launch_decode_nms(cudaStream_t st, const FvtgDecodeParams& p, int B, int Lv, int n_max,
                      const float* cls, const float* conf, const float* coord, const int* vlen,
                      const float* duration, const FvtgDecodeOut& out);
int launch_temporal_nms(cudaStream_t st, const float* windows, const int* count, int B, int M,
                        double thd, int mode, int max_after, float* out_windows, int* order,
                        int* out_count);

int launch_temporal_nms_hull_f64(cudaStream_t st, const double* windows, const int* count, int B,
                                 int M, double thd, int max_after, int* order, int* out_count);

}  // namespace fvtg
