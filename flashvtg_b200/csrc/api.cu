// C-ABI entry points: validate, carve the caller's workspace, and enqueue the kernel sequence of
// the hot path on the caller's stream, chunk by chunk (a chunk = as many videos as the workspace
// budget allows, see chunk_videos).
#include "gemm.cuh"
#include "kernels.cuh"

namespace fvtg {

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static int levels_present(const FvtgCfg& c, int Lv) {
  int n = 0;
  for (int l = 0; l < c.num_levels; ++l)
    if (Lv >= (1 << l)) ++n;
  return n;
}

static PyrGeo make_geo(const FvtgCfg& c, int Lv, const int* vlen) {
  PyrGeo g;
  memset(&g, 0, sizeof(g));
  g.nlev = levels_present(c, Lv);
  const int hk = c.head_k / 2, ck = c.coord_k / 2;
  g.pad = hk > ck ? hk : ck;
  const int align = 1 << (g.nlev > 4 ? g.nlev - 1 : 3);  // >= 8 rows keeps [rows/2][512] views 16B-pitched
  g.P0 = round_up(Lv, align);
  int o = g.pad, n = 0;
  for (int l = 0; l < g.nlev; ++l) {
    g.o1[l] = o;
    o += (Lv >> l) + g.pad;
    n += Lv >> l;
  }
  g.PH1 = o;
  g.n_max = n;
  g.PH2 = g.pad + n + g.pad;
  g.vlen = vlen;
  return g;
}

// Bump allocator over the caller's workspace; with base == nullptr it only measures.
struct Carver {
  uint8_t* base;
  size_t off;
  template <typename T>
  T* take(size_t count) {
    off = round_up_sz(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct Ws {
  // fusion
  bf16 *tmp256, *Xb, *XPb, *Kc, *Yb, *YPb, *qkv, *att, *ffh;
  float *Xf, *Yf, *pos_v, *pos_d, *tsum, *sal_scratch;
  int* Yflags;             // per-128-row-tile completion epochs of the video stream's layer kernels (attn.cu q_flags)
  float2 *Xstat, *Ystat;   // per-row (rstd, -mean * rstd) of a layer kernel's deferred LayerNorm-2 (layer.cu)
  // pyramid + heads
  bf16 *chain0, *chainA[FVTG_MAX_LEVELS], *chainB[FVTG_MAX_LEVELS], *H1, *H2, *hA, *hB, *mA, *mB;
  // per-chunk head logits when the caller does not want them
  float *cls, *conf, *coord;
  size_t h1_rows, h2_rows;
};

static size_t carve(const FvtgCfg& c, int Bc, int Lv, int Lt, uint8_t* base, Ws* w) {
  Carver k{base, 0};
  const size_t S = c.num_dummies + Lt;
  const size_t Rv = static_cast<size_t>(Bc) * Lv, Rs = Bc * S;
  const size_t Rmax = Rv > Rs ? Rv : Rs;
  PyrGeo g = make_geo(c, Lv, nullptr);
  Ws t;
  t.tmp256 = k.take<bf16>(Rmax * 256);
  t.Xf = k.take<float>(round_up_sz(Rs, 128) * 256);   // tile-blocked
  t.Xb = k.take<bf16>(Rs * 256);
  t.XPb = k.take<bf16>(Rs * 256);
  t.Kc = k.take<bf16>(Rs * 256);
  t.Yf = k.take<float>(round_up_sz(Rv, 128) * 256);   // tile-blocked
  t.Yb = k.take<bf16>(Rv * 256);
  t.Xstat = k.take<float2>(round_up_sz(Rs, 128));
  t.Ystat = k.take<float2>(round_up_sz(Rv, 128));
  t.Yflags = k.take<int>(Rv / 8 + 64);   // >= one flag per tile for tiles of >= 8 rows
  t.YPb = k.take<bf16>(Rv * 256);
  t.pos_v = k.take<float>(round_up_sz(Rv, 128) * 256);  // tile-blocked
  t.pos_d = k.take<float>(S * 256);
  t.qkv = k.take<bf16>(Rmax * 768);
  t.att = k.take<bf16>(Rmax * 256);
  t.ffh = k.take<bf16>(Rmax * 1024);
  t.tsum = k.take<float>(8 * Rv * (c.t2v_layers > 0 ? c.t2v_layers : 1));  // [layer][head][row]
  t.sal_scratch = k.take<float>(static_cast<size_t>(Bc) * 513 + 64);
  const size_t Rc = static_cast<size_t>(Bc) * g.P0;
  t.chain0 = k.take<bf16>(Rc * 256);
  // pyramid chains run level-parallel (one grouped launch per step), so every level >= 2 keeps its own
  // ping-pong pair: A for odd steps (<= Rc / 2 rows), B for even steps (<= Rc / 4 rows)
  for (int l = 0; l < FVTG_MAX_LEVELS; ++l) {
    t.chainA[l] = nullptr;
    t.chainB[l] = nullptr;
    if (l >= 2 && l < g.nlev) {
      t.chainA[l] = k.take<bf16>(Rc / 2 * 256 + 256);
      t.chainB[l] = k.take<bf16>(Rc / 4 * 256 + 256);
    }
  }
  t.h1_rows = static_cast<size_t>(Bc) * g.PH1;
  t.h2_rows = static_cast<size_t>(Bc) * g.PH2;
  const size_t Rh = t.h1_rows > t.h2_rows ? t.h1_rows : t.h2_rows;
  t.H1 = k.take<bf16>(t.h1_rows * 256);
  t.H2 = k.take<bf16>(t.h2_rows * 256);
  t.hA = k.take<bf16>(Rh * 256);
  t.hB = k.take<bf16>(Rh * 256);
  t.mA = k.take<bf16>(Rh * 128);
  t.mB = k.take<bf16>(Rh * 128);
  t.cls = k.take<float>(static_cast<size_t>(Bc) * g.n_max);
  t.conf = k.take<float>(static_cast<size_t>(Bc) * g.n_max);
  t.coord = k.take<float>(static_cast<size_t>(Bc) * g.n_max * 2);
  if (w) *w = t;
  return round_up_sz(k.off, 256);
}

static int chunk_videos(const FvtgCfg& c, int Lv, int Lt) {
  const int forced = env_int("FVTG_CHUNK", 0);
  if (forced > 0) return forced;
  // Large chunks: every kernel of the sequence is latency / epilogue bound below a few waves of
  // 128-row tiles (measured: 34 k videos/s at 78-video chunks vs 99 k at 1024, profiles/r01_*),
  // so a chunk is capped only by row count and workspace size, not by L2 residency.
  const size_t per_video = carve(c, 1, Lv, Lt, nullptr, nullptr);
  const size_t budget = static_cast<size_t>(env_int("FVTG_WS_MB", 4096)) << 20;
  long long bc = 131072 / (Lv > 0 ? Lv : 1);
  if (bc < 1) bc = 1;
  if (static_cast<size_t>(bc) * per_video > budget) bc = static_cast<long long>(budget / per_video);
  if (bc < 1) bc = 1;
  return static_cast<int>(bc);
}

static int check_cfg(const FvtgCfg* c) {
  if (!c) return fail(FVTG_EINVAL, "null cfg");
  if (c->abi_version != FVTG_ABI_VERSION) return fail(FVTG_EINVAL, "cfg.abi_version mismatch");
  if (c->v_dim < 1 || c->t_dim < 1 || c->v_dim_pad % 64 || c->t_dim_pad % 64 ||
      c->v_dim_pad < c->v_dim || c->t_dim_pad < c->t_dim)
    return fail(FVTG_EINVAL, "cfg: feature dims must be padded to multiples of 64");
  if (c->num_dummies < 1 || c->num_dummies > 256) return fail(FVTG_EINVAL, "cfg: num_dummies");
  if (c->dummy_layers < 1 || c->dummy_layers > FVTG_MAX_LAYERS || c->t2v_layers < 0 ||
      c->t2v_layers > FVTG_MAX_LAYERS || c->enc_layers < 0 || c->enc_layers > FVTG_MAX_LAYERS)
    return fail(FVTG_EINVAL, "cfg: dummy_layers 1..%d, t2v/enc layers 0..%d", FVTG_MAX_LAYERS, FVTG_MAX_LAYERS);
  if (c->num_levels < 1 || c->num_levels > FVTG_MAX_LEVELS) return fail(FVTG_EINVAL, "cfg: num_levels");
  if (c->head_k < 1 || c->head_k > 7 || !(c->head_k & 1) || c->coord_k != 3)
    return fail(FVTG_EINVAL, "cfg: head_k must be odd <= 7 and coord_k == 3");
  if (c->num_conv_layers < 1 || c->num_conv_layers > FVTG_MAX_CONVS || c->num_mlp_layers < 2 ||
      c->num_mlp_layers > FVTG_MAX_MLP)
    return fail(FVTG_EINVAL, "cfg: num_conv_layers 1..4, num_mlp_layers 2..8");
  return FVTG_OK;
}

// ---------------------------------------------------------------------------------------------
// prev: the layer kernel that produced yf and deferred its LayerNorm-2 to this one (null: yf is final);
// stat: the stream's statistics buffer; defer: the next reader of yf is another layer kernel, which normalises.
static LayerArgs layer_args(const FvtgEncLayer& L, int rows, int mode, float* yf, const FvtgEncLayer* prev,
                            float2* stat, bool defer) {
  LayerArgs a;
  memset(&a, 0, sizeof(a));
  if (prev) {
    a.stats_in = stat;
    a.pg = prev->norm2.g;
    a.pbe = prev->norm2.b;
  }
  a.stats_out = defer ? stat : nullptr;
  a.M = rows;
  a.mode = mode;
  a.prelu = L.prelu;
  a.bo = L.out_proj.b;
  a.g1 = L.norm1.g; a.be1 = L.norm1.b;
  a.b1 = L.ff1.b;
  a.b2 = L.ff2.b;
  a.g2 = L.norm2.g; a.be2 = L.norm2.b;
  a.yf = yf;
  a.trace = dbg_trace();
  return a;
}

// One post-norm self-attention layer (transformer.py:408-421) over `rows` = B * L stream rows:
// QKV projection GEMM -> per-(video, head) attention -> fused out_proj/LN1/FFN/LN2 kernel.
// out_b / out_pb: bf16(x) and bf16(x + pos) for the next layer's V and Q/K projections.
static int sa_layer(cudaStream_t st, const FvtgEncLayer& L, const Ws& w, int B, int Lseq, float* xf,
                    const bf16* xb, const bf16* xpb, bf16* out_b, bf16* out_pb, const float* pos,
                    int pos_mod, int pos_rowlim, const int* klen_src, int kbase, const FvtgEncLayer* prev,
                    float2* stat, bool defer, int pos_cmp_L = 0) {
  const int rows = B * Lseq;
  {  // Q,K from x+pos ; V from x  (in_proj rows 0:512 / 512:768)
    GemmArgs g = gemm_args(rows, 768, 256, 256);
    g.a_switch_ntile = 2;
    g.epi.mode = EPI_TILE;
    g.epi.bias = L.in_proj.b;
    g.epi.out = w.qkv;
    g.epi.ld_out = 768;
    FVTG_TRY(launch_gemm(st, xpb, xb, rows, 256, 256, L.in_proj.w, g));
  }
  {
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = w.qkv; a.ldq = 768;
    a.k = w.qkv + 256; a.ldk = 768;
    a.v = w.qkv + 512; a.ldv = 768;
    a.out = w.att;
    a.B = B; a.Lq = Lseq; a.Lk = Lseq;
    a.klen_src = klen_src; a.kbase = kbase; a.v_first = 0; a.tsum = nullptr;
    a.trace = dbg_trace();
    FVTG_TRY(launch_attention(st, a));
  }
  LayerArgs a = layer_args(L, rows, LAYER_SA, xf, prev, stat, defer);
  a.out_b = out_b;
  a.out_pb = out_pb;
  a.pos = pos;
  a.pos_mod = pos_mod;
  a.pos_cmp_L = pos_cmp_L;
  a.pos_rowlim = pos_rowlim;
  return launch_layer(st, w.att, static_cast<const bf16*>(L.out_proj.w),
                      static_cast<const bf16*>(L.ff1.w), static_cast<const bf16*>(L.ff2.w), a);
}

// FVTG_TILE_HANDOFF=0 restores the grid-wide dependency between a T2V layer kernel and the next attention
static bool tile_handoff() {
  static const bool on = [] { const char* e = getenv("FVTG_TILE_HANDOFF"); return !e || atoi(e) != 0; }();
  return on;
}

// One adaptive cross-attention layer (transformer.py:334-369, crossattention.py:287-396):
// attention over the constant [dummies ‖ text] keys, then the fused layer kernel.
static int t2v_layer(cudaStream_t st, const FvtgCfg& c, const FvtgEncLayer& L, const Ws& w, int B,
                     int Lv, int S, const int* tlen, bool want_b, int layer, int pos_cmp_L,
                     const FvtgEncLayer* prev, bool defer) {
  const int rows = B * Lv;
  {
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = w.YPb; a.ldq = 256;
    a.k = w.Kc; a.ldk = 256;
    a.v = w.Kc; a.ldv = 256;
    a.out = w.att;
    a.B = B; a.Lq = Lv; a.Lk = S;
    a.klen_src = tlen; a.kbase = c.num_dummies; a.v_first = c.num_dummies;
    a.tsum = w.tsum + static_cast<size_t>(layer) * 8 * B * Lv;
    a.trace = dbg_trace();
    if (layer > 0 && tile_handoff()) {   // q = YPb comes from the previous T2V layer kernel, tile by tile
      a.q_flags = w.Yflags;
      a.q_epoch = layer;
      a.q_tile_rows = layer_tile_rows(rows);
    }
    FVTG_TRY(launch_attention(st, a));
  }
  LayerArgs a = layer_args(L, rows, LAYER_T2V, w.Yf, prev, w.Ystat, defer);
  if (tile_handoff()) {
    a.tile_flags = w.Yflags;
    a.flag_epoch = layer + 1;
  }
  a.out_b = want_b ? w.Yb : nullptr;
  a.out_pb = w.YPb;
  a.pos = w.pos_v;
  a.pos_cmp_L = pos_cmp_L;
  return launch_layer(st, w.att, static_cast<const bf16*>(L.out_proj.w),
                      static_cast<const bf16*>(L.ff1.w), static_cast<const bf16*>(L.ff2.w), a);
}

// LinearLayer x2 (model.py:99-110,782-789) for one modality.
static int in_proj(cudaStream_t st, const FvtgInProj& P, const Ws& w, const float* src, int rows,
                   int dim, int dim_pad, GemmEpi final_epi) {
  // layer 0 with its LayerNorm over the raw dim folded in, + ReLU + LayerNorm(256) of layer 1:
  // tmp256 = bf16(LN1(relu(LN0(x) W0^T + b0))), one pass over the fp32 features (inproj.cu)
  FVTG_TRY(launch_inproj(st, src, rows, dim, dim_pad, P.fc0.w, P.fc0_wsum, P.fc0.b, P.ln1.g, P.ln1.b,
                         w.tmp256));
  {
    GemmArgs g = gemm_args(rows, 256, 256, 256);
    g.epi = final_epi;
    g.epi.mode = EPI_ROW;
    g.epi.bias = P.fc1.b;
    FVTG_TRY(launch_gemm(st, w.tmp256, nullptr, rows, 256, 256, P.fc1.w, g));
  }
  return FVTG_OK;
}

static int fusion_chunk(cudaStream_t st, const FvtgCfg& c, const FvtgWeights& W, const Ws& w,
                        int B, int Lv, int Lt, const float* vid, const float* txt,
                        const int* vlen, const int* tlen, float* video_emb, float* saliency,
                        float* t2v, float* dummy_tokens, bool uniform_vlen) {
  const int nd = c.num_dummies, S = nd + Lt;
  // uniform-length chunk (caller's promise, FvtgBatch.uniform_vid_len; always true for one video):
  // the sine table shrinks from B*Lv rows to Lv rows that stay cache resident
  const int pos_cmp_L = (uniform_vlen || B == 1) ? Lv : 0;
  FVTG_TRY(launch_posenc(st, w.pos_v, vlen, B, Lv, pos_cmp_L > 0));
  FVTG_TRY(launch_fill_dummy(st, W.dummy_tok, W.dummy_pos, w.Xf, w.Xb, w.XPb, w.pos_d, B, S, nd));
  {  // text: rows scattered into the [dummies ‖ text] stream, plus the constant text keys/values
    GemmEpi e;
    memset(&e, 0, sizeof(e));
    e.rowmap = RM_TXT; e.rm_a = Lt; e.rm_b = S; e.rm_c = nd;
    e.f32_blocked = 1;
    e.out_f32 = w.Xf; e.out_bf16 = w.Xb; e.out_bf16_pos = w.XPb; e.out_x1 = w.Kc;
    FVTG_TRY(in_proj(st, W.txt, w, txt, B * Lt, c.t_dim, c.t_dim_pad, e));
  }
  {  // video
    GemmEpi e;
    memset(&e, 0, sizeof(e));
    e.f32_blocked = 1;
    e.out_f32 = w.Yf; e.out_bf16 = w.Yb; e.out_bf16_pos = w.YPb; e.pos = w.pos_v;
    e.pos_cmp_L = pos_cmp_L;
    FVTG_TRY(in_proj(st, W.vid, w, vid, B * Lv, c.v_dim, c.v_dim_pad, e));
  }
  // Layer kernels of one stream hand their LayerNorm-2 to the next one (layer.cu): only the last kernel of a
  // stream, whose fp32 output other kernels read, writes normalised rows.
  for (int i = 0; i < c.dummy_layers; ++i) {
    // the last dummy layer writes bf16(dummy + dummy_pos) straight into the key rows of Kc
    const bool last = i == c.dummy_layers - 1;
    FVTG_TRY(sa_layer(st, W.dummy[i], w, B, S, w.Xf, w.Xb, w.XPb, last ? nullptr : w.Xb,
                      last ? w.Kc : w.XPb, w.pos_d, S, last ? nd : 0, tlen, nd,
                      i > 0 ? &W.dummy[i - 1] : nullptr, w.Xstat, !last));
  }
  if (dummy_tokens) FVTG_TRY(launch_unblock(st, w.Xf, dummy_tokens, B, S, nd));
  const int n_video_layers = c.t2v_layers + c.enc_layers;
  if (c.t2v_layers > 1 && tile_handoff()) {
    FVTG_CUDA_OK(cudaMemsetAsync(w.Yflags, 0, sizeof(int) * (static_cast<size_t>(B) * Lv / 8 + 64), st));
    count_launch(1);
  }
  for (int i = 0; i < c.t2v_layers; ++i)
    FVTG_TRY(t2v_layer(st, c, W.t2v[i], w, B, Lv, S, tlen, i == c.t2v_layers - 1, i, pos_cmp_L,
                       i > 0 ? &W.t2v[i - 1] : nullptr, i < n_video_layers - 1));
  for (int i = 0; i < c.enc_layers; ++i) {
    const bool last = i == c.enc_layers - 1;
    const FvtgEncLayer* prev = i > 0 ? &W.enc[i - 1] : (c.t2v_layers > 0 ? &W.t2v[c.t2v_layers - 1] : nullptr);
    FVTG_TRY(sa_layer(st, W.enc[i], w, B, Lv, w.Yf, w.Yb, w.YPb, last ? nullptr : w.Yb,
                      last ? nullptr : w.YPb, w.pos_v, 0, 0, vlen, 0, prev, w.Ystat, !last, pos_cmp_L));
  }
  FVTG_TRY(launch_saliency(st, w.Yf, vlen, W.sal_w1, W.sal_b1, W.sal_w2t, W.sal_b2, w.tsum,
                           c.t2v_layers, w.sal_scratch, saliency, t2v, B, Lv));
  if (video_emb) FVTG_TRY(launch_unblock(st, w.Yf, video_emb, B * Lv, 1, 1));
  return FVTG_OK;
}

static int score_head(cudaStream_t st, const FvtgCfg& c, const FvtgScoreHead& H, const Ws& w,
                      const bf16* in, int rows, int rowmap, const PyrGeo& geo, float* out_logit) {
  const int k = c.head_k;
  const bf16* cur = in;
  bf16* bufs[2] = {w.hA, w.hB};
  for (int ci = 0; ci < c.num_conv_layers; ++ci) {
    GemmArgs g = gemm_args(rows, 256, 256, 256);
    g.ntaps = k;
    g.kb_per_tap = 4;
    for (int t = 0; t < k; ++t) g.tap_shift[t] = t - k / 2;
    g.epi.mode = EPI_TILE;
    g.epi.rowmap = rowmap;
    g.epi.geo = geo;
    g.epi.bias = H.conv[ci].b;
    g.epi.act = ACT_RELU;
    g.epi.out = bufs[ci & 1];
    g.epi.ld_out = 256;
    FVTG_TRY(launch_gemm(st, cur, nullptr, rows, 256, 256, H.conv[ci].w, g));
    cur = bufs[ci & 1];
  }
  {  // the whole MLP behind the convs in one kernel (mlp.cu); FVTG_MLP_FUSED=0 keeps the per-layer GEMMs
    static const bool fused = [] { const char* e = getenv("FVTG_MLP_FUSED"); return !e || atoi(e) != 0; }();
    if (fused && c.num_mlp_layers - 1 <= 7) {
      MlpHostArgs h;
      memset(&h, 0, sizeof(h));
      h.M = rows; h.nl = c.num_mlp_layers - 1; h.h2 = rowmap == RM_H2 ? 1 : 0;
      h.last_b = H.last_b; h.last_w = H.last_w; h.out = out_logit; h.geo = geo;
      const void* wp[7];
      for (int m = 0; m < h.nl; ++m) { h.bias[m] = H.mlp[m].b; wp[m] = H.mlp[m].w; }
      return launch_mlp_chain(st, cur, wp, h);
    }
  }
  bf16* mb[2] = {w.mA, w.mB};
  int kin = 256;
  for (int m = 0; m < c.num_mlp_layers - 1; ++m) {
    const bool last = m == c.num_mlp_layers - 2;
    GemmArgs g = gemm_args(rows, 128, 128, kin);
    g.epi.rowmap = rowmap;
    g.epi.geo = geo;
    g.epi.bias = H.mlp[m].b;
    if (!last) {
      g.epi.mode = EPI_TILE;
      g.epi.act = ACT_RELU;
      g.epi.out = mb[m & 1];
      g.epi.ld_out = 128;
    } else {
      g.epi.mode = EPI_DOT;
      g.epi.dotw = H.last_w;
      g.epi.dotb = H.last_b;
      g.epi.out_dot = out_logit;
    }
    FVTG_TRY(launch_gemm(st, cur, nullptr, rows, kin, kin, H.mlp[m].w, g));
    cur = mb[m & 1];
    kin = 128;
  }
  return FVTG_OK;
}

static int pyramid_heads_chunk(cudaStream_t st, const FvtgCfg& c, const FvtgWeights& W, const Ws& w,
                               int B, int Lv, const float* F, bool f_blocked, const int* vlen,
                               float* cls, float* conf, float* coord) {
  PyrGeo geo = make_geo(c, Lv, vlen);
  // level0 also zeroes the rows of H1 / H2 that no producer writes (pads, positions past a video's length)
  FVTG_TRY(launch_level0(st, F, w.chain0, w.H1, w.H2, B, Lv, geo, f_blocked));
  // Temporal Feature Layering (blocks.py:52-70): level l = l strided convs from ReLU(F), own weights.
  // The chains of different levels are independent, so step j of every level l >= j goes out as ONE
  // grouped launch (4 + 3 + 2 + 1 GEMMs in 4 launches instead of 10; the level-parallel CTAs also share
  // the step-1 input through L2).  FVTG_PYR_GROUP=0 restores one launch per (level, step).
  static const bool grouped = [] { const char* e = getenv("FVTG_PYR_GROUP"); return !e || atoi(e) != 0; }();
  for (int j = 1; j < geo.nlev; ++j) {
    GemmOperands ops[FVTG_MAX_LEVELS];
    GemmArgs ga[FVTG_MAX_LEVELS];
    int n = 0;
    for (int l = j; l < geo.nlev; ++l) {
      const int rows_out = B * (geo.P0 >> j);
      GemmArgs g = gemm_args(rows_out, 256, 256, 512);
      g.epi.mode = EPI_ROW;
      g.epi.rowmap = RM_CHAIN;
      g.epi.rm_a = j; g.epi.rm_b = l; g.epi.rm_c = (j == l) ? 1 : 0;
      g.epi.geo = geo;
      g.epi.bias = W.pyr[l][j - 1].conv.b;
      g.epi.gamma = W.pyr[l][j - 1].ln.g; g.epi.beta = W.pyr[l][j - 1].ln.b;
      g.epi.post_relu = 1;
      const bf16* src = (j == 1) ? w.chain0 : (((j - 1) & 1) ? w.chainA[l] : w.chainB[l]);
      bf16* dst = (j & 1) ? w.chainA[l] : w.chainB[l];
      g.epi.out_bf16 = (j == l) ? nullptr : dst;
      g.epi.out_x1 = w.H1; g.epi.out_x2 = w.H2;
      ops[n].a = src; ops[n].a2 = nullptr;
      ops[n].a_rows = rows_out; ops[n].a_cols = 512; ops[n].a_pitch = 512;
      ops[n].w = W.pyr[l][j - 1].conv.w;
      ga[n] = g;
      ++n;
    }
    if (grouped && n <= 4) {
      FVTG_TRY(launch_gemm_group(st, n, ops, ga));
    } else {
      for (int i = 0; i < n; ++i)
        FVTG_TRY(launch_gemm(st, ops[i].a, nullptr, ops[i].a_rows, 512, 512, ops[i].w, ga[i]));
    }
  }
  FVTG_CUDA_OK(cudaMemsetAsync(cls, 0, sizeof(float) * B * geo.n_max, st));
  FVTG_CUDA_OK(cudaMemsetAsync(conf, 0, sizeof(float) * B * geo.n_max, st));
  FVTG_CUDA_OK(cudaMemsetAsync(coord, 0, sizeof(float) * B * geo.n_max * 2, st));
  count_launch(3);
  FVTG_TRY(score_head(st, c, W.cls, w, w.H1, B * geo.PH1, RM_H1, geo, cls));
  FVTG_TRY(score_head(st, c, W.conf, w, w.H2, B * geo.PH2, RM_H2, geo, conf));
  {  // coord head (blocks.py:90-105): conv k3 + ReLU, conv k3 -> 2, exp * coef[level]
    const int rows = B * geo.PH1;
    GemmArgs g = gemm_args(rows, 256, 256, 256);
    g.ntaps = 3; g.kb_per_tap = 4;
    for (int t = 0; t < 3; ++t) g.tap_shift[t] = t - 1;
    g.epi.mode = EPI_TILE;
    g.epi.rowmap = RM_H1; g.epi.geo = geo;
    g.epi.bias = W.coord1.b;
    g.epi.act = ACT_RELU;
    g.epi.out = w.hA; g.epi.ld_out = 256;
    FVTG_TRY(launch_gemm(st, w.H1, nullptr, rows, 256, 256, W.coord1.w, g));
    GemmArgs h = gemm_args(rows, 16, 16, 256);
    h.ntaps = 3; h.kb_per_tap = 4;
    for (int t = 0; t < 3; ++t) h.tap_shift[t] = t - 1;
    h.epi.mode = EPI_COORD;
    h.epi.rowmap = RM_H1; h.epi.geo = geo;
    h.epi.bias = W.coord2.b;
    h.epi.out_coord = coord;
    for (int l = 0; l < FVTG_MAX_LEVELS; ++l) h.epi.coef[l] = W.coef[l];
    FVTG_TRY(launch_gemm(st, w.hA, nullptr, rows, 256, 256, W.coord2.w, h));
  }
  return FVTG_OK;
}

static FvtgDecodeParams default_decode(const FvtgCfg& c, const FvtgWeights& W) {
  FvtgDecodeParams p;
  memset(&p, 0, sizeof(p));
  p.x = W.x;
  p.clip_len = c.clip_len;
  p.inv_clip_len = static_cast<float>(1.0 / static_cast<double>(c.clip_len));
  p.topk = c.max_num_moment;
  p.num_levels = c.num_levels;
  p.nms_mode = FVTG_NMS_NONE;
  return p;
}

}  // namespace fvtg

using namespace fvtg;

extern "C" {

size_t fvtg_workspace_bytes(const FvtgCfg* cfg, int32_t B, int32_t Lv, int32_t Lt) {
  if (check_cfg(cfg) != FVTG_OK || B < 1 || Lv < 1 || Lt < 1) return 0;
  int bc = chunk_videos(*cfg, Lv, Lt);
  if (bc > B) bc = B;
  return carve(*cfg, bc, Lv, Lt, nullptr, nullptr) + 1024;
}

int32_t fvtg_chunk_videos(const FvtgCfg* cfg, int32_t Lv, int32_t Lt) {
  if (check_cfg(cfg) != FVTG_OK || Lv < 1 || Lt < 1) return 0;
  return chunk_videos(*cfg, Lv, Lt);
}

static int prep_ws(const FvtgCfg* cfg, int B, int Lv, int Lt, void* workspace, size_t ws_bytes,
                   int* bc_out, Ws* w) {
  if (!workspace) return fail(FVTG_EINVAL, "null workspace");
  int bc = chunk_videos(*cfg, Lv, Lt);
  if (bc > B) bc = B;
  uint8_t* base = reinterpret_cast<uint8_t*>(round_up_sz(reinterpret_cast<size_t>(workspace), 1024));
  const size_t need = carve(*cfg, bc, Lv, Lt, base, w) + (base - reinterpret_cast<uint8_t*>(workspace));
  if (need > ws_bytes)
    return fail(FVTG_EWORKSPACE, "workspace too small: need %zu bytes, have %zu", need, ws_bytes);
  *bc_out = bc;
  return FVTG_OK;
}

static int check_shapes(const FvtgCfg* cfg, int B, int Lv, int Lt) {
  if (B < 1 || Lv < 1 || Lt < 1) return fail(FVTG_EINVAL, "B, Lv, Lt must be positive");
  // kernel limit (row-space geometry, top-k candidate buffers); the reference's own bound is cfg.buffer_size
  // (generator.py:60), which the host checks - for the highlight presets (buffer 2048) this limit is the tighter one
  if (Lv > 1024) return fail(FVTG_EINVAL, "Lv %d exceeds the 1024 clips per video the kernels support", Lv);
  if (cfg->num_dummies + Lt > 1024) return fail(FVTG_EINVAL, "num_dummies + Lt too large");
  return FVTG_OK;
}

int32_t fvtg_fusion_fwd(const FvtgCfg* cfg, const FvtgWeights* w, const FvtgBatch* in,
                        const FvtgFusionOut* out, void* workspace, size_t ws_bytes, void* stream) {
  host_state().launches = 0;
  FVTG_TRY(check_cfg(cfg));
  if (!w || !in || !out || !in->vid || !in->txt || !in->vid_len || !in->txt_len || !out->saliency)
    return fail(FVTG_EINVAL, "fusion_fwd: null argument");
  FVTG_TRY(check_shapes(cfg, in->B, in->Lv, in->Lt));
  FVTG_TRY(check_arch());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int bc = 0;
  Ws ws;
  FVTG_TRY(prep_ws(cfg, in->B, in->Lv, in->Lt, workspace, ws_bytes, &bc, &ws));
  const int Lv = in->Lv, Lt = in->Lt, nd = cfg->num_dummies;
  for (int b0 = 0; b0 < in->B; b0 += bc) {
    const int nb = in->B - b0 < bc ? in->B - b0 : bc;
    FVTG_TRY(fusion_chunk(
        st, *cfg, *w, ws, nb, Lv, Lt, in->vid + static_cast<size_t>(b0) * Lv * cfg->v_dim,
        in->txt + static_cast<size_t>(b0) * Lt * cfg->t_dim, in->vid_len + b0, in->txt_len + b0,
        out->video_emb ? out->video_emb + static_cast<size_t>(b0) * Lv * 256 : nullptr,
        out->saliency + static_cast<size_t>(b0) * Lv,
        out->t2v ? out->t2v + static_cast<size_t>(b0) * Lv : nullptr,
        out->dummy_tokens ? out->dummy_tokens + static_cast<size_t>(b0) * nd * 256 : nullptr,
        in->uniform_vid_len != 0));
  }
  return FVTG_OK;
}

int32_t fvtg_pyramid_heads_fwd(const FvtgCfg* cfg, const FvtgWeights* w, int32_t B, int32_t Lv,
                               const float* video_emb, const int32_t* vid_len,
                               const FvtgHeadsOut* out, void* workspace, size_t ws_bytes,
                               void* stream) {
  host_state().launches = 0;
  FVTG_TRY(check_cfg(cfg));
  if (!w || !video_emb || !vid_len || !out || !out->cls_logit || !out->conf_logit || !out->coord)
    return fail(FVTG_EINVAL, "pyramid_heads_fwd: null argument");
  FVTG_TRY(check_shapes(cfg, B, Lv, 1));
  FVTG_TRY(check_arch());
  PyrGeo g0 = make_geo(*cfg, Lv, nullptr);
  if (out->n_max != g0.n_max)
    return fail(FVTG_EINVAL, "pyramid_heads_fwd: n_max %d != %d for Lv %d", out->n_max, g0.n_max, Lv);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int bc = 0;
  Ws ws;
  FVTG_TRY(prep_ws(cfg, B, Lv, 1, workspace, ws_bytes, &bc, &ws));
  for (int b0 = 0; b0 < B; b0 += bc) {
    const int nb = B - b0 < bc ? B - b0 : bc;
    FVTG_TRY(pyramid_heads_chunk(st, *cfg, *w, ws, nb, Lv,
                                 video_emb + static_cast<size_t>(b0) * Lv * 256, false, vid_len + b0,
                                 out->cls_logit + static_cast<size_t>(b0) * g0.n_max,
                                 out->conf_logit + static_cast<size_t>(b0) * g0.n_max,
                                 out->coord + static_cast<size_t>(b0) * g0.n_max * 2));
  }
  return FVTG_OK;
}

int32_t fvtg_decode_nms(const FvtgDecodeParams* p, int32_t B, int32_t Lv, int32_t n_max,
                        const float* cls_logit, const float* conf_logit, const float* coord,
                        const int32_t* vid_len, const float* duration, const FvtgDecodeOut* out,
                        void* stream) {
  host_state().launches = 0;
  if (!p || !cls_logit || !conf_logit || !coord || !vid_len || !out)
    return fail(FVTG_EINVAL, "decode_nms: null argument");
  if (p->num_levels < 1 || p->num_levels > FVTG_MAX_LEVELS)
    return fail(FVTG_EINVAL, "decode_nms: num_levels");
  FVTG_TRY(check_arch());
  return launch_decode_nms(static_cast<cudaStream_t>(stream), *p, B, Lv, n_max, cls_logit,
                           conf_logit, coord, vid_len, duration, *out);
}

int32_t fvtg_temporal_nms(const float* windows, const int32_t* count, int32_t B, int32_t M,
                          double thd, int32_t mode, int32_t max_after_nms, float* out_windows,
                          int32_t* order, int32_t* out_count, void* stream) {
  host_state().launches = 0;
  if (!windows) return fail(FVTG_EINVAL, "temporal_nms: null windows");
  FVTG_TRY(check_arch());
  return launch_temporal_nms(static_cast<cudaStream_t>(stream), windows, count, B, M, thd, mode,
                             max_after_nms, out_windows, order, out_count);
}

int32_t fvtg_temporal_nms_hull_f64(const double* windows, const int32_t* count, int32_t B, int32_t M,
                                   double thd, int32_t max_after_nms, int32_t* order,
                                   int32_t* out_count, void* stream) {
  host_state().launches = 0;
  if (!windows || !order) return fail(FVTG_EINVAL, "temporal_nms_hull_f64: null argument");
  FVTG_TRY(check_arch());
  return launch_temporal_nms_hull_f64(static_cast<cudaStream_t>(stream), windows, count, B, M, thd,
                                      max_after_nms, order, out_count);
}

int32_t fvtg_forward(const FvtgCfg* cfg, const FvtgWeights* w, const FvtgBatch* in,
                     const float* duration, const FvtgDecodeParams* dp, const FvtgFusionOut* fout,
                     const FvtgHeadsOut* hout, const FvtgDecodeOut* dout, void* workspace,
                     size_t ws_bytes, void* stream) {
  host_state().launches = 0;
  FVTG_TRY(check_cfg(cfg));
  if (!w || !in || !fout || !dout || !in->vid || !in->txt || !in->vid_len || !in->txt_len ||
      !fout->saliency)
    return fail(FVTG_EINVAL, "forward: null argument");
  FVTG_TRY(check_shapes(cfg, in->B, in->Lv, in->Lt));
  FVTG_TRY(check_arch());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int bc = 0;
  Ws ws;
  FVTG_TRY(prep_ws(cfg, in->B, in->Lv, in->Lt, workspace, ws_bytes, &bc, &ws));
  const int Lv = in->Lv, Lt = in->Lt, nd = cfg->num_dummies;
  PyrGeo g0 = make_geo(*cfg, Lv, nullptr);
  if (hout && hout->cls_logit && hout->n_max != g0.n_max)
    return fail(FVTG_EINVAL, "forward: heads n_max %d != %d", hout->n_max, g0.n_max);
  FvtgDecodeParams p = dp ? *dp : default_decode(*cfg, *w);
  const bool keep_heads = hout && hout->cls_logit && hout->conf_logit && hout->coord;
  for (int b0 = 0; b0 < in->B; b0 += bc) {
    const int nb = in->B - b0 < bc ? in->B - b0 : bc;
    FVTG_TRY(fusion_chunk(
        st, *cfg, *w, ws, nb, Lv, Lt, in->vid + static_cast<size_t>(b0) * Lv * cfg->v_dim,
        in->txt + static_cast<size_t>(b0) * Lt * cfg->t_dim, in->vid_len + b0, in->txt_len + b0,
        fout->video_emb ? fout->video_emb + static_cast<size_t>(b0) * Lv * 256 : nullptr,
        fout->saliency + static_cast<size_t>(b0) * Lv,
        fout->t2v ? fout->t2v + static_cast<size_t>(b0) * Lv : nullptr,
        fout->dummy_tokens ? fout->dummy_tokens + static_cast<size_t>(b0) * nd * 256 : nullptr,
        in->uniform_vid_len != 0));
    float* cls = keep_heads ? hout->cls_logit + static_cast<size_t>(b0) * g0.n_max : ws.cls;
    float* conf = keep_heads ? hout->conf_logit + static_cast<size_t>(b0) * g0.n_max : ws.conf;
    float* coord = keep_heads ? hout->coord + static_cast<size_t>(b0) * g0.n_max * 2 : ws.coord;
    FVTG_TRY(pyramid_heads_chunk(st, *cfg, *w, ws, nb, Lv, ws.Yf, true, in->vid_len + b0, cls, conf, coord));
    FvtgDecodeOut d = *dout;
    const int tk = p.topk;
    if (d.boundary) d.boundary += static_cast<size_t>(b0) * tk * 3;
    if (d.windows) d.windows += static_cast<size_t>(b0) * tk * 3;
    if (d.nms_windows) d.nms_windows += static_cast<size_t>(b0) * tk * 3;
    if (d.nms_order) d.nms_order += static_cast<size_t>(b0) * tk;
    if (d.count) d.count += b0;
    if (d.nms_count) d.nms_count += b0;
    FVTG_TRY(launch_decode_nms(st, p, nb, Lv, g0.n_max, cls, conf, coord, in->vid_len + b0,
                               duration ? duration + b0 : nullptr, d));
  }
  return FVTG_OK;
}

}  // extern "C"
