/*
 * flashvtg_b200 — C-ABI of the B200 (sm_100a) FlashVTG inference hot path.
 *
 * Drop-in boundary.  The reference has no FFI: its "operator
 * interface" for this path is three Python call sites, each of which one entry
 * point below replaces:
 *
 *   fvtg_fusion_fwd          <- FlashVTG.forward, FlashVTG/model.py:148-184 (input projections,
 *                               dummy-token encoder, T2V cross-attention stack, self-attention
 *                               encoder, saliency head; transformer.py:82-115, crossattention.py:279-398)
 *   fvtg_pyramid_heads_fwd   <- FlashVTG/model.py:186-208 (ConvPyramid blocks/blocks.py:52-70,
 *                               ConfidenceScorer model.py:60-71 x2, ConvHead blocks/blocks.py:90-105)
 *   fvtg_decode_nms          <- FlashVTG/model.py:201,247-266 (ASR mix, sigmoid, span decode, top-k),
 *                               FlashVTG/inference.py:286-290 (clamp + 4-dp), postprocessing.py:25-50,
 *                               FlashVTG/inference.py:36-57 (post_processing_mr_nms)
 *   fvtg_temporal_nms        <- FlashVTG/inference.py:36-57 (modes normal / linear) and
 *                               utils/temporal_nms.py:25-74 (mode hull) on caller-supplied windows
 *
 * Conventions: plain pointers and sizes only; every pointer is DEVICE memory
 * owned by the caller (PyTorch's allocator in the shipped host code) unless the
 * field says "host".  The library allocates nothing, never synchronises, and
 * launches on the stream it is given.  Return 0 on success, a negative
 * FVTG_E* code otherwise; fvtg_last_error() returns a thread-local message.
 * There is no CPU fallback: on a device that is not sm_100 every compute entry
 * point returns FVTG_EARCH.
 */
#ifndef FLASHVTG_B200_H
#define FLASHVTG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVTG_ABI_VERSION 2

#define FVTG_OK 0
#define FVTG_EINVAL (-1)   /* bad shape / config / null pointer */
#define FVTG_EARCH (-2)    /* device is not sm_100 */
#define FVTG_ELAUNCH (-3)  /* CUDA launch or driver error (text in fvtg_last_error) */
#define FVTG_EWORKSPACE (-4) /* workspace too small */

#define FVTG_HIDDEN 256
#define FVTG_HEADS 8
#define FVTG_FFN 1024
#define FVTG_MLP_HIDDEN 128
#define FVTG_MAX_LAYERS 8
#define FVTG_MAX_LEVELS 8
#define FVTG_MAX_CONVS 4
#define FVTG_MAX_MLP 8
#define FVTG_MAX_TOPK 64

/* Model hyper-parameters (reference: FlashVTG/config.py:96-131, data/MR*.py:3-8). */
typedef struct FvtgCfg {
  int32_t abi_version;     /* FVTG_ABI_VERSION */
  int32_t v_dim;           /* raw video feature dim incl. TEF (770 / 2818 / 4098) */
  int32_t t_dim;           /* raw text feature dim (4096 / 512 / 300) */
  int32_t v_dim_pad;       /* v_dim rounded up to a multiple of 64 (bf16 staging pitch) */
  int32_t t_dim_pad;
  int32_t num_dummies;     /* --num_dummies */
  int32_t dummy_layers;    /* --dummy_layers */
  int32_t t2v_layers;      /* --t2v_layers */
  int32_t enc_layers;      /* --enc_layers */
  int32_t num_levels;      /* len(strides); strides are 1,2,4,... */
  int32_t head_k;          /* --kernel_size (odd) */
  int32_t num_conv_layers; /* --num_conv_layers */
  int32_t num_mlp_layers;  /* --num_mlp_layers (>= 2) */
  int32_t coord_k;         /* ConvHead kernal_size (3) */
  int32_t max_num_moment;  /* top-k (50) */
  float clip_len;          /* --clip_length */
} FvtgCfg;

typedef struct FvtgLN {
  const float* g;
  const float* b;
} FvtgLN;

/* Linear / conv weights are bf16, row-major [n_out][k_pad] (K-major, k_pad a multiple of 64, zero
 * padded); conv taps are folded into K as [n_out][tap][c_in].  Biases / LayerNorm params are fp32. */
typedef struct FvtgLinear {
  const void* w;
  const float* b;
} FvtgLinear;

typedef struct FvtgInProj {  /* LinearLayer x2, model.py:767-789 */
  FvtgLN ln0;       /* over the raw feature dim (kept for reference; folded into fc0, see below) */
  FvtgLinear fc0;   /* LayerNorm-folded first layer: w = bf16(W * diag(gamma_ln0)) [256][dim_pad],
                       b = W . beta_ln0 + bias (fp32 [256]) */
  FvtgLN ln1;       /* over 256 */
  FvtgLinear fc1;   /* [256][256]; bias has token_type_embeddings row folded in */
  const float* fc0_wsum;  /* fp32 [256]: row sums of fc0.w as stored (bf16 values), for the mean correction */
} FvtgInProj;

typedef struct FvtgEncLayer {  /* transformer.py:387-421 / :311-369 */
  FvtgLinear in_proj;   /* [768][256]; null for T2V layers (crossattention has no in-proj) */
  FvtgLinear out_proj;  /* [256][256] */
  FvtgLN norm1;
  FvtgLinear ff1;       /* [1024][256] */
  FvtgLinear ff2;       /* [256][1024] */
  FvtgLN norm2;
  float prelu;          /* activation.weight (scalar) */
  int32_t _pad;
} FvtgEncLayer;

typedef struct FvtgPyrConv {  /* blocks.py:37-45: Conv1d(256,256,2,stride 2) + LayerNorm + ReLU */
  FvtgLinear conv;  /* [256][2*256] */
  FvtgLN ln;
} FvtgPyrConv;

typedef struct FvtgScoreHead {  /* ConfidenceScorer, model.py:44-71 */
  FvtgLinear conv[FVTG_MAX_CONVS];  /* [256][head_k*256] */
  FvtgLinear mlp[FVTG_MAX_MLP];     /* layers 0..num_mlp_layers-2: [128][256] then [128][128] */
  const float* last_w;              /* final Linear(128,1) weight, fp32 [128] */
  float last_b;
  int32_t _pad;
} FvtgScoreHead;

typedef struct FvtgWeights {
  FvtgInProj vid, txt;
  const float* dummy_tok;  /* dummy_rep_token fp32 [nd][256] */
  const float* dummy_pos;  /* dummy_rep_pos   fp32 [nd][256] */
  FvtgEncLayer dummy[FVTG_MAX_LAYERS];  /* txtproj_encoder.layers */
  FvtgEncLayer t2v[FVTG_MAX_LAYERS];    /* transformer.t2v_encoder.layers */
  FvtgEncLayer enc[FVTG_MAX_LAYERS];    /* transformer.encoder.layers */
  const float* sal_w1;   /* saliency_proj1.weight fp32 [256 out][256 in] */
  const float* sal_b1;
  const float* sal_w2t;  /* saliency_proj2.weight TRANSPOSED fp32 [256 in][256 out] */
  const float* sal_b2;
  FvtgPyrConv pyr[FVTG_MAX_LEVELS][FVTG_MAX_LEVELS];  /* [level l][step j], j < l */
  FvtgScoreHead cls, conf;
  FvtgLinear coord1;  /* [256][coord_k*256] */
  FvtgLinear coord2;  /* [16][coord_k*256], rows 2..15 zero; bias fp32 [16] */
  float coef[FVTG_MAX_LEVELS];  /* host values of self.coef */
  float x;                      /* host value of self.x (ASR mix) */
  int32_t _pad;
} FvtgWeights;

/* One batch of B videos / queries.  Ragged semantics: video b is processed with its own
 * true lengths vid_len[b] <= Lv, txt_len[b] <= Lt exactly as the reference would at bs=1. */
typedef struct FvtgBatch {
  int32_t B, Lv, Lt;
  int32_t uniform_vid_len; /* hint: 1 = the caller guarantees vid_len[b] == vid_len[0] for all b (the sine
                              position table is then built once for Lv rows instead of per video); 0 = may differ */
  const float* vid;       /* fp32 [B][Lv][v_dim]  (src_vid, TEF appended) */
  const float* txt;       /* fp32 [B][Lt][t_dim]  (src_txt) */
  const int32_t* vid_len; /* [B] */
  const int32_t* txt_len; /* [B] */
} FvtgBatch;

typedef struct FvtgFusionOut {
  float* video_emb;     /* fp32 [B][Lv][256]: encoder output F (before the pyramid's in-place ReLU) */
  float* saliency;      /* fp32 [B][Lv] */
  float* t2v;           /* fp32 [B][Lv]  (t2vattnvalues) */
  float* dummy_tokens;  /* fp32 [B][nd][256] */
} FvtgFusionOut;

typedef struct FvtgHeadsOut {
  int32_t n_max, _pad;  /* points per video at full length: sum_l floor(Lv / 2^l) */
  float* cls_logit;     /* fp32 [B][n_max]  class_head, levels concatenated */
  float* conf_logit;    /* fp32 [B][n_max]  conf_head */
  float* coord;         /* fp32 [B][n_max][2]  exp(coord_head) * coef[level] */
} FvtgHeadsOut;

#define FVTG_NMS_NONE (-1)
#define FVTG_NMS_NORMAL 0  /* inference.py:45-46: score <- 0 where iou >= thd */
#define FVTG_NMS_LINEAR 1  /* inference.py:47-48: score *= 1 - iou */
#define FVTG_NMS_HULL 2    /* utils/temporal_nms.py: hull "union", strict >, removal */

typedef struct FvtgDecodeParams {
  double nms_thd;     /* compared as fp32 in normal / linear mode (torch), as fp64 in hull mode (python) */
  float x;            /* ASR mix weight */
  float clip_len;     /* (float)clip_length */
  float inv_clip_len; /* (float)(1.0 / clip_length), the divisor of model.py:260, rounded once from fp64 */
  float min_ts, max_ts;
  int32_t topk;       /* max_num_moment */
  int32_t num_levels;
  int32_t clip_ts;    /* 1: apply clip_min_max_timestamps (postprocessing.py:38-43) */
  int32_t round_multiple; /* 1: round_to_multiple_clip_lengths (postprocessing.py:45-50) */
  int32_t nms_mode;   /* FVTG_NMS_* */
  int32_t max_after_nms; /* hull mode only */
  int32_t _pad;
} FvtgDecodeParams;

typedef struct FvtgDecodeOut {
  float* boundary;     /* fp32 [B][topk][3]: raw ranked (st, ed, score) == outputs["_out"]["boundary"] */
  float* windows;      /* fp32 [B][topk][3]: after clamp(0,duration), 4-dp, post-processor */
  float* nms_windows;  /* fp32 [B][topk][3]: `windows` after NMS, final order (may be null if NMS_NONE) */
  int32_t* nms_order;  /* [B][topk]: index into `windows` of each nms_windows row */
  int32_t* count;      /* [B]: valid rows in boundary/windows: min(N_b, topk) */
  int32_t* nms_count;  /* [B]: rows in nms_windows (== count except hull mode) */
} FvtgDecodeOut;

/* ---- checkpoint loading --------------------------------------------------------------------------
 * The weight ABI of the drop-in is the reference's state_dict (FlashVTG/inference.py:471,
 * `model.load_state_dict(checkpoint["model"], strict=True)`): one FvtgParam per entry, HOST fp32,
 * contiguous, under the reference's key name (model.py:81-135, transformer.py:311-330,387-405,
 * blocks/blocks.py:23-50,93-101).  fvtg_pack_weights converts them into the packed device layout documented
 * on FvtgLinear / FvtgInProj above (bf16 K-major operands, conv taps folded into K, LayerNorm and token-type
 * folding) inside ONE caller-owned device buffer of fvtg_packed_weights_bytes(cfg) bytes (256-byte aligned)
 * and fills `out` with pointers into it; the copy is enqueued on `stream`.  Strict like torch: a missing key,
 * an unexpected key or a wrong element count returns FVTG_EINVAL with the key in fvtg_last_error(). */
typedef struct FvtgParam {
  const char* name;    /* reference state_dict key */
  const float* data;   /* host fp32, contiguous in the reference's own shape */
  int64_t numel;
} FvtgParam;
size_t fvtg_packed_weights_bytes(const FvtgCfg* cfg);
int32_t fvtg_pack_weights(const FvtgCfg* cfg, const FvtgParam* params, int32_t n_params, void* device_buf,
                          size_t device_bytes, FvtgWeights* out, void* stream);
/* The same packing into a HOST buffer, no CUDA call: `out` points into the address range starting at
 * target_base (null = host_buf itself), i.e. where the caller will place the bytes (its own device upload,
 * a memory-mapped weight file, ...). */
int32_t fvtg_pack_weights_host(const FvtgCfg* cfg, const FvtgParam* params, int32_t n_params, void* host_buf,
                               size_t host_bytes, const void* target_base, FvtgWeights* out);

size_t fvtg_workspace_bytes(const FvtgCfg* cfg, int32_t B, int32_t Lv, int32_t Lt);
/* Videos processed per internal chunk (sized so a chunk's activations stay L2 resident). */
int32_t fvtg_chunk_videos(const FvtgCfg* cfg, int32_t Lv, int32_t Lt);

int32_t fvtg_fusion_fwd(const FvtgCfg* cfg, const FvtgWeights* w, const FvtgBatch* in,
                        const FvtgFusionOut* out, void* workspace, size_t ws_bytes, void* stream);

int32_t fvtg_pyramid_heads_fwd(const FvtgCfg* cfg, const FvtgWeights* w, int32_t B, int32_t Lv,
                               const float* video_emb, const int32_t* vid_len,
                               const FvtgHeadsOut* out, void* workspace, size_t ws_bytes,
                               void* stream);

int32_t fvtg_decode_nms(const FvtgDecodeParams* p, int32_t B, int32_t Lv, int32_t n_max,
                        const float* cls_logit, const float* conf_logit, const float* coord,
                        const int32_t* vid_len, const float* duration, const FvtgDecodeOut* out,
                        void* stream);

/* Standalone temporal NMS on caller windows fp32 [B][M][3] (st, ed, score), count[b] <= M <= 64.
 * Outputs: out_windows [B][M][3] in final order, order [B][M] (index of the source row),
 * out_count [B].  Bit-exact with the reference arithmetic (fp32 for normal/linear, fp64 for hull). */
int32_t fvtg_temporal_nms(const float* windows, const int32_t* count, int32_t B, int32_t M,
                          double thd, int32_t mode, int32_t max_after_nms, float* out_windows,
                          int32_t* order, int32_t* out_count, void* stream);

/* utils/temporal_nms.py:25-74 on fp64 rows (the reference works on python floats there):
 * windows fp64 [B][M][3]; order [B][M] = kept source rows in output order (-1 beyond out_count). */
int32_t fvtg_temporal_nms_hull_f64(const double* windows, const int32_t* count, int32_t B, int32_t M,
                                   double thd, int32_t max_after_nms, int32_t* order,
                                   int32_t* out_count, void* stream);

/* Whole path (fusion -> pyramid+heads -> decode/NMS), chunk by chunk so each chunk's intermediates
 * stay in L2. `duration` fp32 [B]. heads may be null (logits then live only in the workspace). */
int32_t fvtg_forward(const FvtgCfg* cfg, const FvtgWeights* w, const FvtgBatch* in,
                     const float* duration, const FvtgDecodeParams* dp, const FvtgFusionOut* fout,
                     const FvtgHeadsOut* hout, const FvtgDecodeOut* dout, void* workspace,
                     size_t ws_bytes, void* stream);

/* ---- Device-resident input pipeline (what the reference's loader does per item on the host) ----
 * Replaces, for raw feature rows already on the device: l2_normalize_np_array per feature directory
 * (FlashVTG/start_end_dataset.py:524-530, utils/basic_utils.py:84-86), concatenation, the temporal
 * endpoint features (start_end_dataset.py:174-180), zero padding and the 0/1 masks of start_end_collate
 * (utils/tensor_utils.py:5-53).  Raw rows may be fp32 / fp16 / bf16 (the loader's astype(np.float32)). */
#define FVTG_RAW_MAX_GROUPS 4
#define FVTG_RAW_F32 0
#define FVTG_RAW_F16 1
#define FVTG_RAW_BF16 2
typedef struct FvtgRawBatch {
  int32_t B, Lv, Lt;
  int32_t n_groups;                         /* feature directories of the video (--v_feat_dirs) */
  int32_t group_dim[FVTG_RAW_MAX_GROUPS];   /* their feature dims, concatenated in this order */
  int32_t t_dim;                            /* raw text feature dim */
  int32_t dtype;                            /* FVTG_RAW_* of every raw array */
  int32_t normalize_v, normalize_t;         /* not --no_norm_vfeat / --no_norm_tfeat */
  int32_t use_tef;                          /* --ctx_mode contains "tef": append [i/L, (i+1)/L] */
  int32_t _pad;
  const void* vid[FVTG_RAW_MAX_GROUPS];     /* [B][Lv][group_dim[g]], rows >= vid_len[b] ignored */
  const void* txt;                          /* [B][Lt][t_dim] */
  const int32_t* vid_len;                   /* [B] */
  const int32_t* txt_len;                   /* [B] */
} FvtgRawBatch;
/* src_vid fp32 [B][Lv][sum(group_dim) + 2*use_tef], src_txt fp32 [B][Lt][t_dim]; masks fp32 [B][L] or null. */
int32_t fvtg_prepare_inputs(const FvtgRawBatch* raw, float* src_vid, float* src_vid_mask,
                            float* src_txt, float* src_txt_mask, void* stream);

/* ---- QVHighlights evaluation on the device (what standalone_eval does per query on the host) ----
 * Replaces the per-query loops of standalone_eval/eval.py: compute_mr_ap (:24-69, detection AP of
 * utils.py:83-159 at IoU 0.5..0.95), compute_mr_r1 (:72-102), the length ranges of
 * eval_moment_retrieval (:109-170: 0 short (0,10], 1 middle (10,30], 2 long (30,150], 3 full),
 * compute_hl_hit1 / compute_hl_ap (:173-236, get_ap of utils.py:162-209 for minimum scores 2,3,4 and the
 * 3 annotators).  All arithmetic in fp64, in numpy's operation order (pairwise sums included), so the
 * per-query results are bit-identical; the means over queries and the 2-decimal formatting stay on the
 * host (flashvtg_b200/evaluation.py).  One row per query; every pointer is device memory. */
typedef struct FvtgEvalBatch {
  int32_t n_queries;
  int32_t max_pred;              /* row stride of pred_win (windows per query) */
  int32_t max_gt;                /* row stride of gt_win, <= 32 */
  int32_t max_sal;               /* row stride of pred_sal */
  int32_t max_clips;             /* row stride of gt_sal */
  int32_t _pad;
  const double* pred_win;        /* [Q][max_pred][3] (start, end, score) in submission order */
  const int32_t* pred_cnt;       /* [Q] */
  const double* gt_win;          /* [Q][max_gt][2] relevant_windows */
  const int32_t* gt_cnt;         /* [Q] */
  const double* pred_sal;        /* [Q][max_sal] pred_saliency_scores (null: no highlight metrics) */
  const int32_t* pred_sal_len;   /* [Q] */
  const uint8_t* gt_sal;         /* [Q][max_clips][3] annotator scores 0..4 of every clip (mk_gt_scores) */
  const int32_t* gt_clips;       /* [Q] int(duration / clip_length) */
} FvtgEvalBatch;
/* mr_ap fp64 [4][Q][10], mr_iou fp64 [4][Q] (top-1 IoU with the best GT window), mr_valid u8 [4][Q]
 * (query has a GT window in the range); hl_ap fp64 [3][Q][3] (min score, query, annotator), hl_hit u8
 * [3][Q].  Pass mr_ap == null or hl_ap == null to skip a family.  max_pred_windows: the reference's 10. */
int32_t fvtg_eval_submission(const FvtgEvalBatch* batch, int32_t max_pred_windows, double* mr_ap,
                             double* mr_iou, uint8_t* mr_valid, double* hl_ap, uint8_t* hl_hit,
                             void* stream);

/* Number of kernel launches the last fvtg_* compute call on this thread issued. */
int64_t fvtg_last_launch_count(void);
const char* fvtg_last_error(void);
int32_t fvtg_abi_version(void);

/* Bench hook.  fvtg_prof_enable(1): every kernel launched by this thread's later fvtg_* calls is
 * bracketed by a CUDA event pair on its stream.  fvtg_prof_collect waits for them and returns, per
 * kernel class (0 tcgen05 GEMM, 1 attention, 2 fused first input projection, 3 decode/NMS, 4 other,
 * 5 fused tcgen05 transformer-layer kernel),
 * the summed device time in ms and the launch count since the last collect. */
#define FVTG_PROF_CLASSES 6
void fvtg_prof_enable(int32_t on);
int32_t fvtg_prof_collect(double* ms, int64_t* launches, int32_t n_classes);

#ifdef __cplusplus
}
#endif
#endif /* FLASHVTG_B200_H */
