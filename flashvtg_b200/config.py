"""Model hyper-parameters of the FlashVTG hot path and the per-dataset presets.

The reference spreads these over argparse flags (FlashVTG/config.py:96-131,163-168), the nncore
config files (data/MR.py:3-8, data/MR_16.py, data/MR_32.py) and the launch scripts
(FlashVTG/scripts/*/train.sh); SURVEY.md §5 tabulates them.  `ModelConfig.from_opt` accepts the
reference's parsed options object so existing drivers can construct the B200 model unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Tuple


@dataclass(frozen=True)
class ModelConfig:
    v_feat_dim: int                      # --v_feat_dim (+2 when TEF is on, config.py:241-242)
    t_feat_dim: int                      # --t_feat_dim
    num_dummies: int = 40                # --num_dummies
    dummy_layers: int = 2                # --dummy_layers
    t2v_layers: int = 6                  # --t2v_layers
    enc_layers: int = 3                  # --enc_layers
    strides: Tuple[int, ...] = (1, 2, 4, 8)   # cfg.model.strides (data/MR.py:4)
    kernel_size: int = 5                 # --kernel_size (score heads)
    num_conv_layers: int = 1             # --num_conv_layers
    num_mlp_layers: int = 5              # --num_mlp_layers
    coord_kernel: int = 3                # ConvHead kernal_size (data/MR.py:10)
    clip_length: float = 2.0             # --clip_length
    max_num_moment: int = 50             # cfg.model.max_num_moment
    buffer_size: int = 1024              # cfg.model.buffer_size (generator.py:60 assert)
    hidden_dim: int = 256
    nheads: int = 8
    dim_feedforward: int = 1024
    n_input_proj: int = 2
    max_q_l: int = 40                    # only sizes the unused txt_position_embed table
    dset_name: str = "hl"
    nms_thd: float = 0.7                 # --nms_thd
    nms_type: str = "normal"             # --nms_type
    name: str = field(default="custom", compare=False)

    def __post_init__(self):
        if (self.hidden_dim, self.nheads, self.dim_feedforward) != (256, 8, 1024):
            raise ValueError("the B200 kernels are specialised for hidden 256 / 8 heads / ffn 1024 "
                             "(constant in every reference script, SURVEY §5)")
        if self.n_input_proj != 2:
            raise ValueError("n_input_proj must be 2 (every reference script)")
        for i, s in enumerate(self.strides):
            if s != 1 << i:
                raise ValueError("strides must be (1, 2, 4, ...)")
        if not (1 <= len(self.strides) <= 8):
            raise ValueError("1..8 pyramid levels supported")
        if self.kernel_size % 2 != 1 or self.kernel_size > 7 or self.coord_kernel != 3:
            raise ValueError("head kernel_size must be odd and <= 7; coord kernel must be 3")
        if not (1 <= self.num_conv_layers <= 4) or not (2 <= self.num_mlp_layers <= 8):
            raise ValueError("num_conv_layers in 1..4, num_mlp_layers in 2..8")
        if max(self.dummy_layers, self.t2v_layers, self.enc_layers) > 8:
            raise ValueError("at most 8 layers per stack")

    @property
    def num_levels(self) -> int:
        return len(self.strides)

    def level_lengths(self, lv: int):
        """L_l = floor(lv / 2^l), levels with lv < 2^l skipped (blocks.py:56)."""
        return [lv >> l for l in range(self.num_levels) if lv >= (1 << l)]

    def num_points(self, lv: int) -> int:
        return sum(self.level_lengths(lv))

    @classmethod
    def from_opt(cls, opt) -> "ModelConfig":
        """From the reference's parsed options (BaseOptions.parse(), with opt.cfg = nncore.Config)."""
        m = opt.cfg.model
        return cls(
            v_feat_dim=opt.v_feat_dim, t_feat_dim=opt.t_feat_dim, num_dummies=opt.num_dummies,
            dummy_layers=opt.dummy_layers, t2v_layers=opt.t2v_layers, enc_layers=opt.enc_layers,
            strides=tuple(m.strides), kernel_size=opt.kernel_size,
            num_conv_layers=opt.num_conv_layers, num_mlp_layers=opt.num_mlp_layers,
            coord_kernel=(m.coord_head_cfg.get("kernal_size", 3) if m.coord_head_cfg else 3),
            clip_length=opt.clip_length, max_num_moment=m.max_num_moment,
            buffer_size=m.buffer_size, hidden_dim=opt.hidden_dim, nheads=opt.nheads,
            dim_feedforward=opt.dim_feedforward, n_input_proj=opt.n_input_proj,
            max_q_l=(100 if opt.max_q_l == -1 else opt.max_q_l), dset_name=opt.dset_name,
            nms_thd=getattr(opt, "nms_thd", 0.7), nms_type=getattr(opt, "nms_type", "normal"))

    def with_(self, **kw) -> "ModelConfig":
        return replace(self, **kw)


# SURVEY §5 "Model-shape table from the scripts" + BASELINE.json configs.
PRESETS = {
    # scripts/qv_internvideo2/train.sh:15-57 + data/MR_16.py  (BASELINE config #2, the bench workload)
    "qvh_iv2": ModelConfig(v_feat_dim=770, t_feat_dim=4096, num_dummies=40, strides=(1, 2, 4, 8, 16),
                           kernel_size=5, num_conv_layers=1, num_mlp_layers=5, clip_length=2.0,
                           max_q_l=40, dset_name="hl", name="qvh_iv2"),
    # upstream SlowFast+CLIP dims (BASELINE config #1): 2304+512+2 TEF, CLIP text 512, nd=10
    "qvh_sfclip": ModelConfig(v_feat_dim=2818, t_feat_dim=512, num_dummies=10, strides=(1, 2, 4, 8),
                              kernel_size=5, num_conv_layers=1, num_mlp_layers=5, clip_length=2.0,
                              max_q_l=32, dset_name="hl", name="qvh_sfclip"),
    # scripts/charades_sta/train_vgg.sh:17-60,95  (BASELINE config #3)
    "charades_vgg": ModelConfig(v_feat_dim=4098, t_feat_dim=300, num_dummies=40,
                                strides=(1, 2, 4, 8), kernel_size=3, num_conv_layers=2,
                                num_mlp_layers=5, clip_length=0.166666, max_q_l=100,
                                dset_name="charadesSTA", name="charades_vgg"),
    # scripts/charades_sta/train.sh:17-64,99
    "charades_sfclip": ModelConfig(v_feat_dim=2818, t_feat_dim=512, num_dummies=40,
                                   strides=(1, 2, 4, 8), kernel_size=5, num_conv_layers=1,
                                   num_mlp_layers=5, clip_length=1.0, max_q_l=32,
                                   dset_name="charadesSTA", name="charades_sfclip"),
    # scripts/charades_sta_internvideo2/train.sh:15-45,80
    "charades_iv2": ModelConfig(v_feat_dim=770, t_feat_dim=4096, num_dummies=40,
                                strides=(1, 2, 4, 8), kernel_size=7, num_conv_layers=2,
                                num_mlp_layers=3, clip_length=1.0, max_q_l=23,
                                dset_name="charadesSTA", name="charades_iv2"),
    # scripts/tacos/train.sh:18-63,98  (BASELINE config #4)
    "tacos": ModelConfig(v_feat_dim=2818, t_feat_dim=512, num_dummies=35, dummy_layers=3,
                         t2v_layers=8, strides=(1, 2, 4, 8), kernel_size=5, num_conv_layers=2,
                         num_mlp_layers=5, clip_length=2.0, max_q_l=100, dset_name="tacos",
                         name="tacos"),
    # same with data/MR_32.py: the "deep pyramid" stress of BASELINE config #4
    "tacos_deep": ModelConfig(v_feat_dim=2818, t_feat_dim=512, num_dummies=35, dummy_layers=3,
                              t2v_layers=8, strides=(1, 2, 4, 8, 16, 32), kernel_size=5,
                              num_conv_layers=2, num_mlp_layers=5, clip_length=2.0, max_q_l=100,
                              dset_name="tacos", name="tacos_deep"),
    # highlight detection (data/HD.py: a single pyramid level, buffer 2048)
    # scripts/tvsum/train.sh:23-56: SlowFast+CLIP 2816+2 TEF, CLIP text, 3 dummies, t2v 2, k5/c2/m3
    "tvsum": ModelConfig(v_feat_dim=2818, t_feat_dim=512, num_dummies=3, dummy_layers=2, t2v_layers=2,
                         enc_layers=3, strides=(1,), kernel_size=5, num_conv_layers=2, num_mlp_layers=3,
                         clip_length=2.0, buffer_size=2048, max_q_l=100, dset_name="tvsum", name="tvsum"),
    # scripts/youtube_uni/train.sh:20-84: same features, 1 dummy, clip_length 1
    "youtube_uni": ModelConfig(v_feat_dim=2818, t_feat_dim=512, num_dummies=1, dummy_layers=2, t2v_layers=2,
                               enc_layers=3, strides=(1,), kernel_size=5, num_conv_layers=2,
                               num_mlp_layers=3, clip_length=1.0, buffer_size=2048, max_q_l=100,
                               dset_name="youtube_uni", name="youtube_uni"),
}


def postprocessor_preset(cfg: ModelConfig):
    """PostProcessorDETR arguments per dataset (FlashVTG/inference.py:312-352):
    returns (clip_ts: bool, min_ts, max_ts, round_multiple: bool)."""
    if cfg.dset_name == "hl":
        return True, 0.0, 150.0, True
    if cfg.dset_name == "charadesSTA":
        # Quirk kept: the reference tests `opt.v_feat_dim == 4096` AFTER TEF added 2
        # (config.py:241-242), so the 360 s branch only triggers for VGG features without TEF.
        if cfg.v_feat_dim == 4096:
            return True, 0.0, 360.0, True
        return True, 0.0, 150.0, True
    return False, 0.0, 50000.0, True


def gemm_flops_per_video(cfg: ModelConfig, lv: int, lt: int) -> float:
    """Algorithmic FLOPs (2 x MAC, unpadded dims) of the dense d_model contractions of one video -
    every nn.Linear / Conv of the path, i.e. what
    the tcgen05 GEMM kernel computes.  Excludes the per-head attention bmm's (attention kernel)
    and the saliency head (mat-vec kernel).  QVH-IV2 (75, 32): 1.572e9 (of the 1.634e9 total FlopCounterMode reports)."""
    d, ff = 256, 1024
    s = cfg.num_dummies + lt
    mac = lv * (cfg.v_feat_dim * d + d * d) + lt * (cfg.t_feat_dim * d + d * d)
    sa = 3 * d * d + d * d + 2 * d * ff
    mac += cfg.dummy_layers * s * sa
    mac += cfg.t2v_layers * lv * (d * d + 2 * d * ff)
    mac += cfg.enc_layers * lv * sa
    n = 0
    for l in range(cfg.num_levels):
        if lv < (1 << l):
            continue
        n += lv >> l
        for j in range(1, l + 1):
            mac += (lv >> j) * 2 * d * d
    k = cfg.kernel_size
    head = cfg.num_conv_layers * k * d * d + d * 128 + (cfg.num_mlp_layers - 2) * 128 * 128 + 128
    mac += 2 * n * head
    mac += n * (cfg.coord_kernel * d * d + cfg.coord_kernel * d * 2)
    return 2.0 * mac


def flops_by_kernel_class(cfg: ModelConfig, lv: int, lt: int) -> dict:
    """Algorithmic FLOPs (2 x MAC, unpadded dims) of one video split by the kernel class that
    computes them (bench.py roofline numerators; classes as in fvtg_prof_collect):
      layer     - fused layer-tail kernel: out_proj + FFN of every dummy / T2V / encoder layer
      inproj    - first input-projection layer with its LayerNorm folded in (reads the fp32 features once)
      gemm      - persistent tcgen05 GEMM: second projections, QKV in-projections, pyramid convs, heads
      attention - per-(video, head) QK^T and PV products (legacy mma.sync kernel)
    The saliency mat-vecs (9.9 MFLOP at QVH-IV2) are left out, as in gemm_flops_per_video."""
    d, ff, dh, H = 256, 1024, 32, 8
    s = cfg.num_dummies + lt
    tail = d * d + 2 * d * ff
    layer = (cfg.dummy_layers * s + (cfg.t2v_layers + cfg.enc_layers) * lv) * tail
    total = gemm_flops_per_video(cfg, lv, lt) / 2.0
    attn = cfg.dummy_layers * H * (2 * s * s * dh)          # QK^T + PV over all S keys
    attn += cfg.t2v_layers * H * (lv * s * dh + lv * lt * dh)  # scores over S keys, values over text only
    attn += cfg.enc_layers * H * (2 * lv * lv * dh)
    inproj = (lv * cfg.v_feat_dim + lt * cfg.t_feat_dim) * d   # first projection layer (csrc/inproj.cu)
    return {"layer": 2.0 * layer, "gemm": 2.0 * (total - layer - inproj), "attention": 2.0 * attn,
            "inproj": 2.0 * inproj}
