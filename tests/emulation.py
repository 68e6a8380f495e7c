"""TEST INFRASTRUCTURE - the oracle's forward with bf16 rounding inserted at exactly the points
where the CUDA path stores bf16 (GEMM operands, attention probabilities, staged activations);
all accumulation, LayerNorm statistics, softmax, residual streams stay fp32 like the kernels.

Two uses: (1) a much sharper bug detector than the 1e-2 fp32 tolerance - the GPU must agree with
this emulation to ~1e-3; (2) it quantifies how much of the distance to the fp32 oracle is inherent
to bf16 operands (reported in DESIGN.md).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import forward as O

D, H, DH = 256, 8, 32


def q(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _lin(x, sd, prefix, extra_bias=None):
    b = sd[prefix + ".bias"]
    if extra_bias is not None:
        b = b + extra_bias
    return x @ q(sd[prefix + ".weight"]).t() + b


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _prelu(x, a):
    return torch.where(x > 0, x, a * x)


def input_proj(x, sd, name, te_row):
    """csrc/inproj.cu: LayerNorm over the raw dim folded into the first projection.  The kernel rounds the
    SHIFTED raw values (x - m0, m0 = mean of the row's first 64 values) and W * gamma to bf16, accumulates in
    fp32 and applies mean / rstd of the shifted row in the epilogue."""
    w0, b0 = sd[f"{name}.0.net.1.weight"], sd[f"{name}.0.net.1.bias"]
    g0, be0 = sd[f"{name}.0.LayerNorm.weight"], sd[f"{name}.0.LayerNorm.bias"]
    dim = x.shape[-1]
    m0 = x[:, : min(64, dim)].mean(-1, keepdim=True)
    xs = x - m0
    wg = q(w0 * g0[None, :])
    ms = xs.mean(-1, keepdim=True)
    var = ((xs * xs).mean(-1, keepdim=True) - ms * ms).clamp_min(0)
    rstd = torch.rsqrt(var + 1e-5)
    y = rstd * (q(xs) @ wg.t() - ms * wg.sum(1)[None, :]) + (w0 @ be0 + b0)[None, :]
    t = q(_ln(torch.relu(y), sd, f"{name}.1.LayerNorm"))
    return _lin(t, sd, f"{name}.1.net.1", te_row)


def _heads(x):
    return x.view(x.shape[0], H, DH).permute(1, 0, 2)


def _attn(qb, kb, vb, v_first=0):
    """Kernel arithmetic: S = q.k (bf16 in, fp32 acc) * scale, p = exp(s - max) rounded to bf16 for
    the PV product, normalised by the fp32 row sum afterwards."""
    s = (_heads(qb) @ _heads(kb).transpose(1, 2)) * (DH ** -0.5)
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    pv = q(p)[:, :, v_first:] @ _heads(vb)[:, v_first:]
    o = (pv / l).permute(1, 0, 2).reshape(qb.shape[0], D)
    tsum = (p[:, :, v_first:].sum(-1) / l[:, :, 0])          # (H, Lq)
    return q(o), tsum


def sa_layer(x, pos, sd, prefix):
    w = q(sd[prefix + ".self_attn.in_proj_weight"])
    b = sd[prefix + ".self_attn.in_proj_bias"]
    xb, xpb = q(x), q(x + pos)
    qq = q(xpb @ w[:D].t() + b[:D])
    kk = q(xpb @ w[D:2 * D].t() + b[D:2 * D])
    vv = q(xb @ w[2 * D:].t() + b[2 * D:])
    att, _ = _attn(qq, kk, vv)
    x1 = _ln(x + _lin(att, sd, prefix + ".self_attn.out_proj"), sd, prefix + ".norm1")
    h = q(_prelu(_lin(q(x1), sd, prefix + ".linear1"), sd[prefix + ".activation.weight"]))
    return _ln(x1 + _lin(h, sd, prefix + ".linear2"), sd, prefix + ".norm2")


def t2v_layer(y, pos_v, kc, nd, sd, prefix):
    att, tsum = _attn(q(y + pos_v), kc, kc, v_first=nd)
    z = y + _lin(att, sd, prefix + ".self_attn.out_proj")
    t = q(_ln(z, sd, prefix + ".norm1"))
    h = q(_prelu(_lin(t, sd, prefix + ".linear1"), sd[prefix + ".activation.weight"]))
    return _ln(z + _lin(h, sd, prefix + ".linear2"), sd, prefix + ".norm2"), tsum


def score_head(u, sd, head, cfg):
    k = cfg.kernel_size
    x = u.t().unsqueeze(0)
    for c in range(cfg.num_conv_layers):
        w = q(sd[f"{head}.convs.{c}.weight"])[:, :, 0, :]
        x = q(torch.relu(F.conv1d(x, w, sd[f"{head}.convs.{c}.bias"], padding=k // 2)))
    x = x[0].t()
    for m in range(cfg.num_mlp_layers - 1):
        x = torch.relu(_lin(x, sd, f"{head}.fc.layers.{m}"))
        if m < cfg.num_mlp_layers - 2:
            x = q(x)
    last = cfg.num_mlp_layers - 1
    return (x @ sd[f"{head}.fc.layers.{last}.weight"].t() + sd[f"{head}.fc.layers.{last}.bias"])[:, 0]


def coord_head(u, sd, cfg):
    p = cfg.coord_kernel // 2
    x = u.t().unsqueeze(0)
    x = q(torch.relu(F.conv1d(x, q(sd["coord_head.module.1.weight"]), sd["coord_head.module.1.bias"],
                              padding=p)))
    x = F.conv1d(x, q(sd["coord_head.module.3.weight"]), sd["coord_head.module.3.bias"], padding=p)
    return x[0].t()


def pyramid(f, sd, cfg):
    p0 = q(torch.relu(f))
    levels = [p0]
    for l in range(1, cfg.num_levels):
        if f.shape[0] < (1 << l):
            continue
        z = p0
        for j in range(l):
            w = q(sd[f"pyramid.blocks.{l}.{1 + 5 * j}.weight"])
            z = F.conv1d(z.t().unsqueeze(0), w, sd[f"pyramid.blocks.{l}.{1 + 5 * j}.bias"], stride=2)[0].t()
            z = q(torch.relu(_ln(z, sd, f"pyramid.blocks.{l}.{3 + 5 * j}")))
        levels.append(z)
    return levels


def forward_single(sd, cfg, vid, txt):
    nd = cfg.num_dummies
    te = sd["token_type_embeddings.weight"]
    v = input_proj(vid, sd, "input_vid_proj", te[1])
    t = input_proj(txt, sd, "input_txt_proj", te[0])
    lv = v.shape[0]
    pos_v = O.sine_pos(lv, torch.float32)
    dpos = sd["dummy_rep_pos"]
    x = torch.cat([sd["dummy_rep_token"], t], 0)
    p = torch.cat([dpos, torch.zeros_like(t)], 0)
    for i in range(cfg.dummy_layers):
        x = sa_layer(x, p, sd, f"txtproj_encoder.layers.{i}")
    dummy = x[:nd]
    kc = torch.cat([q(dummy + dpos), q(t)], 0)
    y = v
    tacc = torch.zeros(H, lv)
    for i in range(cfg.t2v_layers):
        y, ts = t2v_layer(y, pos_v, kc, nd, sd, f"transformer.t2v_encoder.layers.{i}")
        tacc = tacc + ts
    t2v = (tacc.sum(0) / (H * cfg.t2v_layers)).clamp(0, 1)
    for i in range(cfg.enc_layers):
        y = sa_layer(y, pos_v, sd, f"transformer.encoder.layers.{i}")
    f = y
    g = f.mean(0)
    u = sd["saliency_proj2.weight"] @ g + sd["saliency_proj2.bias"]
    sal = ((f @ sd["saliency_proj1.weight"].t() + sd["saliency_proj1.bias"]) * u[None]).sum(-1) / math.sqrt(D)
    levels = pyramid(f, sd, cfg)
    cls = torch.cat([score_head(u_, sd, "class_head", cfg) for u_ in levels])
    conf = score_head(torch.cat(levels, 0), sd, "conf_head", cfg)
    xm = sd["x"]
    logit = xm * cls + (1 - xm) * conf
    coord = torch.cat([coord_head(u_, sd, cfg).exp() * sd["coef"][i] for i, u_ in enumerate(levels)], 0)
    return dict(video_emb=f, saliency=sal, t2vattn=t2v, dummy_tokens=dummy, cls=cls, conf=conf,
                logit=logit, coord=coord)


def forward_batch(sd, cfg, batch):
    outs = []
    with torch.no_grad():
        for b in range(batch["src_vid"].shape[0]):
            lv, lt = int(batch["vid_len"][b]), int(batch["txt_len"][b])
            outs.append(forward_single(sd, cfg, batch["src_vid"][b, :lv].float(),
                                       batch["src_txt"][b, :lt].float()))
    return outs
