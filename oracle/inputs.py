"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference loader's per-item arithmetic between
the feature files and model.forward (the oracle of flashvtg_b200/inputs.py).

  l2_normalize          utils/basic_utils.py:84-86 (eps 1e-5, division by norm + eps)
  video features        FlashVTG/start_end_dataset.py:508-531 (astype(float32), normalise per feature
                        directory, concatenate)
  TEF                   FlashVTG/start_end_dataset.py:174-180 (torch fp32: arange/L, + 1.0/L)
  padding / masks       utils/tensor_utils.py:5-53 (zeros, mask 1 = valid)

Pinned against the unmodified `l2_normalize_np_array` and `pad_sequences_1d` when /root/reference is
present (tests/test_oracle_vs_reference.py) and against tests/golden/inputs_case.npz generated from them.
"""
from __future__ import annotations

import numpy as np
import torch


def l2_normalize(a, eps=1e-5):
    a = np.asarray(a)
    return a / (np.linalg.norm(a, axis=-1, keepdims=True) + eps)


def tef(L):
    st = torch.arange(0, L, 1.0) / L
    ed = st + 1.0 / L
    return torch.stack([st, ed], dim=1).numpy()


def prepare_item(groups, query, normalize_v=True, normalize_t=True, use_tef=True):
    """groups: list of (L, D_g) raw arrays of one video (any float dtype); query (Lq, Dt)."""
    feats = []
    for g in groups:
        f = np.asarray(g).astype(np.float32)
        feats.append(l2_normalize(f) if normalize_v else f)
    n = min(len(f) for f in feats)
    v = np.concatenate([f[:n] for f in feats], axis=1)
    if use_tef:
        v = np.concatenate([v, tef(n)], axis=1)
    q = np.asarray(query).astype(np.float32)
    if normalize_t:
        q = l2_normalize(q)
    return v.astype(np.float32), q.astype(np.float32)


def collate(items):
    """items: list of (L_i, D) arrays -> (padded (B, Lmax, D) fp32, mask (B, Lmax) fp32)."""
    lmax = max(len(x) for x in items)
    out = np.zeros((len(items), lmax) + items[0].shape[1:], np.float32)
    mask = np.zeros((len(items), lmax), np.float32)
    for i, x in enumerate(items):
        out[i, :len(x)] = x
        mask[i, :len(x)] = 1
    return out, mask
