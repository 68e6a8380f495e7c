"""CPU: the bf16-emulating oracle (tests/emulation.py) stays inside the 1e-2 budget of the fp32
oracle on scores / spans - i.e. the precision plan of the kernels (bf16 operands, fp32 accumulate /
LayerNorm / softmax / residual) is sound before any GPU is involved."""
import pytest
import torch

import emulation as E
from helpers import load_forward_index, max_rel, regen_case
from oracle import forward as O


@pytest.mark.parametrize("entry", load_forward_index()[:2] + load_forward_index()[4:5], ids=lambda e: e["file"][:-4])
def test_bf16_emulation_within_budget(entry):
    cfg, sd, batch, _ = regen_case(entry)
    a = E.forward_batch(sd, cfg, batch)
    b = O.forward_batch(sd, cfg, batch)
    for x, y in zip(a, b):
        for k in ("video_emb", "saliency", "t2vattn", "dummy_tokens", "coord"):
            assert max_rel(x[k].numpy(), y[k].numpy()) < 1e-2, k
        assert max_rel(torch.sigmoid(x["logit"]).numpy(), y["score"].numpy()) < 1e-2
