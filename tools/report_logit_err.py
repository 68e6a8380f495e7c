"""Per-fixture max-norm relative error of the head logits (cls / conf / ASR logit / sigmoid score) of the CUDA
path against the fp32 oracle and against the bf16-emulating oracle (tests/emulation.py): the evidence behind the
logit tolerance written in tests/test_gpu_parity.py.  Prints a markdown table."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emulation as E  # noqa: E402
from helpers import load_forward_index, max_rel, regen_case  # noqa: E402

from flashvtg_b200.model import FlashVTGB200  # noqa: E402
from oracle import forward as O  # noqa: E402

dev = torch.device("cuda:0")
print("| fixture | video | cls vs fp32 | conf vs fp32 | logit vs fp32 | score vs fp32 | cls: emu vs fp32 | conf: emu vs fp32 |")
print("|---|---|---|---|---|---|---|---|")
worst = {}
for entry in load_forward_index():
    cfg, sd, batch, gold = regen_case(entry)
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(sd, strict=True)
    r = m.infer(batch["src_vid"].to(dev), batch["vid_len"].to(dev), batch["src_txt"].to(dev), batch["txt_len"].to(dev),
                duration=batch["duration"].to(dev), want_heads=True)
    outs = O.forward_batch(sd, cfg, batch)
    emu = E.forward_batch(sd, cfg, batch) if hasattr(E, "forward_batch") else None
    x = float(sd["x"])
    for b, o in enumerate(outs):
        n = o["logit"].shape[0]
        cls, conf = r.cls_logit[b, :n].cpu(), r.conf_logit[b, :n].cpu()
        logit = x * cls + (1 - x) * conf
        row = [max_rel(cls.numpy(), o["cls"].numpy()), max_rel(conf.numpy(), o["conf"].numpy()),
               max_rel(logit.numpy(), o["logit"].numpy()),
               max_rel(torch.sigmoid(logit).numpy(), o["score"].numpy())]
        if emu is not None:
            row += [max_rel(emu[b]["cls"].numpy(), o["cls"].numpy()), max_rel(emu[b]["conf"].numpy(), o["conf"].numpy())]
        else:
            row += [float("nan")] * 2
        for k, v in zip(("cls", "conf", "logit", "score", "emu_cls", "emu_conf"), row):
            worst[k] = max(worst.get(k, 0.0), v)
        print(f"| {entry['file'][:-4]} | {b} | " + " | ".join(f"{v:.2e}" for v in row) + " |")
print("\nworst:", {k: f"{v:.2e}" for k, v in worst.items()})
