"""Experiment: one 1024-video forward vs the same videos as k sub-batches on k concurrent streams
(work-conserving fill of the layer kernel's 5th-wave tail).  Prints ms/step of each variant."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flashvtg_b200 import synth
from flashvtg_b200.config import PRESETS
from flashvtg_b200.model import FlashVTGB200

cfg = PRESETS["qvh_iv2"]
sd = synth.make_state_dict(cfg, 2024)
dev = torch.device("cuda:0")
B, LV, LT = 1024, 75, 32
base = synth.make_inputs(cfg, 64, LV, LT, seed=1234)
d = {k: v.repeat(B // 64, *([1] * (v.dim() - 1))).contiguous().to(dev) for k, v in base.items()}
res = {}


def timeit(fn, steps=20, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_split(k, sizes=None):
    models = []
    for _ in range(k):
        m = FlashVTGB200(cfg).eval()
        m.load_state_dict(sd, strict=True)
        models.append(m)
    streams = [torch.cuda.Stream(device=dev) for _ in range(k)]
    if sizes is None:
        sizes = [B // k] * k
        sizes[-1] += B - sum(sizes)
    offs = [sum(sizes[:i]) for i in range(k)]
    parts = [{n: v[o:o + s].contiguous() for n, v in d.items()} for o, s in zip(offs, sizes)]

    def step():
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for m, st, p in zip(models, streams, parts):
            with torch.cuda.stream(st):
                m.infer(p["src_vid"], p["vid_len"], p["src_txt"], p["txt_len"], duration=p["duration"],
                        nms="normal", uniform_len=True)
        for st in streams:
            cur.wait_stream(st)
    return timeit(step)


m0 = FlashVTGB200(cfg).eval()
m0.load_state_dict(sd, strict=True)
res["single_1024"] = timeit(lambda: m0.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"],
                                             duration=d["duration"], nms="normal", uniform_len=True))
for k in (2, 3, 4):
    res[f"split_{k}"] = run_split(k)
res["split_2_505_519"] = run_split(2, [505, 519])
res["split_2_1010_14"] = run_split(2, [1010, 14])
print(json.dumps(res))
