"""Per-phase clock64 trace of the fused layer kernel's CTA 0 (debug aid, run on a B200).

Registers a device buffer with fvtg_dbg_set_trace, runs one forward, and prints for the LAST layer
launch the cycle offsets of each pipeline event of the first tiles (MMA issuer and one epilogue
thread), so bubbles between the roles are visible without a profiler.
"""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")   # the hook-carrying build (libflashvtg_b200_dbg.so)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib, synth  # noqa: E402
from flashvtg_b200.config import PRESETS  # noqa: E402
from flashvtg_b200.model import FlashVTGB200  # noqa: E402


def main():
    cfg = PRESETS["qvh_iv2"]
    dev = torch.device("cuda:0")
    m = FlashVTGB200(cfg).eval()
    m.load_state_dict(synth.make_state_dict(cfg, 2024))
    base = synth.make_inputs(cfg, 64, 75, 32, seed=1)
    d = {k: v.repeat(16, *([1] * (v.dim() - 1))).contiguous().to(dev) for k, v in base.items()}
    lib = _lib.load()
    for _ in range(2):
        m.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"])
    torch.cuda.synchronize()
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    lib.fvtg_dbg_set_trace(buf.data_ptr())
    m.infer(d["src_vid"], d["vid_len"], d["src_txt"], d["txt_len"])
    torch.cuda.synchronize()
    lib.fvtg_dbg_set_trace(None)
    full = buf.cpu()
    ns, cyc = int(full[514] - full[512]), int(full[515] - full[513])
    if ns > 0:
        print(f"CTA lifetime after griddepcontrol.wait: {ns / 1e3:.1f} us, {cyc} cycles -> {cyc / ns * 1e3:.0f} MHz")
    t = full[:512].view(2, 8, 32)
    t0 = int(t[0, 0, 0])
    names_m = {0: "tile start", 1: "z_empty ok", 2: "a_full ok", 3: "out_proj issued", 4: "ln_ready ok",
               5: "ff1(0,1) issued", 22: "z2 commit"}
    names_e = {0: "tile start", 1: "z1_full ok", 2: "ep1 pass1 done", 3: "ep1 done (ln_ready)",
               20: "z2_full ok", 21: "drain pass done", 22: "drain done (raw_ready)",
               24: "NORM raw_ready ok", 25: "NORM half 0 done", 26: "NORM tile done"}
    for it in range(5):
        print(f"--- tile {it} (cycles since first tile start; 1 us ~ 1900 cycles at full clock)")
        ev = []
        for e in range(32):
            if int(t[0, it, e]):
                nm = names_m.get(e, f"ff2({(e - 6) // 2}) issued" if e % 2 == 0 else f"ff1({(e - 7) // 2 + 2}) issued")
                ev.append((int(t[0, it, e]) - t0, "MMA", nm))
            if int(t[1, it, e]):
                nm = names_e.get(e, f"ep2({(e - 4) // 2}) start" if e % 2 == 0 else f"ep2({(e - 5) // 2}) done")
                ev.append((int(t[1, it, e]) - t0, "EPI", nm))
        for c, role, nm in sorted(ev):
            print(f"{c:>9d}  {role}  {nm}")


if __name__ == "__main__":
    main()
