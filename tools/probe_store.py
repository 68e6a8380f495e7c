"""Per-SM global store / load rates with the layer kernel's epilogue access shapes (fvtg_dbg_store_probe)."""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402

NAMES = {0: "fp32 blocked st (128 KB/tile)", 1: "bf16 rows st x2 (128 KB/tile)", 2: "LN2 mix st (256 KB/tile)",
         3: "fp32 blocked ld (128 KB/tile)", 4: "TMA bulk st 4 x 32 KB", 5: "TMA bulk st 16 x 8 KB",
         6: "TMA bulk st 64 x 2 KB"}
BYTES = {0: 128 << 10, 1: 128 << 10, 2: 256 << 10, 3: 128 << 10, 4: 128 << 10, 5: 128 << 10, 6: 128 << 10}


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    reps = 8
    buf = torch.zeros(148 * reps * (256 << 10), dtype=torch.uint8, device=dev)
    out = torch.zeros(148 * 2, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for grid in (1, 8, 148):
        for mode in (0, 1, 2, 3, 4, 5, 6):
            best = None
            for _ in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.fvtg_dbg_store_probe(mode, reps, grid, buf.data_ptr(), out.data_ptr(), st)
                e1.record()
                assert rc == 0
                torch.cuda.synchronize()
                o = out[: grid * 2].view(grid, 2).float().mean(0).tolist()
                us = e0.elapsed_time(e1) * 1e3
                if best is None or us < best[0]:
                    best = (us, o)
            us, o = best
            b = BYTES[mode] * reps
            print(f"grid {grid:3d} {NAMES[mode]:32s}: issue {o[0] / reps:8.0f} cyc/tile ({b / o[0]:5.1f} B/cyc/SM)  "
                  f"performed {o[1] / reps:8.0f} cyc/tile ({b / o[1]:5.1f} B/cyc/SM)  kernel {us:7.1f} us "
                  f"{b * grid / us / 1e6:5.2f} TB/s", flush=True)


if __name__ == "__main__":
    main()
