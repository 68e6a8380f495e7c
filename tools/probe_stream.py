"""HBM streaming rate of the first-projection access shape vs. contiguous shapes (fvtg_dbg_stream_probe)."""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")   # the hook-carrying build (libflashvtg_b200_dbg.so)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    out = torch.zeros(4, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for rows, dim in ((148 * 256, 4096), (148 * 512, 768)):
        x = torch.randn(rows, dim, device=dev)
        for seg in (1, 2, 4, 8):
            if dim % (64 * seg):
                continue
            for slots in (2, 3, 4, 6):
                ts = []
                for i in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = lib.fvtg_dbg_stream_probe(x.data_ptr(), rows, dim, seg, slots, out.data_ptr(), st)
                    e1.record()
                    assert rc == 0
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                us = min(ts[1:])
                print(f"rows {rows} dim {dim} seg {seg} slots {slots}: {us:7.1f} us  {rows * dim * 4 / us / 1e6:5.2f} TB/s",
                      flush=True)


if __name__ == "__main__":
    main()
