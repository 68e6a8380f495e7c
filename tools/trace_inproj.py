"""Timing and a numeric spot check of the fused first-projection kernel alone through fvtg_dbg_inproj
(tuning aid, run on a B200).  Usage: trace_inproj.py [rows dim]"""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")   # the hook-carrying build (libflashvtg_b200_dbg.so)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402


def run(lib, rows, dim):
    dev = torch.device("cuda:0")
    dim_pad = (dim + 63) // 64 * 64
    x = torch.randn(rows, dim, device=dev)
    wg = torch.zeros(256, dim_pad, device=dev)
    wg[:, :dim] = torch.randn(256, dim, device=dev) / dim ** 0.5
    wg = wg.to(torch.bfloat16)
    wsum = wg.float().sum(1).contiguous()
    cf = torch.zeros(256, device=dev)
    g1 = torch.ones(256, device=dev)
    b1 = torch.zeros(256, device=dev)
    out = torch.empty(rows, 256, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    times = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.fvtg_dbg_inproj(x.data_ptr(), rows, dim, dim_pad, wg.data_ptr(), wsum.data_ptr(), cf.data_ptr(),
                                 g1.data_ptr(), b1.data_ptr(), out.data_ptr(), st)
        e1.record()
        assert rc == 0
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    us = min(times[1:5])
    print(f"rows {rows} dim {dim}: {us:.1f} us, {rows * dim * 4 / us / 1e6:.2f} TB/s of features")
    # reference check of a few rows (fp32 torch)
    xs = x[:256].double()
    ln = (xs - xs.mean(1, keepdim=True)) / torch.sqrt(xs.var(1, unbiased=False, keepdim=True) + 1e-5)
    y = torch.relu(ln @ wg[:, :dim].double().T)
    y = (y - y.mean(1, keepdim=True)) / torch.sqrt(y.var(1, unbiased=False, keepdim=True) + 1e-5)
    print("  max |err| on 256 rows:", float((out[:256].double() - y).abs().max()))


def main():
    lib = _lib.load()
    if len(sys.argv) > 2:
        run(lib, int(sys.argv[1]), int(sys.argv[2]))
    else:
        run(lib, 76800, 770)
        run(lib, 32768, 4096)


if __name__ == "__main__":
    main()
