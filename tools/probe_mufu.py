"""MUFU (ex2.approx) vs FMA warp-instruction throughput per SM sub-partition (fvtg_dbg_mufu_probe, debug library):
the denominator for the attention kernels, whose softmax is exp-bound rather than tensor-bound."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
out = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for warps in (4, 8, 16, 32):
    for _ in range(2):
        rc = lib.fvtg_dbg_mufu_probe(warps, 2000, out.data_ptr(), st)
        assert rc == 0
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    print(f"{warps:2d} warps/SM: ex2.approx {o[0]:.2f} cycles per warp instruction per sub-partition "
          f"(= {32 / o[0]:.2f} lanes/clk/SMSP, {4 * 32 / o[0]:.1f} per SM), fma {o[1]:.2f} cycles, "
          f"cvt.rn.bf16x2.f32 {o[3]:.2f} cycles")
