"""Debug: TMA weight-streaming throughput per SM vs ring depth and CTA count (run on a B200)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402

lib = C.CDLL(str(_lib.LIB_PATH))
lib.fvtg_dbg_tma_probe.restype = C.c_int32
lib.fvtg_dbg_tma_probe.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
dev = torch.device("cuda:0")
w = torch.randn(2304, 256, device=dev).to(torch.bfloat16)   # 1.15 MB like one layer's weights
units = 72 * 8
for grid in (1, 16, 74, 148):
    for stages in (2, 3, 5, 7, 10, 13):
        out = torch.zeros(grid, dtype=torch.int64, device=dev)
        for _ in range(2):
            rc = lib.fvtg_dbg_tma_probe(w.data_ptr(), 2304, stages, units, grid, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
            assert rc == 0
            torch.cuda.synchronize()
        cyc = out.float()
        print(f"grid {grid:4d} stages {stages:2d}: {cyc.mean().item() / units:8.1f} cycles / 16 KB unit "
              f"(max CTA {cyc.max().item() / units:8.1f})  -> {16384 / (cyc.mean().item() / units):6.1f} B/cycle/SM")
