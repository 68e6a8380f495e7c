"""Debug: tcgen05.mma cycles per instruction for cta_group 1/2 x A-in-smem / A-in-TMEM (run on a B200)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
lib.fvtg_dbg_mma_probe2.restype = C.c_int32
lib.fvtg_dbg_mma_probe2.argtypes = [C.c_int32] * 5 + [C.c_void_p, C.c_void_p]
lib.fvtg_last_error.restype = C.c_char_p
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
iters = 2048
for cg in (1, 2):
    for ts in (0, 1):
        for N in (128, 256):
            for grid in (cg, 148):
                out = torch.zeros(grid, dtype=torch.int64, device=dev)
                for _ in range(2):
                    rc = lib.fvtg_dbg_mma_probe2(N, iters, cg, ts, grid, out.data_ptr(), st)
                    assert rc == 0, (rc, lib.fvtg_last_error())
                    torch.cuda.synchronize()
                v = out[out > 0].float()
                c = v.mean().item() / (iters * 4)
                m = 128 * cg
                print(f"cta_group {cg} {'TS' if ts else 'SS'} {m}x{N}x16 grid {grid:4d}: {c:7.1f} cycles/MMA -> "
                      f"{128 * N * 16 * 2 / c:8.0f} flop/cycle/SM", flush=True)
