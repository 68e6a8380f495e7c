"""Debug: TMA weight-streaming throughput per SM vs ring depth / CTA count / copy flavour, and the
tcgen05.mma issue-rate floor for N = 128 / 256 (run on a B200)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402

lib = _lib.load_debug()
lib.fvtg_dbg_tma_probe.restype = C.c_int32
lib.fvtg_dbg_tma_probe.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p]
lib.fvtg_dbg_mma_probe.restype = C.c_int32
lib.fvtg_dbg_mma_probe.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
w = torch.randn(2304, 256, device=dev).to(torch.bfloat16)   # 1.15 MB like one layer's weights
names = {0: "2D box 64x128 (16KB)", 1: "1D bulk 16KB", 2: "2D box 64x256 (32KB)", 3: "2D 16KB, 2 producers"}
for mode in (0, 1, 2, 3):
    ub = 32768 if mode == 2 else 16384
    units = 72 * 8
    for grid in (1, 148):
        for stages in ((2, 4, 6) if mode == 2 else (2, 5, 10)):
            out = torch.zeros(grid, dtype=torch.int64, device=dev)
            for _ in range(2):
                rc = lib.fvtg_dbg_tma_probe(w.data_ptr(), 2304, stages, units, grid, mode, out.data_ptr(), st)
                assert rc == 0
                torch.cuda.synchronize()
            cyc = out.float()
            print(f"{names[mode]:24s} grid {grid:4d} stages {stages:2d}: {cyc.mean().item() / units:8.1f} cycles/unit "
                  f"(max CTA {cyc.max().item() / units:8.1f}) -> {ub / (cyc.mean().item() / units):6.1f} B/cycle/SM", flush=True)
for N in (64, 128, 256):
    for nbuf in (1, 4):
        for grid in (1, 148):
            iters = 2048
            out = torch.zeros(grid, dtype=torch.int64, device=dev)
            for _ in range(2):
                rc = lib.fvtg_dbg_mma_probe(N, iters, nbuf, grid, out.data_ptr(), st)
                assert rc == 0
                torch.cuda.synchronize()
            c = out.float().mean().item() / (iters * 4)
            print(f"mma 128x{N}x16 nbuf {nbuf} grid {grid:4d}: {c:7.1f} cycles/MMA -> {128 * N * 16 * 2 / c:8.0f} flop/cycle/SM", flush=True)

print("--- probe3: honest producer loop, cluster multicast")
lib.fvtg_dbg_tma_probe3.restype = C.c_int32
lib.fvtg_dbg_tma_probe3.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]
passes = 8
for csize in (1, 2, 4):
    for nprod in (1, 2):
        for stages in (4, 8, 12):
            if stages % nprod:
                continue
            for grid in (csize, 148):
                out = torch.zeros(grid, dtype=torch.int64, device=dev)
                for _ in range(2):
                    rc = lib.fvtg_dbg_tma_probe3(w.data_ptr(), csize, nprod, stages, passes, grid, out.data_ptr(), st)
                    assert rc == 0, rc
                    torch.cuda.synchronize()
                c = out.float().mean().item() / (passes * 72)
                print(f"cluster {csize} producers {nprod} stages {stages:2d} grid {grid:4d}: {c:7.1f} cycles/unit -> "
                      f"{16384 / c:6.1f} B/cycle/SM delivered, {16384 / c / csize:6.1f} B/cycle/SM from L2", flush=True)
