"""Per-phase clock64 trace of the persistent GEMM kernel's CTA 0 through fvtg_dbg_gemm (debug aid,
run on a B200): producer / MMA issuer / epilogue-leader stamps for the first tiles."""
import os
import sys

import torch

os.environ.setdefault("FVTG_DEBUG_LIB", "1")   # the hook-carrying build (libflashvtg_b200_dbg.so)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402


def main():
    M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (76800, 768, 256)))
    dev = torch.device("cuda:0")
    lib = _lib.load()
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    out = torch.empty(M * N, device=dev, dtype=torch.float32 if N == 256 else torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    for i in range(3):
        if i == 2:
            lib.fvtg_dbg_set_trace(buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.fvtg_dbg_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 0, st)
        e1.record()
        assert rc == 0
        torch.cuda.synchronize()
    lib.fvtg_dbg_set_trace(None)
    us = e0.elapsed_time(e1) * 1e3
    print(f"M {M} N {N} K {K}: {us:.1f} us, {2 * M * N * K / us / 1e6:.1f} TFLOP/s")
    t = buf.cpu()[1024:1024 + 3 * 16 * 8].view(3, 16, 8)
    t0 = int(t[1, 0, 0])
    names = {(0, 0): "PROD tile start", (0, 1): "PROD tile loads issued", (1, 0): "MMA tile start",
             (1, 1): "MMA tempty ok", (1, 2): "MMA tile issued", (2, 0): "EPI tile start",
             (2, 1): "EPI staging free", (2, 2): "EPI tfull ok", (2, 3): "EPI math done",
             (2, 4): "EPI tile done"}
    ev = []
    for role in range(3):
        for it in range(6):
            for e in range(8):
                v = int(t[role, it, e])
                if v:
                    ev.append((v - t0, it, names.get((role, e), f"{role}/{e}")))
    for c, it, nm in sorted(ev):
        print(f"{c:>9d}  tile {it}  {nm}")


if __name__ == "__main__":
    main()
