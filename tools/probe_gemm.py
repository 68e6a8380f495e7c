"""GPU probe: the tcgen05 GEMM against torch on random bf16 operands.

Run on a B200 (gpurun).  Prints one line per shape and writes gpurun_out/probe_gemm.json.
"""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashvtg_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load_debug()
    lib.fvtg_dbg_gemm.restype = C.c_int32
    lib.fvtg_dbg_gemm.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 4 + [C.c_void_p]
    lib.fvtg_last_error.restype = C.c_char_p
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    results = []
    shapes = [(128, 256, 64), (128, 256, 256), (300, 256, 256), (9600, 256, 256),
              (9600, 256, 832), (2400, 256, 4096), (9600, 1024, 256), (9600, 256, 1024),
              (9600, 768, 256), (5000, 128, 256), (5000, 128, 128), (18944, 256, 1280)]
    ok_all = True
    for (M, N, K) in shapes:
        a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.1).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        ref = torch.relu(a.float() @ w.float().t() + bias)
        if N == 256:
            out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
        else:
            out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.fvtg_dbg_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                               M, N, K, 1, st)
        if rc != 0:
            print("rc", rc, lib.fvtg_last_error().decode())
            ok_all = False
            results.append(dict(M=M, N=N, K=K, rc=rc, err=lib.fvtg_last_error().decode()))
            continue
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("sync failed", e)
            results.append(dict(M=M, N=N, K=K, rc=-99, err=str(e)))
            ok_all = False
            break
        err = (out.float() - ref).abs().max().item()
        scale = ref.abs().max().item()
        tol = 2e-2 * scale if N != 256 else 1e-3 * scale
        good = bool(err <= tol) and bool(torch.isfinite(out.float()).all())
        ok_all &= good
        # timing
        for _ in range(3):
            lib.fvtg_dbg_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                              M, N, K, 1, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            lib.fvtg_dbg_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
                              M, N, K, 1, st)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
        line = dict(M=M, N=N, K=K, max_err=err, scale=scale, ok=good, us=us, tflops=tf)
        print(line, flush=True)
        results.append(line)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_gemm.json", "w") as f:
        json.dump(dict(ok=ok_all, results=results), f, indent=1)
    print("PROBE_GEMM", "PASS" if ok_all else "FAIL")
    return 0 if ok_all else 1


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print("elapsed", time.time() - t0)
    sys.exit(rc)
