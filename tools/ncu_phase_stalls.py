"""Per-phase warp-stall breakdown of the fused layer kernel from an `ncu --set full --import-source on` report:
the SASS stream is cut at the instructions that delimit the kernel's phases (mbarrier waits, bar.sync, tcgen05.ld / st),
and the warp-state samples of every phase are summed by stall reason.

usage: python tools/ncu_phase_stalls.py gpurun_out/prof_r02d_layer_kernel.ncu-rep > profiles/r02d_layer_kernel_stalls.md
"""
import csv
import io
import re
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr, data = rows[hdr_i], rows[hdr_i + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    src = [re.sub(r"\s+", " ", r[ix["Source"]]).strip() for r in data]
    # phase boundaries of the epilogue warps, found by their marker instructions in program order
    bars = [i for i, s in enumerate(src) if s.startswith("BAR.SYNC") and "0x1" in s]
    ldtm = [i for i, s in enumerate(src) if s.startswith("LDTM")]
    sttm8 = [i for i, s in enumerate(src) if s.startswith("STTM.x8")]
    mma = [i for i, s in enumerate(src) if s.startswith("UTCHMMA") or s.startswith("UTCQMMA")]
    first_ldtm = ldtm[0]
    phases = []
    if len(bars) >= 2 and sttm8 and mma:
        ln1_bar, ln2_bar = bars[0], bars[1]
        prelu_first = max(i for i in ldtm if i < sttm8[0])          # LDTM of the first PReLU half
        prelu_last = sttm8[-1]
        ln2_first = min(i for i in ldtm if i > prelu_last)
        phases = [("control warps (TMA producer, prologue)", 0, first_ldtm - 60),
                  ("epilogue 1 pass 1 (z1_full wait, residual add, statistics)", first_ldtm - 60, ln1_bar),
                  ("epilogue 1 pass 2 (LayerNorm-1 -> sA, Z init)", ln1_bar, prelu_first - 40),
                  ("epilogue 2 (hacc_full wait = FFN time, PReLU pieces)", prelu_first - 40, prelu_last + 20),
                  ("epilogue 3 pass 1 (z2_full wait, statistics + fp32 row)", prelu_last + 20, ln2_bar),
                  ("epilogue 3 pass 2 (bf16 operands)", ln2_bar, mma[0] - 120),
                  ("MMA issuer + kernel tail", mma[0] - 120, len(data))]
    total = sum(f(r, "# Samples") for r in data)
    print(f"# Warp-state samples of `layer_kernel` by phase ({rep.split('/')[-1]}, {int(total)} samples, {len(data)} SASS "
          "instructions)\n")
    print("Phases are cut at marker instructions of the SASS stream (approximate at the seams). The control warps and the "
          "final `bar.sync 0` hold samples of warps that only wait.\n")
    print("| phase | samples | share | top stall reasons |")
    print("|---|---|---|---|")
    for name, lo, hi in phases:
        lo, hi = max(lo, 0), min(hi, len(data))
        agg = {s: 0.0 for s in stalls}
        n = 0.0
        for r in data[lo:hi]:
            n += f(r, "# Samples")
            for s in stalls:
                agg[s] += f(r, s)
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
        print(f"| {name} | {int(n)} | {100 * n / total:.1f} % | " +
              ", ".join(f"{k[6:]} {100 * v / max(n, 1):.0f} %" for k, v in top if v > 0) + " |")
    print("\nHottest instructions:\n")
    print("| # | samples | instruction | top stall |")
    print("|---|---|---|---|")
    for i in sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:16]:
        st = max(stalls, key=lambda s: f(data[i], s))
        print(f"| {i} | {int(f(data[i], '# Samples'))} | `{src[i][:80]}` | {st[6:]} |")


if __name__ == "__main__":
    main(sys.argv[1])
