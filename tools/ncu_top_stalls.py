"""Warp-stall summary of every kernel in a set of `ncu --set full --import-source on` reports: total samples by stall
reason and the hottest SASS instructions (source-level page).

usage: python tools/ncu_top_stalls.py gpurun_out/prof_r02d_*.ncu-rep > profiles/r02d_kernel_stalls.md
"""
import csv
import io
import re
import subprocess
import sys


def one(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    try:
        hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    except StopIteration:
        return
    name = rows[0][1] if rows and len(rows[0]) > 1 else rep
    name = re.sub(r"\(.*", "", name).replace("fvtg::", "").replace("void ", "")
    hdr, data = rows[hdr_i], rows[hdr_i + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    total = sum(f(r, "# Samples") for r in data)
    agg = {s: sum(f(r, s) for r in data) for s in stalls}
    issued = sum(f(r, "Instructions Executed") for r in data)
    print(f"## `{name}` ({rep.split('/')[-1]})\n")
    print(f"{int(total)} warp-state samples over {len(data)} SASS instructions, {int(issued)} warp instructions executed.\n")
    print("| stall reason | share of samples |")
    print("|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        if v > 0:
            print(f"| {k[6:]} | {100 * v / max(total, 1):.1f} % |")
    print("\n| # | samples | instruction | top stall |")
    print("|---|---|---|---|")
    for i in sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:10]:
        st = max(stalls, key=lambda s: f(data[i], s))
        src = re.sub(r"\s+", " ", data[i][ix["Source"]]).strip()
        print(f"| {i} | {int(f(data[i], '# Samples'))} | `{src[:84]}` | {st[6:]} |")
    print()


def main():
    print("# Warp-stall summaries of the hot kernels (source-level pages of the r02d `ncu --set full` reports)\n")
    print("`selected` = issuing, `long_sb` = waiting on global / TMEM / mbarrier-probe results, `short_sb` = shared memory / "
          "MUFU, `mio` / `lg` = memory-instruction queues full, `barrier` = bar.sync, `sleep` = nanosleep in a spin.\n")
    for rep in sys.argv[1:]:
        one(rep)


if __name__ == "__main__":
    main()
