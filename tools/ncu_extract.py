"""Condense `ncu --set full` reports (gpurun_out/prof_*.ncu-rep) into the committed evidence:
profiles/<round>_ncu_<kernel>.csv (every raw metric of the first captured launch of each kernel) and a
summary table + profiles/r02_ncu_traffic.json (dram bytes per launch per kernel class, read by bench.py).

usage: python tools/ncu_extract.py r02 gpurun_out/prof_a.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
       "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
       "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic"]
CLASS = {"layer_kernel": "layer", "gemm_pair_kernel": "gemm", "gemm_kernel": "gemm", "gemm_group_kernel": "gemm",
         "mlp_chain_kernel": "gemm", "inproj_kernel": "inproj", "attn_video_kernel": "attention",
         "attn_tc_kernel": "attention", "attn_tc1_kernel": "attention"}


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    seen, rows_out, traffic = {}, [], {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[ki]).replace("fvtg::", "").replace("void ", "")
            base = re.sub(r"<.*", "", name)
            if name in seen:
                continue
            seen[name] = True
            d = dict(zip(hdr, r))
            path = os.path.join(ROOT, "profiles", f"{tag}_ncu_{re.sub(r'[^A-Za-z0-9_]+', '_', name).strip('_')}.csv")
            with open(path, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["metric", "unit", "value"])
                for h, u in zip(hdr, units):
                    if d.get(h, "") != "":
                        w.writerow([h, u, d[h]])

            def val(k):
                if k not in d:   # section-prefixed names (e.g. "TPC.TriageCompute.sm__pipe_tensor_...")
                    cand = [h for h in hdr if h.endswith("." + k)]
                    if not cand:
                        return None
                    k = cand[0]
                try:
                    v = float(d[k].replace(",", ""))
                except (KeyError, ValueError):
                    return None
                u = units[hdr.index(k)]
                scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3,
                         "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
                return v * scale
            rows_out.append((name, [val(k) for k in KEY]))
            cls = CLASS.get(base)
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            # the class's representative launch at the bench workload: gemm_pair_kernel for "gemm", attn_video_kernel for "attention"
            if cls and rd is not None and wr is not None and (cls not in traffic or base in ("gemm_pair_kernel", "attn_video_kernel")):
                traffic[cls] = rd + wr
    print(f"# ncu --set full, first captured launch of each kernel ({tag}; raw metrics per kernel in profiles/{tag}_ncu_*.csv)\n")
    short = ["us", "dram rd MB", "dram wr MB", "tensor pipe %", "tensor mem %", "tmem inst %", "issue %", "warps %",
             "dram %", "L2 %", "xu %", "regs", "grid", "block", "smem"]
    print("| kernel | " + " | ".join(short) + " |")
    print("|---|" + "---|" * len(short))
    for name, v in rows_out:
        def f(i, div=1.0, nd=1):
            return "-" if v[i] is None else f"{v[i] / div:.{nd}f}"
        print(f"| {name} | {f(0)} | {f(1, 1e6)} | {f(2, 1e6)} | " + " | ".join(f(i) for i in range(3, 11)) + " | " +
              " | ".join(f(i, 1.0, 0) for i in range(11, 15)) + " |")
    with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w") as f:
        json.dump(traffic, f)
    print("\ndram bytes per launch by class (profiles/r02_ncu_traffic.json):", traffic)


if __name__ == "__main__":
    main()
