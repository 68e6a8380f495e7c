"""Per-kernel SASS evidence of the Blackwell-native path: counts of the tcgen05 / TMEM / TMA mnemonics (and of the
legacy HMMA) in every kernel of libflashvtg_b200.so, from `cuobjdump -sass`.

usage: python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flashvtg_b200", "libflashvtg_b200.so")
MNEMONICS = OrderedDict([
    ("UTCHMMA", r"\bUTC[A-Z]*MMA\b"),       # tcgen05.mma
    ("LDTM", r"\bLDTM\b"),                  # tcgen05.ld
    ("STTM", r"\bSTTM\b"),                  # tcgen05.st
    ("UTMALDG", r"\bUTMALDG\b"),            # cp.async.bulk.tensor (load)
    ("UTMASTG", r"\bUTMASTG\b"),            # cp.async.bulk.tensor (store)
    ("UTCBAR", r"\bUTCBAR\b"),              # tcgen05.commit
    ("SYNCS", r"\bSYNCS\b"),                # mbarrier
    ("HMMA", r"\bHMMA\b"),                  # mma.sync (legacy tensor path)
    ("MUFU.EX2", r"\bMUFU\.EX2\b"),
    ("FFMA2/FADD2", r"\bF(FMA|ADD|MUL)2\b"),  # packed fp32 pairs
    ("LDGSTS", r"\bLDGSTS\b"),              # cp.async
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("fvtg::", "").replace("void ", "")
            cur = kernels.setdefault(name, {"instr": 0, **{k: 0 for k in MNEMONICS}})
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1]
        if not re.search(r"\b[A-Z][A-Z0-9_.]+\b", body):
            continue
        cur["instr"] += 1
        for k, pat in MNEMONICS.items():
            if re.search(pat, body):
                cur[k] += 1
    print("# SASS mnemonic counts per kernel (libflashvtg_b200.so, `cuobjdump -sass`; regenerate with "
          "`python tools/sass_summary.py`)\n")
    print("UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = "
          "tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync, LDGSTS = cp.async (B200_PROFILING.md).\n")
    cols = ["instr"] + list(MNEMONICS)
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    tot = {c: 0 for c in cols}
    for name, c in kernels.items():
        print(f"| {name} | " + " | ".join(str(c[k]) for k in cols) + " |")
        for k in cols:
            tot[k] += c[k]
    print("| **total** | " + " | ".join(str(tot[k]) for k in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
