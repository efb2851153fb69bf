"""Static (no GPU) evidence for the CUDA kernels: registers / spills from `ptxas -v` and the SASS mnemonics that
prove TMA / bulk-copy / f64 reduce use, per kernel, for the library exactly as barcode_b200/build.py compiles it.

    python tools/static_report.py > profiles/static_r01_ptxas_sass.txt
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from barcode_b200 import build as b  # noqa: E402

MNEMONICS = ["UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKRED", "SYNCS", "LDGSTS", "RED.E.ADD.F64", "REDG.E.ADD.F64", "ATOMS", "MUFU.RSQ64H",
             "SHFL", "DFMA", "DMUL", "DADD", "MUFU.RCP64H", "LDS", "STS", "LDG", "STG", "BAR.SYNC", "WARPSYNC"]


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines()


def short(name):
    name = name.replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("bgpu::", "").replace("void ", "")
    return name if len(name) <= 110 else name[:107] + "..."


def main():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    rows = []
    for src in b.CU_SOURCES:
        cmd = [nvcc] + b.NVCC_FLAGS + ["-Xptxas", "-v", "-c", os.path.join(b.CSRC, src), "-o", "/tmp/static_%s.o" % src]
        r = subprocess.run(cmd, capture_output=True, text=True, check=True)
        cur = None
        for line in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = {"src": src, "name": m.group(1), "regs": 0, "spill": 0, "stack": 0, "smem": 0}
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                cur["stack"], cur["spill"] = int(m.group(1)), int(m.group(2)) + int(m.group(3))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                cur["regs"] = int(m.group(1))
                m2 = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(m2.group(1)) if m2 else 0
    # SASS mnemonic counts per kernel
    sass = collections.defaultdict(collections.Counter)
    for src in b.CU_SOURCES:
        r = subprocess.run(["cuobjdump", "-sass", "/tmp/static_%s.o" % src], capture_output=True, text=True, check=True)
        cur = None
        for line in r.stdout.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                continue
            if cur and "/*" in line:
                for mn in MNEMONICS:
                    if re.search(r"\b" + re.escape(mn), line):
                        sass[cur][mn] += 1
                sass[cur]["_total"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
    names = demangle([r_["name"] for r_ in rows])
    print("# static report: %d kernels, nvcc flags: %s" % (len(rows), " ".join(f for f in b.NVCC_FLAGS if not f.startswith("/"))))
    print("# columns: source | registers | static smem B | stack B | spill B | SASS instructions | mnemonics seen")
    spills = 0
    for r_, dn in zip(rows, names):
        c = sass.get(r_["name"], {})
        mn = " ".join("%s=%d" % (k, c[k]) for k in MNEMONICS if c.get(k))
        spills += r_["spill"] > 0
        print("%-12s %3d regs %6d smem %4d stack %4d spill %6d sass | %s\n    %s" %
              (r_["src"], r_["regs"], r_["smem"], r_["stack"], r_["spill"], c.get("_total", 0), mn, short(dn)))
    print("# kernels with register spills: %d" % spills)


if __name__ == "__main__":
    main()
