"""Turn ncu output brought back from the GPU box into the text summaries kept under profiles/ (runs here, no GPU).

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/launches_<tag>.summary.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [--json out]   > profiles/ncu_full_<tag>.txt
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]


def short(name):
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    name = name.replace("bgpu::", "")
    return re.sub(r"\(.*$", "", name).strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
        t = tot[short(r[ik])]
        t[0] += 1
        t[1] += us
    total = sum(t[1] for t in tot.values())
    print(f"{sum(t[0] for t in tot.values())} launches, {total:.1f} us in kernels (cold-cache, serialised: compare SHARES)\n")
    for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:84]:84s} {c:4d} launches {us:10.1f} us {100 * us / total:6.1f} %  {us / c:8.1f} us each")


def full(path, json_out=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    groups = OrderedDict()
    for r in rows[2:]:
        groups.setdefault(short(r[ik]), []).append(r)
    traffic = {}
    for k, rs in groups.items():
        print(f"{k}   ({len(rs)} launches; first shown)")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:68s} {rs[0][i]} {units[i]}")
        if "gpu__time_duration.sum" in hdr:
            i = hdr.index("gpu__time_duration.sum")
            print("  gpu__time_duration of all launches (%s): %s" % (units[i], ", ".join(r[i] for r in rs)))
        if "dram__bytes_read.sum" in hdr:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tr = [float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]] for r in rs]
            traffic[k] = sum(tr) / len(tr)
            print(f"  dram read+write per launch, mean over launches: {traffic[k] / 1e6:.1f} MB")
        print()
    if json_out:
        json.dump(traffic, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        jo = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
        full(sys.argv[2], jo)
