"""Summarise ncu outputs into small tracked files under profiles/.
  launches: ncu --metrics gpu__time_duration.sum --csv log  -> per-kernel count / total / share
  full    : .ncu-rep (--set full)                          -> key metrics per captured launch
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void\s+", "", name).replace("b200sa::", "")
        rows.append((name, ns))
    agg = OrderedDict()
    for name, ns in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised: compare shares)\n")
        f.write("# source: %s ; launches=%d ; total=%.3f ms\n" % (path, len(rows), total / 1e6))
        f.write("%-72s %8s %12s %8s\n" % ("kernel", "launches", "total_ms", "share"))
        for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %8d %12.3f %7.2f%%\n" % (name[:72], cnt, ns / 1e6, 100 * ns / total))
    print("wrote", out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    idx = {w: hdr.index(w) for w in WANT if w in hdr}
    kn = hdr.index("Kernel Name")
    res = []
    stall_cols = [(i, h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
    extra = [h for h in hdr if h in ("lts__throughput.avg.pct_of_peak_sustained_elapsed",
                                     "l1tex__throughput.avg.pct_of_peak_sustained_active",
                                     "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
                                     "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__cycles_active.avg",
                                     "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                                     "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct")]
    for row in r[2:]:
        d = {"kernel": row[kn]}
        for w, i in idx.items():
            d[w] = "%s %s" % (row[i], units[i])
        for h in extra:
            d[h] = "%s %s" % (row[hdr.index(h)], units[hdr.index(h)])
        st = []
        for i, h in stall_cols:
            try:
                st.append((float(row[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        d["top_stalls_pct_of_warp_active"] = {nm: v for v, nm in st[:6]}
        res.append(d)
    with open(out, "w") as f:
        json.dump({"source": rep, "launches": res}, f, indent=1)
    print("wrote", out, len(res))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
