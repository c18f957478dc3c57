"""stdin: `ncu --page raw --csv`; prints, per captured launch, every metric whose name mentions a
warp stall reason (sorted), plus a few throughput metrics.  Small text for profiles/."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    sys.exit(0)
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name")
for row in rows[2:]:
    print("==", row[kn][:100])
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "not_issued" not in h and h.endswith(".pct"):
            try:
                st.append((float(row[i].replace(",", "")), h))
            except ValueError:
                pass
    st.sort(reverse=True)
    for v, h in st[:10]:
        print("   %6.2f  %s" % (v, h))
    for h in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
              "launch__registers_per_thread", "smsp__cycles_active.avg"):
        if h in hdr:
            print("   %s = %s %s" % (h, row[hdr.index(h)], units[hdr.index(h)]))
