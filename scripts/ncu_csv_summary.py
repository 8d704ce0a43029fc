"""One line per kernel from an `ncu --page raw --csv` export: duration, DRAM traffic/throughput, achieved occupancy,
issue-slot utilisation, top stall reasons.   python scripts/ncu_csv_summary.py file.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"): return default
    try: return float(r[i].replace(",", ""))
    except ValueError: return default
def unit(name): return units[col[name]] if name in col else ""
stall_cols = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")] or \
             [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print("%-34s %-16s %9s %8s %8s %7s %6s %6s %6s  stalls" % ("kernel", "grid", "us", "rd MB", "wr MB", "GB/s", "dram%", "occ%", "ipc"))
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")[:34]
    dur = g(r, "gpu__time_duration.sum"); du = unit("gpu__time_duration.sum")
    dur_us = dur / 1e3 if du in ("ns", "nsecond") else (dur if du in ("us", "usecond") else dur * 1e3)
    rd = g(r, "dram__bytes_read.sum"); ru = unit("dram__bytes_read.sum")
    wr = g(r, "dram__bytes_write.sum"); wu = unit("dram__bytes_write.sum")
    sc = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    rd *= sc.get(ru, 1.0); wr *= sc.get(wu, 1.0)
    stalls = sorted(((g(r, h, 0.0), h.split("issue_stalled_")[1].split("_per")[0].replace(".ratio", "")) for h in stall_cols), reverse=True)[:3]
    print("%-34s %-16s %9.1f %8.1f %8.1f %7.0f %6.1f %6.1f %6.2f  %s" % (
        name, r[col["Grid Size"]].replace(" ", ""), dur_us, rd, wr, (rd + wr) / dur_us * 1e-3 * 1e3 if dur_us else 0,
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__inst_executed.avg.per_cycle_active"), " ".join("%s=%.1f" % (n, v) for v, n in stalls)))
