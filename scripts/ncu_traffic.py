"""Reduce an `ncu --page raw --csv` dump (scripts/ncu_r2.sh) to the per-kernel numbers the bench and DESIGN.md quote:
duration, DRAM bytes read / written per launch, tensor-pipe activity, registers, occupancy.
    python scripts/ncu_traffic.py gpurun_out/r2_ncu_raw.csv  ->  profiles/r2_ncu_summary.txt, profiles/r2_ncu_traffic.json"""
import csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_ncu_raw.csv")
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name) if name in hdr else None


def num(r, name, scale=1.0):
    i = col(name)
    if i is None or r[i] in ("", "no data", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(u, 1.0)
    return v * mult * scale


TENSOR = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
TENSOR_ALT = "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"
out, lines = {}, []
seen = {}
for r in data:
    name = re.sub(r"\(.*", "", r[col("Kernel Name")]).replace("void ", "")
    k = seen.get(name, 0); seen[name] = k + 1
    key = "%s#%d" % (name, k)
    t = num(r, "gpu__time_duration.sum")
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    tens = num(r, TENSOR) if col(TENSOR) is not None else num(r, TENSOR_ALT)
    rec = {"us": t, "dram_read": rd, "dram_write": wr, "dram_total": (rd or 0) + (wr or 0), "tensor_pipe_pct": tens,
           "regs": num(r, "launch__registers_per_thread"), "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "grid": r[col("Grid Size")] if col("Grid Size") is not None else None}
    out[key] = rec
    lines.append("%-52s %8.1f us  dram rd %7.2f MB wr %7.2f MB  (%5.0f GB/s)  tensor %5.1f%%  regs %3d  warps %5.1f%%  grid %s" % (
        key[:52], t, (rd or 0) / 1e6, (wr or 0) / 1e6, rec["dram_total"] / t / 1e3 if t else 0, tens or 0, int(rec["regs"] or 0),
        rec["warps_active_pct"] or 0, rec["grid"]))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "r2_ncu_summary.txt"), "w") as f:
    f.write("ncu --set full --clock-control none, one launch per kernel at the K2 shapes (scripts/ncu_r2.sh, scripts/prof_all.py); cold-cache, serialised\n")
    f.write("\n".join(lines) + "\n")
with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print("\n".join(lines))
