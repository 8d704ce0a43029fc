"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
    python scripts/launch_summary.py gpurun_out/launches.csv [first_row] [last_row]"""
import collections, csv, re, sys

def load(path):
    lines = open(path, errors="replace").read().splitlines()
    i = [n for n, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(lines[i:]):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") == "us":
            v *= 1e3
        elif r.get("Metric Unit") == "ms":
            v *= 1e6
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((int(r["ID"]), name, v, r["Grid Size"], r["Block Size"]))
    return rows

def main():
    rows = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
    rows = rows[lo:hi]
    tot, cnt = collections.Counter(), collections.Counter()
    for _, k, v, *_ in rows:
        tot[k[:80]] += v; cnt[k[:80]] += 1
    T = sum(tot.values())
    print("%d launches, %.3f ms total (serialised, cold-cache)" % (len(rows), T / 1e6))
    for k, v in tot.most_common(60):
        print("%9.1f us %5d x %7.1f us %5.1f%%  %s" % (v / 1e3, cnt[k], v / 1e3 / cnt[k], 100 * v / T, k))

if __name__ == "__main__":
    main()
