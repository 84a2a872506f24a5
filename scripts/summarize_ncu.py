"""Turns ncu outputs under gpurun_out/ into the small tracked summaries under profiles/.
usage: python scripts/summarize_ncu.py <round-tag> [launches.csv] [prof.ncu-rep]"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof.ncu-rep"

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "")[:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {sum(v[0] for v in agg.values())} launches, "
            f"{tot / 1000:.1f} us total (cold-cache, serialised: compare SHARES)\n")
    f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:72s} {v[0]:8d} {v[1] / 1000:10.1f} {v[1] / v[0] / 1000:8.2f} {v[1] / tot:7.3f}\n")
print(open(f"profiles/{tag}_launches_summary.txt").read()[:3000])

try:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(f"profiles/{tag}_ncu_full_summary.csv", "w") as f:
        f.write(",".join(w for w, _ in idx) + "\n")
        f.write(",".join(rows[1][i] for _, i in idx) + "\n")
        for r in rows[2:]:
            f.write(",".join((r[i].split("(")[0] if w == "Kernel Name" else r[i].replace(",", "")) for w, i in idx) + "\n")
    print(open(f"profiles/{tag}_ncu_full_summary.csv").read()[:2000])
except Exception as e:
    print("no full capture summarised:", e)
