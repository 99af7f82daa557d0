"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/rNN_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_X.ncu-rep|X_raw.csv profiles/rNN_full.md
"""
import collections, csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        a = agg.setdefault(row["Kernel Name"].split("(")[0][:80], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | us/launch | total ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.1f | %.3f | %.1f %% |\n" % (k, a[0], a[1] / a[0], a[1] / 1e3, 100 * a[1] / tot))
        f.write("\ntotal %.3f ms over %d launches (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n" % (tot / 1e3, sum(a[0] for a in agg.values())))


def full(src, dst):
    # src: an .ncu-rep, or the CSV that `ncu -i X.ncu-rep --page raw --csv` printed on the GPU box (scripts/gpu_profile.sh)
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        for d in data:
            f.write("### %s\n\n| metric | value | unit |\n|---|---:|---|\n" % d[ix["Kernel Name"]][:100])
            for k in KEYS:
                if k in ix:
                    f.write("| %s | %s | %s |\n" % (k, d[ix[k]], units[ix[k]]))
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
