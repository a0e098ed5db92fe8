"""Summaries of ncu output for profiles/ (run here, on files a gpurun call brought back).
    python tools/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel aggregate of a --metrics gpu__time_duration.sum pass
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [regex]        # selected metrics of a --set full capture (needs ncu here)"""
import csv, io, re, subprocess, sys
from collections import defaultdict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]


def rows_of(text):
    lines = text.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def launches(path):
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows_of(open(path).read()):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        a = agg[r["Kernel Name"]]
        a[0] += 1; a[1] += us
    total = sum(a[1] for a in agg.values())
    print("# columns: launches, mean us, total us, share of all captured GPU time, kernel\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:5d} {t / n:10.2f} {t:12.1f} {100 * t / total:6.2f}%  {k[:150]}")


def full(path, pattern=None):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(header, r))
        u = dict(zip(header, units))
        if pattern and not re.search(pattern, d["Kernel Name"]):
            continue
        print(d["Kernel Name"][:140])
        for m in KEEP:
            if m in d:
                print(f"    {m:100s} {d[m]:>16s} {u[m]}")
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u["dram__bytes_write.sum"]]
            print(f"    -> DRAM traffic {(rd + wr) / 1e6:.1f} Mbyte per launch\n")
        except Exception:
            print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
