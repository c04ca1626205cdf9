"""Developer tool: condense ncu CSV output into the per-kernel tables committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv            # --metrics ... --csv --log-file output
    python tools/ncu_summary.py raw gpurun_out/prof_raw.csv [regex]         # `ncu -i x.ncu-rep --page raw --csv` output

`launches`: one line per (kernel, grid, block): launches, total / average device time, average DRAM MB read / written, share
of the summed device time.  `raw`: one line per captured launch with the metrics the roofline discussion uses.
"""
import csv
import io
import re
import sys
from collections import OrderedDict, defaultdict

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def read_csv(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        sys.exit(f"{path}: no ncu CSV header found")
    return list(csv.reader(io.StringIO(text[start:])))


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def to_us(value, unit):
    return value * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1.0)


def to_mb(value, unit):
    return value * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)


def launches(path):
    rows = read_csv(path)
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    per_id = OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        d = per_id.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        m, u, v = r[col["Metric Name"]], r[col["Metric Unit"]], num(r[col["Metric Value"]])
        if m == "gpu__time_duration.sum":
            d["us"] = to_us(v, u)
        elif m == "dram__bytes_read.sum":
            d["rd"] = to_mb(v, u)
        elif m == "dram__bytes_write.sum":
            d["wr"] = to_mb(v, u)
    agg = OrderedDict()
    for d in per_id.values():
        a = agg.setdefault((d["name"], d["grid"], d["block"]), defaultdict(float))
        a["n"] += 1
        a["us"] += d.get("us", 0.0)
        a["rd"] += d.get("rd", 0.0)
        a["wr"] += d.get("wr", 0.0)
    total = sum(a["us"] for a in agg.values())
    for (name, grid, block), a in agg.items():
        n = a["n"]
        print(f"{name[:72]:72s} grid {grid:16s} block {block:14s} n={int(n):4d} total {a['us']:9.1f} us avg {a['us'] / n:8.2f} us"
              f"  read {a['rd'] / n:8.2f} MB  write {a['wr'] / n:7.3f} MB  share {100 * a['us'] / total:5.1f}%")


def raw(path, pattern=None):
    rows = read_csv(path)
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    keep = [m for m in RAW_METRICS if m in col]
    print(" | ".join(["Kernel Name", "Grid Size", "Block Size"] + [f"{m} [{units[col[m]]}]" for m in keep]))
    for r in rows[2:]:
        if len(r) < len(hdr) or (pattern and not re.search(pattern, r[col["Kernel Name"]])):
            continue
        print(" | ".join([r[col["Kernel Name"]][:60], r[col["Grid Size"]], r[col["Block Size"]]] + [r[col[m]] for m in keep]))


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "raw"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
