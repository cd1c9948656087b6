"""Key launch metrics and stall reasons from an `ncu -i X.ncu-rep --page raw --csv` dump (first profiled kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
m = dict(zip(rows[0], zip(rows[1], rows[2])))
print(sys.argv[2] if len(sys.argv) > 2 else "")
keys = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active"]
for k in keys:
    if k in m:
        print(f"{k} [{m[k][0]}] = {m[k][1]}")
print("--- warps stalled per issue, by reason")
st = [(float(v[1]), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[1] not in ("", "n/a")]
for v, k in sorted(st, reverse=True)[:10]:
    print(f"  {v:.2f} {k}")
