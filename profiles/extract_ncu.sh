#!/bin/bash
# profiles/extract_ncu.sh <tag>: excerpts of gpurun_out/<tag>_*.ncu-rep -> profiles/<tag>/ and profiles/current/
TAG=$1
mkdir -p profiles/$TAG profiles/current /tmp/ncu_x
cp gpurun_out/${TAG}_launches.csv profiles/$TAG/launches.csv
for K in score_tc_wide_kernel accumulate_tcx_kernel fwdbwd_warp_kernel; do
  ncu -i gpurun_out/${TAG}_$K.ncu-rep --page raw --csv 2>/dev/null > /tmp/ncu_x/$K.raw.csv
  python - "$K" "$TAG" <<'PY'
import csv, sys
k, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(f"/tmp/ncu_x/{k}.raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for out in (f"profiles/{tag}/{k}.csv", f"profiles/current/{k}.csv"):
    with open(out, "w") as f:
        f.write("metric,value,unit\n")
        for i, h in enumerate(hdr):
            if h in want:
                f.write("%s,%s,%s\n" % (h, vals[i], units[i]))
d = {h: vals[i] for i, h in enumerate(hdr)}
print(k, d["gpu__time_duration.sum"], "us | tensor", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
      "% | issue", d["smsp__issue_active.avg.pct_of_peak_sustained_active"], "% | dram r/w MB", d["dram__bytes_read.sum"],
      d["dram__bytes_write.sum"], "| xu", d.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"))
PY
done
echo "$TAG" > profiles/current/TAG
