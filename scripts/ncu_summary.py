"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    units = rows[1] if len(rows) > 2 else [""] * len(hdr)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "?")[:100])
        for w in WANT:
            if w in d:
                print(f"  {w} = {d[w]} {units[hdr.index(w)]}")
        # stall breakdown (top 6)
        stalls = [(float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("_per_warp_active.pct") and v not in ("", "n/a")]
        for v, k in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {k.split('issue_stalled_')[1].split('_per_warp')[0]} = {v:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
