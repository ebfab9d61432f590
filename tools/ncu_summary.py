"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`, no GPU needed) into a small text
file for profiles/ and, optionally, the per-launch DRAM traffic JSON bench.py reports as
roofline.traffic.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_x.txt [traffic.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_membar",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + " ".join(sys.argv)]
    traffic = []
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")][:110])
        vals = {}
        for k in KEYS:
            if k in hdr:
                vals[k] = (r[hdr.index(k)], units[hdr.index(k)])
                lines.append("  %-75s %s %s" % (k, *vals[k]))

        def to_bytes(k):
            v, u = vals[k]
            m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            return float(v) * m
        if "dram__bytes_read.sum" in vals:
            t = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            traffic.append(t)
            lines.append("  dram bytes read+write per launch: %.0f" % t)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    if len(sys.argv) > 3 and traffic:
        with open(sys.argv[3], "w") as f:
            json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic),
                       "source": "ncu --set full capture " + rep}, f)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
