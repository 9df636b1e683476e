"""Condense an .ncu-rep (ncu --set full) into the columns worth committing under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_ncu_x.csv

Keeps one row per captured launch: duration, DRAM bytes, throughput percentages, occupancy limits, registers,
pipe utilisation, issue activity and every per-issue stall reason.  Never a bench value: ncu serialises and
replays the kernel (B200_PROFILING.md)."""
import csv
import io
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct",
    "gpu__dram_throughput.avg.pct", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "lts__t_sector_hit_rate.pct",
    "launch__block_size", "launch__grid_size", "launch__cluster_size", "launch__occupancy_limit", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__shared_mem_per_block", "sm__pipe_fp64_cycles_active.avg.pct",
    "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled", "sass__inst_executed_local", "smsp__sass_inst_executed_op_local",
    "sm__inst_executed_pipe_fp64", "smsp__inst_executed_pipe_fp64", "sm__sass_thread_inst_executed_op_dfma",
    "sm__sass_thread_inst_executed_op_dadd", "sm__sass_thread_inst_executed_op_dmul",
)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size") or any(h.startswith(k) for k in KEEP)]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in data:
            w.writerow([r[i] for i in cols])
    print("%s: %d launch(es), %d columns" % (out, len(data), len(cols)))


if __name__ == "__main__":
    main()
