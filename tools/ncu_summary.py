"""Condense an .ncu-rep into the metric,value,unit CSV kept under profiles/ (first kernel of the report).

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/rN_ncu_prof_x.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit_barriers", "launch__occupancy_limit_blocks", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__registers_per_thread",
    "launch__registers_per_thread_allocated", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main(rep, out):
    # accepts the report itself or its `--page raw --csv` export (the GPU job exports on the box and drops the large report)
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        f.write("metric,value,unit\n")
        for key in ("Kernel Name", "Block Size", "Grid Size"):
            w.writerow([key, vals[col[key]], ""])
        for key in KEEP:
            if key in col:
                w.writerow([key, vals[col[key]], units[col[key]]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
