#!/usr/bin/env python
"""Key metrics of every kernel in one or more .ncu-rep files, as markdown (for profiles/).

    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__cluster_size", "sm__cycles_elapsed.max.per_second", "smsp__inst_executed.sum"]


def main():
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            print(f"- {rep}: unreadable")
            continue
        hdr, units = rows[0], rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        print(f"### `{rep.split('/')[-1]}`\n")
        for r in rows[2:]:
            name = r[ci["Kernel Name"]]
            grid, block = r[ci["Grid Size"]], r[ci["Block Size"]]
            print(f"- `{name[:110]}` grid {grid} block {block}")
            for k in KEYS:
                if k in ci and r[ci[k]] != "":
                    print(f"    - {k} = {r[ci[k]]} {units[ci[k]]}")
        print()


if __name__ == "__main__":
    main()
