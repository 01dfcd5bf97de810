#!/usr/bin/env python
"""Condense an `ncu --set full` report of the conv-stack kernel into the per-launch summary CSV kept under profiles/
(one column per captured launch; bench.py reads the dram__bytes rows for roofline.traffic).

    python scripts/ncu_summary.py gpurun_out/stack_full.ncu-rep profiles/r02_stack_kernel_ncu_full_summary.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "derived__memory_l1_conflicts_shared_nway",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    names = [r[col["Kernel Name"]] for r in launches]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d of a 148-candidate pass (segment %s)" % (i + 1, "layers 1-2" if i % 2 == 0 else "layers 3-7") for i in range(len(launches))])
        w.writerow(["Kernel Name", ""] + names)
        for k in KEEP:
            if k in col:
                w.writerow([k, units[col[k]]] + [r[col[k]] for r in launches])
    print("wrote", out, "launches:", len(launches))


if __name__ == "__main__":
    main()
