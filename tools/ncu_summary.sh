#!/bin/bash
# usage: tools/ncu_summary.sh gpurun_out/prof.ncu-rep profiles/name.txt
# Compact, committed summary of one `ncu --set full` capture: key raw metrics + stall/hot-spot table.
set -e
rep=$1; out=$2
raw=$(mktemp); src=$(mktemp)
ncu -i "$rep" --page raw --csv 2>/dev/null > "$raw"
ncu -i "$rep" --page source --csv 2>/dev/null > "$src"
{
echo "# ncu summary of $(basename "$rep") (ncu --set full --clock-control none --import-source on)"
python - "$raw" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("## launch", r[0])
    for w in want:
        if w in h:
            i = h.index(w)
            print(f"{w} = {r[i]} {u[i]}")
PY
python "$(dirname "$0")/ncu_hot.py" "$src" | head -48
} > "$out"
rm -f "$raw" "$src"
echo "wrote $out"
