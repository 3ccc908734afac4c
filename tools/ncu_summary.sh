#!/bin/sh
# usage: tools/ncu_summary.sh gpurun_out/prof.ncu-rep profiles/r01_name   -> profiles/r01_name.{raw.txt,lines.txt}
set -e
REP="$1"; OUT="$2"
ncu -i "$REP" --page raw --csv > /tmp/_raw.csv 2>/dev/null
python - "$OUT.raw.txt" <<'PY'
import csv, sys
rows = list(csv.reader(open('/tmp/_raw.csv')))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc', 'sm__pipe_tensor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput', 'smsp__pcsamp_warps_issue_stalled', 'sm__cycles_elapsed.max', 'launch__shared_mem_per_block',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__waves_per_multiprocessor']
with open(sys.argv[1], 'w') as f:
    for r in rows[2:]:
        for h, u, v in zip(hdr, units, r):
            if any(w in h for w in want) and 'not_issued' not in h:
                f.write(f"{h} [{u}] = {v}\n")
        f.write("\n")
PY
ncu -i "$REP" --page source --print-source cuda,sass --csv > /tmp/_src.csv 2>/dev/null
( echo "== top source lines by stall samples =="; python tools/ncu_lines.py /tmp/_src.csv 25 s; echo; echo "== top source lines by executed warp-instructions =="; python tools/ncu_lines.py /tmp/_src.csv 25 ) > "$OUT.lines.txt"
echo "wrote $OUT.raw.txt $OUT.lines.txt"
