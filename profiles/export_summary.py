"""Turns an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small CSV summaries
committed here:  python profiles/export_summary.py gpurun_out/<name>.ncu-rep [...]"""
import csv
import os
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']

for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.path.basename(rep).replace('.ncu-rep', '.summary.csv'))
    with open(out, 'w') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'metric', 'unit', 'value'])
        for li, r in enumerate(rows[2:]):
            for i, h in enumerate(hdr):
                if h in KEEP or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
                    w.writerow([li, h, units[i], r[i]])
    print('wrote', out)
