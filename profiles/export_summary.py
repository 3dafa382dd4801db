"""Turns an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small CSV summaries
committed here:  python profiles/export_summary.py gpurun_out/<name>.ncu-rep [...]

  --counters <env_steps_per_launch>   additionally writes profiles/rollout_counters.json from the FIRST report:
      the per-launch counters bench.py's roofline reads at run time (warp instructions, DRAM bytes) together
      with the hash of the kernel sources they were captured from (bench.kernel_source_hash), so that a
      later change of the kernel shows up as `counters_stale` in the bench line instead of silently."""
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

args = sys.argv[1:]
counters_for = None
if '--counters' in args:
    k = args.index('--counters')
    counters_for = int(args[k + 1])
    del args[k:k + 2]

for ri, rep in enumerate(args):
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
    if counters_for is not None and ri == 0:
        import json
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        first = dict(zip(hdr, rows[2]))
        unit = dict(zip(hdr, units))
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        cj = {'kernel': first['Kernel Name'],
              'source': f'profiles/{os.path.basename(out)} (ncu --set full --clock-control none, one launch of bench.py)',
              'source_hash': bench.kernel_source_hash(), 'env_steps_per_launch': counters_for,
              'smsp__inst_executed.sum': float(first['smsp__inst_executed.sum']),
              'dram_bytes_read': float(first['dram__bytes_read.sum']) * scale[unit['dram__bytes_read.sum']],
              'dram_bytes_write': float(first['dram__bytes_write.sum']) * scale[unit['dram__bytes_write.sum']],
              'gpu__time_duration_us_under_ncu': float(first['gpu__time_duration.sum']),
              'issue_active_pct_under_ncu': float(first['smsp__issue_active.avg.pct_of_peak_sustained_active'])}
        cpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'rollout_counters.json')
        with open(cpath, 'w') as f:
            json.dump(cj, f, indent=1)
        print('wrote', cpath)
