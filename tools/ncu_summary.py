"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]} {units[i]}")
    st = [(float(r[i]), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    print('  stalls (warps per issue):', ', '.join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}"
        for v, h in sorted(st, reverse=True)[:7]))
