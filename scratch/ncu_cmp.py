"""Side-by-side of selected raw metrics of ncu reports: python scratch/ncu_cmp.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys, io
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_xu.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print('==', f)
    print('%-90s' % 'kernel', ' '.join('%10s' % d[idx['Kernel Name']].replace('void k_unet_level', 'L')[:10] for d in data))
    for w in want:
        if w in idx:
            print('%-90s' % (w + ' [' + units[idx[w]] + ']')[:90], ' '.join('%10s' % d[idx[w]][:10] for d in data))
