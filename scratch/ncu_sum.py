"""Key metrics of every kernel in an ncu report: python scratch/ncu_sum.py report.ncu-rep"""
import csv, io, subprocess, sys
want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dramR'), ('dram__bytes_write.sum', 'dramW'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('lts__t_bytes.sum', 'L2bytes'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2%'), ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__occupancy_limit_registers', 'occR'), ('launch__occupancy_limit_shared_mem', 'occS'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long'), ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_bar'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short'), ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'st_mio'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'st_lg'), ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'st_math'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'), ('smsp__inst_executed.sum', 'inst'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf'), ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wf')]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
for d in data:
    print('==', d[idx['Kernel Name']][:100])
    print('   ' + '  '.join('%s=%s%s' % (n, d[idx[m]][:9], units[idx[m]].replace('byte', 'B')[:6] if n in ('time', 'dramR', 'dramW', 'L2bytes') else '') for m, n in want if m in idx))
