"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
with open(out, 'w') as f:
    f.write(f'# ncu --set full --clock-control none summary of {rep}\n')
    for vals in rows[2:]:
        f.write('\n')
        for h, u, v in zip(hdr, units, vals):
            if h in want:
                f.write(f'{h} [{u}] = {v}\n')
print(open(out).read())
