"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/ and, optionally, record the
kernel's DRAM traffic per launch in profiles/roofline_traffic.json (the file bench.py reads `roofline.traffic` from).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt [--traffic-key l4|tail]"""
import csv, io, json, os, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
key = sys.argv[sys.argv.index('--traffic-key') + 1] if '--traffic-key' in sys.argv else None
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
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
last = {}
with open(out, 'w') as f:
    f.write(f'# ncu --set full --clock-control none summary of {os.path.basename(rep)}\n')
    for vals in rows[2:]:
        f.write('\n')
        for h, u, v in zip(hdr, units, vals):
            if h in want:
                f.write(f'{h} [{u}] = {v}\n')
            if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and v:
                last[h] = float(v.replace(',', '')) * SCALE.get(u, 1.0)
            if h in ('Kernel Name', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):
                last[h] = v
print(open(out).read())
if key:
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'profiles', 'roofline_traffic.json')
    d = json.load(open(p)) if os.path.exists(p) else {}
    d[key] = {'dram_bytes': last['dram__bytes_read.sum'] + last['dram__bytes_write.sum'],
              'dram_bytes_read': last['dram__bytes_read.sum'], 'dram_bytes_write': last['dram__bytes_write.sum'],
              'kernel': last.get('Kernel Name'), 'source': os.path.relpath(out, os.path.join(os.path.dirname(p), '..'))}
    json.dump(d, open(p, 'w'), indent=1)
    print('updated', p, d[key])
