"""Turn an `ncu --metrics gpu__time_duration.sum --csv --log-file x.csv` launch list into the per-kernel share table kept
under profiles/ (read here, no GPU needed).  Usage: python tools/launch_list.py x.csv out.txt "<command line that was profiled>" """
import csv
import sys
from collections import OrderedDict

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ''
rows = []
with open(src, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for rec in csv.DictReader(lines):
    if rec.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = rec['Kernel Name']
    name = name[5:] if name.startswith('void ') else name
    name = name.split('(')[0][-70:]
    ms = float(rec['Metric Value']) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(rec['Metric Unit'], 1e-6)
    rows.append((int(rec['ID']), name, ms))
tot = OrderedDict()
for _, n, ms in rows:
    c, t = tot.get(n, (0, 0.0))
    tot[n] = (c + 1, t + ms)
total = sum(t for _, t in tot.values())
with open(dst, 'w') as f:
    f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none {cmd}\n')
    f.write('# per-launch device times are cold-cache and serialised: compare SHARES, not absolutes\n')
    f.write('# kernel, launches, total ms, share\n')
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f'{n}, {c}, {t:.3f}, {100 * t / total:.1f}%\n')
    f.write('\n# id, kernel, ms\n')
    for i, n, ms in rows:
        f.write(f'{i}, {n}, {ms:.4f}\n')
print(dst, len(rows), 'launches', round(total, 2), 'ms')
