"""Top SASS instructions of an .ncu-rep source page by stall samples (read here, no GPU needed).
Usage: python tools/ncu_hot.py x.ncu-rep [top_n]   (needs -lineinfo + --import-source on at capture time)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
print('total samples', tot, 'instructions', len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:top]
for i in sorted(order):
    r = body[i]
    s = int(r[ix['# Samples']] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f'{i:5d} {100.0 * s / tot:5.1f}% exec {r[ix["Instructions Executed"]]:>10s}  {r[ix["Source"]].strip()[:70]:70s} '
          + ' '.join(f'{n}:{c}' for c, n in st if c))
