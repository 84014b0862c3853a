"""A/B of two builds of liba3d on the SAME box in one call (the boxes' power-capped clocks differ by +-8 %, so numbers from
different gpurun calls do not compare): runs `bench.py --no-aux --no-cpu-baseline` alternately with A3D_LIB=<lib A> and
A3D_LIB=<lib B> and prints value, stage times and clock of every run.
Usage: python tests/tools/ab_libs.py liba3d_prev.so liba3d.so [rounds]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
libs = sys.argv[1:3]           # a library file name, or ENV=VALUE[,ENV=VALUE] to compare settings of the default library
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
acc = {l: [] for l in libs}
for i in range(rounds):
    for l in libs:
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--no-aux', '--no-cpu-baseline'], capture_output=True, text=True,
                           env=dict(os.environ, **(dict(kv.split('=', 1) for kv in l.split(',') if kv) if ('=' in l or l == '-') else
                                                   {'A3D_LIB': l})), cwd=ROOT)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        st = {k: round(v, 3) for k, v in d['roofline']['stage_ms'].items()}
        acc[l].append(d['value'])
        print(f"{l:18s} {d['value']:8.0f} objects/s  stages {st}  SM {d['clocks']['sm_mhz']} MHz {d['clocks'].get('power_w_median')} W", flush=True)
for l in libs:
    v = acc[l]
    print(f'{l:18s} mean {sum(v) / len(v):8.0f}  min {min(v):8.0f}  max {max(v):8.0f}')
