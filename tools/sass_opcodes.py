"""Per-kernel SASS opcode histogram of liba3d.so (read here, no GPU needed): the mnemonics that prove the Blackwell path
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM
alloc) next to the legacy ones (HMMA = mma.sync) and the instruction count.

    python tools/sass_opcodes.py [anytime-3d-reconstruction_b200/liba3d.so] > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..',
                                                         'anytime-3d-reconstruction_b200', 'liba3d.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
WATCH = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'MUFU',
         'FFMA2', 'FADD2', 'FMUL2', 'LDS', 'STS', 'LDG', 'STG', 'REDUX', 'SHFL', 'BAR']
kern, hist, order = None, {}, []
for ln in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', ln)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        order.append(kern)
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', ln)
    if m and kern:
        op, mods = m.group(1), m.group(2)
        hist[kern]['_total'] += 1
        hist[kern][op] += 1
        if op in ('UTCHMMA', 'UTMALDG', 'UTCBAR') and mods:
            hist[kern][op + mods] += 1
print(f'# cuobjdump -sass opcode histogram of {os.path.basename(lib)} ({len(order)} kernels; sm_100a)')
print('# columns: ' + ' '.join(WATCH))
tot = collections.Counter()
for k in order:
    h = hist[k]
    tot.update(h)
    name = demangle(k)
    name = re.sub(r'\(anonymous namespace\)::|a3d::|\(CUtensorMap_st.*', '', name)[:110]
    cols = ' '.join(f'{w}={h[w]}' for w in WATCH if h[w])
    extra = ' '.join(f'{w}={c}' for w, c in sorted(h.items()) if '.' in w)
    print(f'{name}\n    instr={h["_total"]} {cols}' + (f'\n    {extra}' if extra else ''))
print('\n# library totals: ' + ' '.join(f'{w}={tot[w]}' for w in WATCH if tot[w]) + ' ' +
      ' '.join(f'{w}={c}' for w, c in sorted(tot.items()) if '.' in w))
