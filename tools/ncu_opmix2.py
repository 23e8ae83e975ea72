"""per-kernel opcode mix and hot-loop size from an ncu report's source page:  python tools/ncu_opmix2.py rep.ncu-rep [cells]"""
import csv, re, collections, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(out.splitlines()))
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if r and r[0] == 'Address':
        cur['hdr'] = r; continue
    if cur is not None and r:
        cur['rows'].append(r)
for k in kern:
    h = k['hdr']; isrc, iex, ismp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    tot = 0; by = collections.Counter(); smp = collections.Counter()
    for r in k['rows']:
        n = int(r[iex]); tot += n
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[isrc])
        op = m.group(2) if m else '?'
        by[op] += n; smp[op] += int(r[ismp])
    print('=====', k['name'], 'warp instr', tot, 'static', len(k['rows']), ('thread-instr/cell %.1f' % (tot * 32 / cells)) if cells else '')
    print('  '.join(f'{op} {100*n/tot:.1f}%' for op, n in by.most_common(28)))
    cnts = collections.Counter(int(r[iex]) for r in k['rows'])
    print('exec-count groups (count, static instrs):', sorted(cnts.items(), key=lambda x: -x[0] * x[1])[:8])
