"""top stall sites of each kernel in an ncu report:  python tools/ncu_hot.py rep.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(out.splitlines()))
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if r and r[0] == 'Address':
        cur['hdr'] = r; continue
    if cur is not None and r:
        cur['rows'].append(r)
seen = set()
for k in kern:
    if k['name'] in seen: continue
    seen.add(k['name'])
    h = k['hdr']
    isrc, ismp = h.index('Source'), h.index('# Samples')
    st = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(int(r[ismp]) for r in k['rows'])
    print('=====', k['name'], 'samples', tot)
    agg = {h[i]: sum(int(r[i]) for r in k['rows']) for i in st}
    print('  by reason:', ', '.join(f'{a[6:]} {100*b/tot:.1f}%' for a, b in sorted(agg.items(), key=lambda x: -x[1])[:9]))
    idx = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][ismp]))[:N]
    for i in sorted(idx):
        r = k['rows'][i]
        why = sorted(((int(r[j]), h[j][6:]) for j in st), reverse=True)[:3]
        print(f'  #{i:5d} {100*int(r[ismp])/tot:5.1f}%  {r[isrc].strip()[:70]:70s} ' + ' '.join(f'{n}:{c}' for c, n in why if c))
