"""executed warp instructions and stall samples per CUDA source line:
     python tools/ncu_lines.py rep.ncu-rep path/to/kernels.o kernel_substring [N]
The report's SASS rows (ncu --page source) are matched, instruction by instruction, with nvdisasm's line
annotations of the SAME object file the profiled library was linked from (build with -lineinfo).
BY_SAMPLES=1 sorts by stall samples instead of executed instructions."""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, obj, want = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
want_rep = os.environ.get('REP_NAME', want)      # demangled name in the report when it differs (templates)
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40

with tempfile.TemporaryDirectory() as td:
    subprocess.run(['cuobjdump', '-xelf', 'all', obj], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(td, cubin)], capture_output=True, text=True).stdout

funcs = defaultdict(list)      # section -> [(opcode, file, line)]
cur = None; f = None; ln = None
for l in dis.split('\n'):
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m: cur = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: f = os.path.basename(m.group(1)); ln = int(m.group(2)); continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)', l)
    if m and cur: funcs[cur].append((m.group(1), f, ln))
name = [k for k in funcs if want in k]
assert len(name) == 1, name
ins = funcs[name[0]]

out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; data = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        if kern is not None and data: break
        kern = r[1] if want_rep in r[1] else None; data = []; continue
    if r and r[0] == 'Address': hdr = r; continue
    if kern is not None and r: data.append(r)
assert data, 'kernel not in report'
assert len(data) == len(ins), f'report has {len(data)} instructions, object has {len(ins)}: not the same build'
ie, ismp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
st = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
per = defaultdict(lambda: [0, 0, defaultdict(int), defaultdict(int)])
for r, (op, f, ln) in zip(data, ins):
    assert op.split('.')[0] in r[isrc], (op, r[isrc])
    p = per[(f, ln)]
    p[0] += int(r[ie]); p[1] += int(r[ismp]); p[2][op.split('.')[0]] += int(r[ie])
    for i in st: p[3][hdr[i][6:]] += int(r[i])
ti = sum(p[0] for p in per.values()); ts = sum(p[1] for p in per.values())
print(f'{kern}: {ti} warp instructions, {ts} samples')
src = {}
def text(f, ln):
    if f not in src:
        for d in (os.path.dirname(obj), '/usr/local/cuda/include', '/usr/local/cuda/include/crt'):
            if os.path.exists(os.path.join(d, f)): src[f] = open(os.path.join(d, f)).read().split('\n'); break
        else: src[f] = []
    return src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ''
key = (lambda x: -x[1][1]) if os.environ.get('BY_SAMPLES') else (lambda x: -x[1][0])
for (f, ln), p in sorted(per.items(), key=key)[:N]:
    ops = ' '.join(f'{k}:{100 * v / ti:.1f}' for k, v in sorted(p[2].items(), key=lambda x: -x[1])[:4])
    why = ' '.join(f'{k}:{100 * v / max(ts, 1):.1f}' for k, v in sorted(p[3].items(), key=lambda x: -x[1])[:3] if v)
    print(f'{100 * p[0] / ti:5.1f}% inst {100 * p[1] / max(ts, 1):5.1f}% smp  {f}:{ln:<4d} {text(f, ln)[:60]:60s} {ops} | {why}')
