"""profiles/traffic.json from an `ncu --set full` report:  python tools/ncu_traffic.py rep.ncu-rep cfg2:reference
Sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches (one forward + one adjoint = one step)."""
import csv, json, os, subprocess, sys
rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
def col(r, name):
    i = h.index(name)
    v = float(r[i].replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
per = {}
for r in rows[2:]:
    name = r[h.index('Kernel Name')].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
    per[name] = {"read": col(r, 'dram__bytes_read.sum'), "write": col(r, 'dram__bytes_write.sum')}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'profiles', 'traffic.json')
try:
    t = json.load(open(path))
except Exception:
    t = {}
t[key] = {"bytes_per_step": sum(v["read"] + v["write"] for v in per.values()), "per_kernel": per,
          "source": f"profiles/{os.path.basename(rep).replace('prof_', '').replace('.ncu-rep', '')}_kernels_ncu_full.txt (ncu --set full, one forward + one adjoint launch)"}
json.dump(t, open(path, 'w'), indent=1)
print(json.dumps(t[key], indent=1))
