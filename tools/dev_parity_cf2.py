import sys, os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'tests')); sys.path.insert(0,R)
import numpy as np, torch
import util as U
O = U.O; srm = U.srm
W,H,D,T,K,seed,kw = (24,20,6,3,2,2002,dict(all_layers=True))
if len(sys.argv) > 1: kw['near_knots'] = False
ocfg, otab, spec, ptab, batch = U.make_case(W,H,D,T,K,seed,**kw)
o64 = U.oracle_run(ocfg, otab, batch, dtype=torch.float64)
c = U.cuda_run(spec, ptab, batch, numerics='closed_form')
knots = otab.c.astype(np.float64)
p0 = batch.p0.numpy().astype(np.float64); p1 = batch.p1.numpy().astype(np.float64)
def nk(x): return np.abs(x.reshape(-1,1) - knots[None,:]).min(1).reshape(x.shape)
wells = set((w.k,w.j,w.i) for w in spec.wells)
for k in ['dom','gp0','gp1']:
    d = np.abs(c[k].astype(np.float64)-o64[k]); mx = np.abs(o64[k]).max()
    idx = np.argsort(d.reshape(-1))[::-1][:8]
    print(k, 'max', mx, 'rel', d.max()/mx)
    for f in idx:
        b,kk,j,i = np.unravel_index(f, d.shape)
        nb = [(kk,j,i-1),(kk,j,i+1),(kk,j-1,i),(kk,j+1,i),(kk-1,j,i),(kk+1,j,i)]
        print('  ', (int(b),int(kk),int(j),int(i)), 'cf', c[k][b,kk,j,i], 'o64', o64[k][b,kk,j,i], 'p0kd %.4f p1kd %.4f'%(nk(p0[b,kk,j,i:i+1])[0], nk(p1[b,kk,j,i:i+1])[0]), 'well' if (kk,j,i) in wells else '', 'nbrwell' if any(n in wells for n in nb) else '')
print('gdt1 cf', c['gdt1'], '\n o64', o64['gdt1'])
