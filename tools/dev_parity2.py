import sys, os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'tests')); sys.path.insert(0,R)
import numpy as np, torch
import util as U
O = U.O; srm = U.srm
W,H,D,T,K,seed = 39,39,1,8,4,2001
ocfg, otab, spec, ptab, batch = U.make_case(W,H,D,T,K,seed)
o = U.oracle_run(ocfg, otab, batch)
c = U.cuda_run(spec, ptab, batch)
dev = torch.device('cuda',0)
eng = srm.SrmPhysics(spec, ptab)
knots = otab.c
for nm in ['p0','p1']:
    p = getattr(batch, nm).reshape(-1)
    v, d = eng.pvt_eval(p.to(dev).contiguous())
    pt = p.clone()
    vals, ders = O.pvt_eval(pt, otab, ocfg, props=(0,1), need_deriv=(0,1))
    for q in (0,1):
        print(nm, 'prop', q, 'val ulp', U.ulp_diff(v[q].cpu().numpy(), vals[q].numpy()), 'exact', np.mean(v[q].cpu().numpy()==vals[q].numpy()),
              'der ulp', U.ulp_diff(d[q].cpu().numpy(), ders[q].detach().numpy()), 'exact', np.mean(d[q].cpu().numpy()==ders[q].detach().numpy()))
bad = np.argwhere(c['dom'] != o['dom'])
print('n bad', len(bad), 'of', c['dom'].size)
p0 = batch.p0.numpy(); p1 = batch.p1.numpy()
def nk(x): return np.abs(x - knots[:,None]).min(0)
cnt_bdry = 0
for (b,k,j,i) in bad[:12]:
    print((b,k,j,i), 'cuda', c['dom'][b,k,j,i], 'orcl', o['dom'][b,k,j,i], 'p0 knotdist', nk(np.array([p0[b,k,j,i]]))[0], 'p1 kd', nk(np.array([p1[b,k,j,i]]))[0])
bm = (c['dom'] != o['dom'])
isb = np.zeros_like(bm); isb[...,0,:]=1; isb[...,-1,:]=1; isb[...,:,0]=1; isb[...,:,-1]=1
print('bad on boundary frac', (bm&isb.astype(bool)).sum()/bm.sum(), ' boundary share', isb.mean())
# per-sample distribution
print('bad per sample', bm.reshape(bm.shape[0],-1).sum(1))
for k in ['gp0','gp1']:
    d = np.abs(c[k].astype(np.float64)-o[k]); mx = np.abs(o[k]).max()
    idx = np.argsort(d.reshape(-1))[::-1][:8]
    print(k, 'max', mx)
    for f in idx:
        b,kk,j,i = np.unravel_index(f, d.shape)
        print('  ', (b,kk,j,i), 'cuda', c[k][b,kk,j,i], 'orcl', o[k][b,kk,j,i], 'p0 kd', nk(np.array([p0[b,kk,j,i]]))[0], 'p1 kd', nk(np.array([p1[b,kk,j,i]]))[0])
    # exclude near-knot cells
    near = (nk(p0.reshape(-1)).reshape(p0.shape) < 3) | (nk(p1.reshape(-1)).reshape(p0.shape) < 3)
    print(k, 'rel-to-max excluding near-knot cells', d[~near].max()/mx, ' frac near', near.mean())
