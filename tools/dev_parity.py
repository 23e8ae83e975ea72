import sys; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'tests')); sys.path.insert(0,R)
import numpy as np, torch, time
import util as U
for (W,H,D,T,K,seed,kw) in [(39,39,1,8,4,2001,{}), (24,20,6,3,2,2002,dict(all_layers=True)), (16,16,4,2,2,2003,dict(use_blocking_factor=True, all_layers=True))]:
    ocfg, otab, spec, ptab, batch = U.make_case(W,H,D,T,K,seed,**kw)
    o = U.oracle_run(ocfg, otab, batch)
    o64 = U.oracle_run(ocfg, otab, batch, dtype=torch.float64)
    c = U.cuda_run(spec, ptab, batch)
    print('case', W,H,D,T,K,kw)
    print(' dom ulp', U.ulp_diff(c['dom'], o['dom']), 'rel', U.rel_to_max(c['dom'], o['dom']), 'exact frac', np.mean(c['dom']==o['dom']))
    print(' terms cuda', c['terms'][:4], '\n terms orcl', o['terms'][:4], '\n counts', c['counts'])
    print(' qw rel', U.rel_to_max(c['qw'], o['qw']), 'pwf rel', U.rel_to_max(c['pwfw'], o['pwfw']))
    for k in ['gp0','gp1','gdt1','gdt2']:
        print(' ',k, 'rel-to-max vs o32', U.rel_to_max(c[k], o[k]), ' o32 vs o64', U.rel_to_max(o[k], o64[k]), 'max', np.abs(o[k]).max())
