import sys, os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'tests')); sys.path.insert(0,R)
import numpy as np, torch
import util as U
O = U.O; srm = U.srm
for (W,H,D,T,K,seed,kw) in [(64,32,16,4,2,2011,dict(all_layers=True)), (39,39,1,8,4,2001,{}), (24,20,6,3,2,2002,dict(all_layers=True)), (16,16,4,2,2,2003,dict(use_blocking_factor=True, all_layers=True)), (36,10,9,3,2,2004,{})]:
    ocfg, otab, spec, ptab, batch = U.make_case(W,H,D,T,K,seed,**kw)
    o32 = U.oracle_run(ocfg, otab, batch)
    o64 = U.oracle_run(ocfg, otab, batch, dtype=torch.float64)
    c = U.cuda_run(spec, ptab, batch, numerics='closed_form')
    print('case', W,H,D,T,K,kw)
    print(' dom  cf-vs-o64 rel', U.rel_to_max(c['dom'], o64['dom']), '  o32-vs-o64', U.rel_to_max(o32['dom'], o64['dom']), ' max', np.abs(o64['dom']).max())
    print(' terms cf ', c['terms'][:4], '\n terms 64 ', o64['terms'][:4], '\n terms 32 ', o32['terms'][:4])
    if ocfg.wells: print(' qw rel', U.rel_to_max(c['qw'], o64['qw']), 'pwf rel', U.rel_to_max(c['pwfw'], o64['pwfw']))
    for k in ['gp0','gp1','gdt1','gdt2']:
        print('  ',k, 'cf-vs-o64', U.rel_to_max(c[k], o64[k]), ' o32-vs-o64', U.rel_to_max(o32[k], o64[k]), 'max64', np.abs(o64[k]).max(), 'maxcf', np.abs(c[k]).max())
