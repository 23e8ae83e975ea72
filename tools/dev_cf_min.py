import sys, os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'tests')); sys.path.insert(0,R)
import numpy as np, torch
import util as U
W,H,D,T,K = [int(a) for a in sys.argv[1:6]]
ocfg, otab, spec, ptab, batch = U.make_case(W,H,D,T,K,2011, wells=sys.argv[6] if len(sys.argv)>6 else 'default')
eng = U.srm.SrmPhysics(spec, ptab, numerics='closed_form')
d = U.to_dev(batch, 'cuda')
fw = eng.forward(want_dom=True, **d)
torch.cuda.synchronize()
print('fwd ok', fw['terms'][0,:4].cpu().numpy())
g = eng.backward(dterms=torch.tensor(U.WEIGHTS, device='cuda'), **d)
torch.cuda.synchronize()
print('bwd ok', float(g[0].abs().max()))
