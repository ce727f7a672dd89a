import sys; sys.path.insert(0,'.')
import numpy as np, torch
from tests import cases
from tests.test_gpu_parity_solve import _run_pair
B=2048+37
y0,th,c=cases.mm_inputs(B,0,np.float64)
ts=np.linspace(0,100,16)
ref,out=_run_pair(cases.michaelis_menten(),y0,th,c,ts,(0,0,0,None),np.float64,kernel=1)
ys=out.ys.cpu().numpy()
d=np.abs(ys-ref.ys)/np.maximum(np.abs(ref.ys),1e-6)
for k in np.argsort(d.ravel())[::-1][:8]:
    b,t,i=np.unravel_index(k,d.shape)
    print(b,t,i,'rel',d[b,t,i],'ys',ys[b,t,i],'ref',ref.ys[b,t,i],'acc/rej gpu',out.n_accepted[b].item(),out.n_rejected[b].item(),'ref',ref.n_accepted[b],ref.n_rejected[b],'y0',y0[b],'th',th[b])
