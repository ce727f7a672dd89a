import sys; sys.path.insert(0,'.')
import numpy as np, torch
from tests import cases
from tests.test_gpu_parity_solve import _run_pair
from oracle import diffrax_oracle as do, symbolic as sy
B=2048+37
m=cases.michaelis_menten()
rhs,*_=sy.make_rhs(*m[:4],m[4])
for T in (16,200):
    y0,th,c=cases.mm_inputs(B,0,np.float32)
    ts=np.linspace(0,100,T).astype(np.float32)
    truth=do.solve(rhs,y0.astype(np.float64),th.astype(np.float64),c.astype(np.float64),ts.astype(np.float64),rtol=1e-10,atol=1e-10,dtype=np.float64,max_steps=100000).ys
    for kernel in (1,2):
        ref,out=_run_pair(m,y0,th,c,ts,(0,0,0,None),np.float32,kernel=kernel)
        ys=out.ys.cpu().numpy().astype(np.float64)
        scale=1e-6+1e-6*np.abs(truth)
        eg=np.abs(ys-truth)/scale; er=np.abs(ref.ys-truth)/scale; d=np.abs(ys-ref.ys)/scale
        print('T',T,'kernel',kernel,'gpu err/scale max %.1f p99.9 %.1f mean %.2f | oracle32 max %.1f p99.9 %.1f mean %.2f | gpu-vs-oracle max %.1f p99.9 %.1f'%(eg.max(),np.percentile(eg,99.9),eg.mean(),er.max(),np.percentile(er,99.9),er.mean(),d.max(),np.percentile(d,99.9)))
        k=np.argmax(d); bb,t,i=np.unravel_index(k,d.shape)
        print('  worst',bb,t,i,ys[bb,t,i],ref.ys[bb,t,i],truth[bb,t,i],'steps gpu',out.n_accepted[bb].item(),out.n_rejected[bb].item(),'ref',ref.n_accepted[bb],ref.n_rejected[bb])
        sg=(out.n_accepted+out.n_rejected).cpu().numpy().sum(); sr=(ref.n_accepted+ref.n_rejected).sum(); print('  steps',sg,sr)
