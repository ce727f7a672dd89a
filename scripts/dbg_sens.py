import sys; sys.path.insert(0,'.')
import numpy as np, torch
from tests import cases
from catalax_b200 import engine
from catalax_b200.tools import codegen as cg
from oracle import diffrax_oracle as do, symbolic as sy
model=cases.michaelis_menten()
rng=np.random.default_rng(2)
M=8
S0=rng.uniform(50,300,M);E0=rng.uniform(5,20,M)
y0=np.stack([E0,np.zeros(M),S0],1); th=np.array([0.05,0.1]); ts=np.linspace(0,100,20)
rhs,N,D,ia=sy.make_rhs(*model[:4],model[4],sens_theta=True)
ref=do.solve(rhs,y0,th,np.zeros((M,0)),ts,rtol=1e-5,atol=1e-5,dtype=np.float64,n_err=3,init_aug=ia(M,np.float64))
field=cg.generate_field(cases.field_spec(model),sens_theta=True)
m=engine.CompiledModel(field,'float64')
dev=torch.device('cuda:0'); tt=lambda a: torch.as_tensor(a,device=dev)
for kernel in (1,2):
    out=m.solve(tt(y0),tt(th),tt(np.zeros((M,0))),tt(ts),in_axes=(0,None,0,None),rtol=1e-5,atol=1e-5,tangents=True,kernel=kernel)
    sens=out.sens.cpu().numpy(); ys=out.ys.cpu().numpy()
    rs=ref.ys[...,3:].reshape(M,20,2,3).transpose(0,1,3,2)
    print('kernel',kernel,'ys diff',np.abs(ys-ref.ys[...,:3]).max(),'sens diff',np.abs(sens-rs).max(),'sens max',np.abs(rs).max(), 'acc', out.n_accepted.cpu().numpy(), ref.n_accepted)
    k=np.argmax(np.abs(sens-rs)); print(np.unravel_index(k,sens.shape), sens.ravel()[k], rs.ravel()[k])
