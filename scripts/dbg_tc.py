import sys, json; sys.path.insert(0,'.')
import numpy as np, torch
import catalax_b200 as ctx
from catalax_b200.neural import NeuralODE, RateFlowODE
dev=torch.device('cuda:0')
for kind in ("neuralode","rateflow"):
    net = NeuralODE(4,64,3,["a","b","c","d"],activation="tanh",max_time=10.0,key=3) if kind=="neuralode" else RateFlowODE(4,3,64,3,["a","b","c","d"],activation="celu",max_time=10.0,key=3)
    rng=np.random.default_rng(4)
    B=int(sys.argv[1]) if len(sys.argv)>1 else 1000
    y0=rng.uniform(0,2,(B,4)).astype(np.float32); ts=np.linspace(0,10,16).astype(np.float32)
    y2=net.predict_arrays(ts,y0,kernel=2); o2=net.last
    y3=net.predict_arrays(ts,y0,kernel=3); o3=net.last
    scale=1e-6+1e-3*np.abs(y2)
    err=np.abs(y3-y2)/scale
    print(kind,'tc vs cuda-core: median %.3g p99 %.3g max %.3g'%(np.median(err),np.percentile(err,99),err.max()),'steps',int((o2.n_accepted+o2.n_rejected).sum()),int((o3.n_accepted+o3.n_rejected).sum()),'fails',int((o3.status!=0).sum()), 'finite', np.isfinite(y3).all())
    print(y2[0,:3],y3[0,:3])
