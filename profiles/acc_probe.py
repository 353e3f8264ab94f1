import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import bignn_b200 as B
from bignn_b200 import ops
B._lib.load()
for K in (8, 32, 64, 320):
    g=torch.Generator().manual_seed(K)
    a=torch.randn(4096,K,generator=g); b=torch.randn(64,K,generator=g)
    want=(a.double()@b.double().t())
    for name,fn in (('simt',lambda: ops.gemm(a.cuda(),b.cuda(),False,True)),('tc',lambda: ops.gemm_tc(a.cuda(),b.cuda(),True))):
        got=fn().double().cpu()
        err=(got-want)
        print(K,name,'max',float(err.abs().max()/want.abs().max()),'rms',float(err.pow(2).mean().sqrt()/want.pow(2).mean().sqrt()),'mean(signed*sign(want))',float((err*want.sign()).mean()/want.abs().mean()))
    # positive inputs (like post-relu activations): bias visible
    a2=a.abs(); b2=b.abs(); want2=(a2.double()@b2.double().t())
    for name,fn in (('simt+',lambda: ops.gemm(a2.cuda(),b2.cuda(),False,True)),('tc+',lambda: ops.gemm_tc(a2.cuda(),b2.cuda(),True))):
        got=fn().double().cpu(); err=got-want2
        print(K,name,'max',float(err.abs().max()/want2.abs().max()),'rms',float(err.pow(2).mean().sqrt()/want2.pow(2).mean().sqrt()),'bias',float(err.mean()/want2.mean()))
