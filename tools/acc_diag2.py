import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bignn_b200 as B
from bignn_b200 import fused, ops
from bignn_b200.engine_lower import LowerOnlyEngine
from bignn_b200.graph import MergedGraph
DEV='cuda:0'
gold=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'tests','golden')
z=np.load(gold+'/bignn_ll_gnn_step.npz')
B.set_flags(B.make_flags(model='lower_level_gnn', device=DEV))
data=B.BiGNNData.from_npz(gold+'/drugbank_packed.npz', device=DEV)
model=B.Model(data).to(DEV)
sd={k[4:]:torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
model.load_state_dict(sd, strict=False); model.train()
eng=LowerOnlyEngine(data, model, fused_lower=True)
rows, ids, labels = eng.stage(z['batch_gids'], z['y_true'])
mg=MergedGraph(data.packed, rows, pad_features=True)
spec=fused.StackSpec(eng.gin_layers, eng.agg, mg, None, mg.G, None)
class Ctx: needs_input_grad=(False,False,False,True)
ctx=Ctx()
params=spec.params()
out=fused._GinStack.forward(ctx, spec, True, mg.x, *params)
def rel(a,b): return float((a.double()-b.double()).abs().max()/b.double().abs().max())
X=mg.x.double(); fm=fa=fb=None
rp=mg.row_ptr.cpu().numpy(); ci=mg.col_idx.cpu().long()
rowid=torch.repeat_interleave(torch.arange(mg.A), torch.as_tensor(np.diff(rp))).to(DEV)
for li,(Y,Z,T,mean,rstd,a_in,a_out,din) in enumerate(ctx.saved):
    W1,b1,W2,b2,g,be=[p.detach().double() for p in params[6*li:6*li+6]]
    xin = X if li==0 else (Xprev_bn)
    agg=torch.zeros(mg.A, xin.shape[1], dtype=torch.float64, device=DEV).index_add_(0, rowid, xin[ci.to(DEV)])
    zref = xin + agg
    f={0:lambda v:v,1:torch.relu}
    tref=f[a_in](zref[:,:din]@W1.t()+b1)
    yref=f[a_out](tref@W2.t()+b2)
    print('layer',li,'Z %.2e'%rel(Z[:,:din],zref[:,:din]),'T %.2e'%rel(T,tref),'Y %.2e'%rel(Y,yref), 'nan', bool(torch.isnan(Z).any()), bool(torch.isnan(T).any()),
          'T mask mismatches', int(((T>0)!=(tref>0)).sum()), 'of', T.numel(), ' |tref| at mismatches max %.2e' % (float(tref[(T>0)!=(tref>0)].abs().max()) if int(((T>0)!=(tref>0)).sum()) else 0.0))
    mu=yref.mean(0); var=yref.var(0,unbiased=False)
    Xprev_bn=(yref-mu)/torch.sqrt(var+1e-5)*g+be
