import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bignn_b200 as B
from bignn_b200 import fused, ops
from bignn_b200.engine_lower import LowerOnlyEngine
from oracle import bignn_oracle as O
DEV='cuda:0'
gold='/root/repo/tests/golden'
z=np.load(gold+'/bignn_ll_gnn_step.npz')
B.set_flags(B.make_flags(model='lower_level_gnn', device=DEV))
data=B.BiGNNData.from_npz(gold+'/drugbank_packed.npz', device=DEV)
ds=O.PackedDataset.load(gold+'/drugbank_packed.npz')
lines=open(gold+'/bignn_ll_gnn_layers.txt').read().split()
om=O.OracleModel(O.parse_specs(lines), O.state_from_npz(z,'sd0/'), dtype=torch.float64)
m64, acts64, pooled64, pred64, l64 = O.lower_only_step_forward(om, ds, z['batch_gids'], z['y_true'])
l64.backward()
g64={k:v.grad.numpy() for k,v in om.params().items()}
def rel(a,b):
    a=np.asarray(a.detach().cpu().double().numpy() if isinstance(a,torch.Tensor) else a); b=np.asarray(b.detach().cpu().double().numpy() if isinstance(b,torch.Tensor) else b)
    return float(np.abs(a-b).max()/np.abs(b).max())
for fused_on in (False, True):
    model=B.Model(data).to(DEV)
    sd={k[4:]:torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith('sd0/')}
    model.load_state_dict(sd, strict=False); model.train()
    eng=LowerOnlyEngine(data, model, fused_lower=fused_on)
    rows, ids, labels = eng.stage(z['batch_gids'], z['y_true'])
    if fused_on:
        from bignn_b200.graph import MergedGraph
        mg=MergedGraph(data.packed, rows, pad_features=True)
        spec=fused.StackSpec(eng.gin_layers, eng.agg, mg, None, mg.G, None); spec.keep_acts=True
        with torch.no_grad():
            pooled=fused.gin_stack(spec, mg.x, True)
        acts=fused.level_activations(spec)
        print('fused acts vs fp64:', ['%.2e'%rel(a,b) for a,b in zip(acts, acts64)], 'pooled %.2e'%rel(pooled,pooled64))
        model.load_state_dict(sd, strict=False)
    model.zero_grad()
    loss=eng.forward(rows, ids, labels); loss.backward()
    if not fused_on:
        print('layers pooled vs fp64 %.2e'%rel(eng.last['pooled'],pooled64))
    scale={}
    for k,g in g64.items(): scale[k.split('.')[1]]=max(scale.get(k.split('.')[1],0.0), float(np.abs(g).max()))
    worst={}
    for k,p in model.named_parameters():
        if k in g64:
            e=float(np.abs(p.grad.double().cpu().numpy()-g64[k]).max())/scale[k.split('.')[1]]
            r=float(np.abs(z['grad/'+k].astype(np.float64)-g64[k]).max())/scale[k.split('.')[1]]
            worst[k]=(e,r)
    print('fused' if fused_on else 'layers', 'loss err %.2e'%abs(float(loss)-float(l64)))
    for k,(e,r) in worst.items():
        if int(k.split('.')[1])<5: print('   %-34s ours %.2e ref %.2e'%(k,e,r))
