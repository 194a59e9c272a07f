import os, sys, numpy as np, torch, torch.nn as nn, torch.nn.functional as F
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "safe-grid-agents_b200"), os.path.join(ROOT,"tests")): sys.path.insert(0,p)
import gridfast
from test_gpu_dqn import build_Q
dev=torch.device("cuda",0)
for env_id in ["BoatRace-v0","SideEffectsSokoban-v0","TomatoWatering-v0"]:
  for batch in (64,1000,20000):
    torch.manual_seed(batch); env=gridfast.BatchedEnv(env_id,4,seed=1); rs=np.random.RandomState(batch)
    s=torch.as_tensor(rs.randint(0,6,size=(batch,env.hw)).astype(np.uint8)).to(dev); s2=torch.as_tensor(rs.randint(0,6,size=(batch,env.hw)).astype(np.uint8)).to(dev)
    a=torch.as_tensor(rs.randint(0,4,size=batch).astype(np.uint8)).to(dev); r=torch.as_tensor(rs.choice([-1.0,2.0,49.0],size=batch)).to(dev); term=torch.as_tensor((rs.rand(batch)<0.1).astype(np.uint8)).to(dev)
    Q=build_Q(env.hw,2,100,4).to(dev); T=build_Q(env.hw,2,100,4).to(dev)
    Qs=Q(s.float()).gather(1,a.long().reshape(-1,1)).reshape(-1); nxt=T(s2.float()).max(1)[0]; nxt[term.bool()]=0
    F.mse_loss(Qs,0.99*nxt+r.float()).backward()
    ref=torch.cat([p.grad.reshape(-1) for m in Q.modules() if isinstance(m,nn.Linear) for p in (m.weight,m.bias)])
    ag=gridfast.BatchedDeepQ(env,batch_size=batch,reference_bxb_loss=False); ag.load_torch_module(Q,0); ag.load_torch_module(T,1); ag.set_tensor_cores(True); ag.learn_batch(s,a,r,s2,term); g=ag.get_grads()
    off=0; out=[]
    for m in Q.modules():
        if isinstance(m,nn.Linear):
            for prm in (m.weight,m.bias):
                n=prm.numel(); gr,gt=ref[off:off+n],g[off:off+n]; out.append("%.1e/%.1e"%(((gt-gr).norm()/gr.norm()).item(), ((gt-gr).abs().max()/gr.abs().max()).item())); off+=n
    print(env_id,batch," ".join(out), "cos=%.6f"%(torch.dot(g,ref)/(g.norm()*ref.norm())).item())
