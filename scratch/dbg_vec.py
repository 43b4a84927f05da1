import numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
import genjax_b200 as gj
from genjax_b200.inference.pf import ParticleFilter
from genjax_b200.workloads import lgssm_step_vec
from oracle import rng as orng, gfi as ogfi, smc as osmc
A_,Q_,C_,R_=0.9,1.0,1.0,0.5
d,n,T=8,5000,6
ys=osmc.simulate_lgssm(2,T,d,A_,Q_,C_,R_)
def o_step_vec(h, x_prev, q, r):
    x = h.mv_normal_diag('x', np.float32(A_) * x_prev, q)
    h.mv_normal_diag('y', np.float32(C_) * x, r)
    return x
x0=torch.randn(n,d,generator=torch.Generator().manual_seed(0))
q=torch.full((d,),Q_); r=torch.full((d,),R_)
for use_graph in (False, True):
    res=ParticleFilter(lgssm_step_vec,n).run(gj.key(5),x0,gj.C['y'].set(torch.from_numpy(ys)),shared_args=(q,r),record=True,use_graph=use_graph)
    anc=res.ancestors.cpu().numpy(); xs=res.history['state'][0].cpu().numpy(); lws=res.history['log_weights'].cpu().numpy()
    x_in=x0.numpy(); okey=orng.key(5)
    for t in range(T):
        kp,kr=osmc.pf_step_keys(okey,t)
        otr,ow=ogfi.generate(o_step_vec,orng.split(kp,n),{'y':ys[t]},(x_in,q.numpy(),r.numpy()))
        print(use_graph,t,'x err',np.abs(xs[t]-otr.choices['x']).max(),'w err',np.abs(lws[t]-ow).max(),'anc eq',np.array_equal(anc[t],osmc.resample_systematic(lws[t],kr)), 'inc', res.log_increments[t].item(), osmc.log_mean_exp(ow))
        x_in=xs[t][anc[t]]
