import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import genjax_b200 as gj
from genjax_b200.inference.mcmc import hmc_chain
from oracle import mcmc as om, rng as orng, dists as od
F32=np.float32
@gj.gen
def model():
    x = gj.normal(0.0, 1.0) @ "x"
    gj.normal(x, 0.5) @ "y"
def lpg(q):
    x=q[:,0].astype(F32)
    lp=(od.normal_logpdf(x,F32(0),F32(1))+od.normal_logpdf(F32(1.0),x,F32(0.5))).astype(F32)
    g=(-x+(F32(1.0)-x)/F32(0.25)).astype(F32)
    return lp,g[:,None]
n=8
tr,_=model.importance(gj.split(gj.key(0),n), gj.C.kw(y=1.0), ())
q0=tr.get_choices()["x"].cpu().numpy().reshape(n,1)
for compat in (False, True):
  for it in (1,2,5):
    res=hmc_chain(gj.split(gj.key(4),n), tr, gj.S["x"], eps=0.15, L=5, n_iters=it, compat_stale_grad=compat)
    q,lp,acc,al=om.hmc_chain(lpg,q0,orng.split(orng.key(4),n),it,0.15,5,compat_stale_grad=compat)
    print("compat",compat,"iters",it)
    print(" gpu x", res.trace.get_choices()["x"].cpu().numpy())
    print(" ora x", q[:,0])
    print(" gpu alpha", res.alpha.cpu().numpy()); print(" ora alpha", al)
    print(" acc", res.accept_count.cpu().numpy(), acc)
