import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import liesel_ptm_b200 as lp
n=int(sys.argv[1]) if len(sys.argv)>1 else 1_000_000
C=int(sys.argv[2]) if len(sys.argv)>2 else 1024
F32=len(sys.argv)>3 and sys.argv[3]=='f32'
DT=np.float32 if F32 else np.float64
TDT=torch.float32 if F32 else torch.float64
rng=np.random.default_rng(0)
T=10
X=rng.uniform(-2,2,size=(T,n)).astype(DT); y=(rng.normal(size=n)*1.5).astype(DT)
m=lp.LocScalePTM.new_ptm(y,nparam=20,to_float32=F32)
t0=time.time()
for t in range(T):
    b=lp.ps(X[t],nbases=20,xname=f"x{t}",dtype=DT)
    tm=lp.term.f_ig(b)
    if t<5: m.loc+=tm
    else: m.scale+=tm
print("setup",time.time()-t0)
m.build()
L,P=m.param_layout()
par=np.zeros((C,P)); par[:,2:L['delta_z']]=rng.normal(size=(C,L['delta_z']-2))*0.01
par[:,L['delta_z']:L['delta_z']+19]=rng.normal(size=(C,19))*0.2
par[:,L['tau'][0]:]=1.0
p=torch.from_numpy(par.astype(DT)).cuda()
lpv=torch.empty(C,dtype=TDT,device='cuda'); g=torch.empty_like(p)
for _ in range(3): m.log_prob_and_grad(p,lpv,g)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
K=5
e0.record()
for _ in range(K): m.log_prob_and_grad(p,lpv,g)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/K
print(f"logp+grad: {ms:.2f} ms  evals/s {n*C/ms*1e3:.3e}  frac {n*C/ms*1e3*(44 if F32 else 88)/6560.3e9:.3f}")
e0.record()
for _ in range(K): m.log_prob(p,lpv)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/K
print(f"logp only: {ms:.2f} ms  evals/s {n*C/ms*1e3:.3e}")
print(lpv[:3])
