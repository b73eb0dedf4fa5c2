import torch, sys, torch.nn.functional as F
sys.path.insert(0,'/root/repo')
from racformer_b200 import points
QG,C,P_out,p_in=3600,64,128,96
x=torch.randn(QG,p_in,C,device='cuda'); params=torch.randn(QG,C*C+P_out*p_in,device='cuda')*0.2
def chain():
    m,s=params.split([C*C,P_out*p_in],1)
    t=torch.matmul(x,m.reshape(QG,C,C)); t=F.relu(F.layer_norm(t,[p_in,C]))
    r=torch.matmul(s.reshape(QG,P_out,p_in),t); return F.relu(F.layer_norm(r,[P_out,C]))
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n*1e3
print('chain us', t(chain)); print('fused us', t(lambda: points.adaptive_mixing_core(x,params,P_out)))
