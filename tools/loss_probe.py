import sys, numpy as np, torch
sys.path.insert(0, ".")
from aladin_b200 import loss as L
B=8192
r=np.random.RandomState(B)
S=torch.tensor(r.standard_normal((B,B)).astype(np.float32),device="cuda")
M=torch.tensor(np.clip(r.standard_normal((B,B))*0.3,-1,1).astype(np.float32),device="cuda")
T=S*2+3
for _ in range(3):
    L.triplet_fwd_bwd(S,0.2,True); L.triplet_fwd_bwd(S,0.2,False); L.listnet_fwd_bwd(T,M)
torch.cuda.synchronize()
