import os, sys, torch
sys.path.insert(0, "/root/repo")
from composable_diffusion_models_b200 import _lib, steps as S
B=4096
x=torch.randn(B,1,28,28,device="cuda"); e1=torch.randn_like(x); e2=torch.randn_like(x)
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
for _ in range(5): S.step_sde(x,[e1,e2],[1.0,1.0],-5.0,3.0,1e-3,0.1,rng=(1,2),out=x)
torch.cuda.synchronize()
_lib.prof_enable(True)
for i in range(20):
    flush.zero_()
    S.step_sde(x,[e1,e2],[1.0,1.0],-5.0,3.0,1e-3,0.1,rng=(1,i),out=x)
torch.cuda.synchronize()
c=_lib.prof_summary()["step"]
print(os.environ.get("CDM_LIB_PATH","base")[-12:], "step kernel avg us (cold L2):", round(1000*c["ms"]/c["launches"],2))
_lib.prof_enable(False)
