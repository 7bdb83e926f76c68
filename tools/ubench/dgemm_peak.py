import torch, time, json
torch.backends.cuda.matmul.allow_tf32=False
dev='cuda'
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return best
res={}
for N in (4096,8192):
    a=torch.randn(N,N,dtype=torch.float64,device=dev); b=torch.randn(N,N,dtype=torch.float64,device=dev)
    ms=t(lambda: a@b); res[f'dgemm_{N}']=2*N**3/ms/1e9
    print(N, ms, 'ms', 2*N**3/ms/1e9,'TF')
B=1024
a=torch.randn(B,256,256,dtype=torch.float64,device=dev); b=torch.randn(B,256,256,dtype=torch.float64,device=dev)
ms=t(lambda: torch.bmm(a,b)); res['bmm_1024x256']=2*B*256**3/ms/1e9
print('bmm', ms, res['bmm_1024x256'])
ms=t(lambda: torch.bmm(a,b.transpose(1,2))); print('bmm NT', ms, 2*B*256**3/ms/1e9)
ms=t(lambda: torch.bmm(a.transpose(1,2),b)); print('bmm TN', ms, 2*B*256**3/ms/1e9)
# sustained
N=8192
a=torch.randn(N,N,dtype=torch.float64,device=dev); b=torch.randn(N,N,dtype=torch.float64,device=dev)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(60): a@b
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)
res['dgemm_8192_sustained']=60*2*N**3/ms/1e9
print('sustained', res['dgemm_8192_sustained'])
import os
os.makedirs('gpurun_out',exist_ok=True)
json.dump(res, open('gpurun_out/fp64_peaks.json','w'), indent=1)
