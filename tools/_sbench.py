import torch, os, sys
sys.path.insert(0, '/root/repo')
from hma_b200 import ops
frames,n,H=128,320,8
qkvs = [torch.randn(frames*n, 768, device='cuda').bfloat16() for _ in range(6)]
def run(reps=30):
    for i in range(3): ops.attn_spatial_fwd(qkvs[i],frames,n,H,0.17,True)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    tot=0
    for i in range(reps):
        e0.record(); ops.attn_spatial_fwd(qkvs[i%6],frames,n,H,0.17,True); e1.record(); torch.cuda.synchronize(); tot+=e0.elapsed_time(e1)
    return tot/reps*1e3
print(os.environ.get('HMA_STAGGER_NS','default'), 'spatial fwd us', round(run(),1))
