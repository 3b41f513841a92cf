import torch, os, sys
sys.path.insert(0, '/root/repo')
from hma_b200 import ops
B,T,n,H=8,16,320,8
qkv = torch.randn(B*T*n, 768, device='cuda').bfloat16()
def run(reps=20):
    for _ in range(3): ops.attn_temporal_fwd(qkv,B,T,n,H,0.17,False)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    # flush L2 between reps by touching a big buffer
    big = torch.empty(256*1024*1024, device='cuda', dtype=torch.uint8)
    tot=0
    for _ in range(reps):
        big.zero_(); e0.record(); ops.attn_temporal_fwd(qkv,B,T,n,H,0.17,False); e1.record(); torch.cuda.synchronize(); tot+=e0.elapsed_time(e1)
    return tot/reps*1e3
print(os.environ.get('HMA_TEMPORAL_DEBUG','0'), 'fwd us', run())
