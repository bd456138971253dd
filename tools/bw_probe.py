import torch, statistics
dev='cuda'
x=torch.empty(1<<30, dtype=torch.bfloat16, device=dev).normal_()   # 2 GiB
y=torch.empty_like(x)
def t(fn,n=10):
    for _ in range(3): fn()
    ts=[]
    for _ in range(n):
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
ms=t(lambda: y.copy_(x)); print(f"copy: {2*x.numel()*2/ms/1e6:.0f} GB/s (read+write)")
ms=t(lambda: x.sum()); print(f"sum (read only): {x.numel()*2/ms/1e6:.0f} GB/s")
xf=x.view(torch.int32)
ms=t(lambda: xf.max()); print(f"max int32 (read only): {x.numel()*2/ms/1e6:.0f} GB/s")
ms=t(lambda: y.zero_()); print(f"memset (write only): {x.numel()*2/ms/1e6:.0f} GB/s")
