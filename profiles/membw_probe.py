import torch, time
dev='cuda'
n=655360000//4
x=torch.empty(n,device=dev); y=torch.empty(n,device=dev)
def t(f,it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/it*1e3
us=t(lambda: x.zero_()); print('memset 655MB: %.1f us  %.0f GB/s'%(us, 655.36e6/us/1e3))
us=t(lambda: y.copy_(x)); print('copy 655MB->655MB: %.1f us  %.0f GB/s (r+w)'%(us, 2*655.36e6/us/1e3))
us=t(lambda: x.sum()); print('sum (read) 655MB: %.1f us  %.0f GB/s'%(us, 655.36e6/us/1e3))
us=t(lambda: x.fill_(1.0)); print('fill 655MB: %.1f us  %.0f GB/s'%(us, 655.36e6/us/1e3))
z=torch.empty(n*2,device=dev)
us=t(lambda: z.zero_()); print('memset 1.3GB: %.1f us  %.0f GB/s'%(us, 2*655.36e6/us/1e3))
