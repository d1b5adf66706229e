"""Pinned-memory H2D / D2H / duplex bandwidth of the box (context for the e2e number of bench.py)."""
import torch, time
d=torch.device('cuda')
h=torch.empty(64<<20,dtype=torch.uint8).pin_memory(); g=torch.empty(64<<20,dtype=torch.uint8,device=d)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for name,fn in (('h2d',lambda: g.copy_(h,non_blocking=True)),('d2h',lambda: h.copy_(g,non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    print(name, 64/1024/dt, 'GiB/s')
h2=torch.empty(64<<20,dtype=torch.uint8).pin_memory(); g2=torch.empty(64<<20,dtype=torch.uint8,device=d)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): g.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(g2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
print('duplex each', 64/1024/dt, 'GiB/s')
