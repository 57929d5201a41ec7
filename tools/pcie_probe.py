import torch, time
n=1_000_000_000; m=1_080_000_000
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(m,dtype=torch.uint8).pin_memory(); d2=torch.empty(m,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f,reps=3):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
def h2d():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
def both():
    h2d(); d2h()
print('H2D 1.00GB ms',t(h2d)); print('D2H 1.08GB ms',t(d2h)); print('both ms',t(both))
