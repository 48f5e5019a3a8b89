import torch, time
n=256<<20
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
for name,src,dst in (("H2D",h,d),("D2H",d,h)):
    for _ in range(2): dst.copy_(src,non_blocking=True); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(5): dst.copy_(src,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, "%.1f GB/s"%(n/dt/1e9))
# bidirectional
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
h2=torch.empty(n,dtype=torch.uint8).pin_memory(); d2=torch.empty(n,dtype=torch.uint8,device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print("bidir each %.1f GB/s"%(n/dt/1e9))
