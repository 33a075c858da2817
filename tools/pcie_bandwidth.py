import torch, time
x=torch.empty(256<<20,dtype=torch.uint8).pin_memory(); d=torch.empty_like(x,device='cuda')
for n in (2<<20, 22<<20, 256<<20):
    torch.cuda.synchronize(); 
    for _ in range(3): d[:n].copy_(x[:n],non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): d[:n].copy_(x[:n],non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    print('H2D',n>>20,'MB',round(n/dt/1e9,1),'GB/s')
    t=time.perf_counter()
    for _ in range(10): x[:n].copy_(d[:n],non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    print('D2H',n>>20,'MB',round(n/dt/1e9,1),'GB/s')
