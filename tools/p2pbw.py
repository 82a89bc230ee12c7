import torch, time
print(torch.cuda.device_count(), torch.cuda.can_device_access_peer(0,1))
a=torch.empty(1<<30, dtype=torch.uint8, device='cuda:0'); b=torch.empty(1<<30, dtype=torch.uint8, device='cuda:1')
for _ in range(2): b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t=time.perf_counter()
for _ in range(10): b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print("copy GB/s", 10*(1<<30)/(time.perf_counter()-t)/1e9)
