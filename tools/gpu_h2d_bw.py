import torch, time
x = torch.empty(48,3,32,112,112).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): d.copy_(x, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)/10
print("H2D %.1f MB in %.2f ms = %.1f GB/s" % (x.numel()*4/1e6, ms, x.numel()*4/ms/1e6))
