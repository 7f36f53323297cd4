import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
n, B = 512, 64
g = torch.Generator(device='cuda'); g.manual_seed(5)
m = torch.rand((B, n, n), dtype=torch.float64, device='cuda', generator=g) * 2 - 1
m = (torch.triu(m) + torch.triu(m, 1).transpose(1, 2)).contiguous()
out = jpd.diagonalize_batched(m); torch.cuda.synchronize()
for i in range(8):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); jpd.diagonalize_batched(m, out=out); e1.record(); torch.cuda.synchronize()
    print(i, round(e0.elapsed_time(e1), 1), "ms", flush=True)
