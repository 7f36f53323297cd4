# dev tool: one tier-5 call (n = 512, 64 matrices) for ncu captures / launch lists
import sys; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
n, B = 512, 64
g = torch.Generator(device='cuda'); g.manual_seed(5)
a = torch.rand((B, n, n), dtype=torch.float64, device='cuda', generator=g) * 2 - 1; m = (a + a.transpose(1, 2)).contiguous()
out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
print('failed', (out[2] == 0).sum().item())
