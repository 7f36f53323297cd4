# dev tool: one launch of the tier-4 kernel (n = 256, 66 matrices = two per resident cluster) for an ncu capture
import sys; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
n, B = 256, 66
dt = torch.float32 if len(sys.argv) > 1 and sys.argv[1] == 'f32' else torch.float64
m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 1234)
out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
print('failed', (out[2] == 0).sum().item(), 'mean iters', out[2].float().mean().item())
