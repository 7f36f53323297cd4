import sys; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
B = 1 << 24
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(reps)]; e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
corr = torch.rand((9, B), dtype=torch.float64, device='cuda') * 2 - 1
ms = t(lambda: jpd.quaternion_from_correlation(corr, layout=1))
print(f"quaternion_from_correlation (SoA, leading pair): {ms:.3f} ms  {B/ms*1e3:.3e}/s  {B*116/ms/1e6:.0f} GB/s algorithmic (116 B/matrix)")
ms = t(lambda: jpd.quaternion_from_correlation(corr, layout=1, all_eigenpairs=True))
print(f"quaternion_from_correlation (SoA, all pairs):    {ms:.3f} ms  {B/ms*1e3:.3e}/s")
i6 = torch.rand((6, B), dtype=torch.float64, device='cuda') + torch.tensor([3.,3,3,0,0,0], dtype=torch.float64, device='cuda')[:,None]
ms = t(lambda: jpd.principal_axes(i6, layout=1))
print(f"principal_axes (SoA): {ms:.3f} ms  {B/ms*1e3:.3e}/s  {B*148/ms/1e6:.0f} GB/s algorithmic (148 B/body)")
