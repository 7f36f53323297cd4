import sys; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
for n, B, kind in ((4, 1<<24, 1), (3, 1<<24, 0)):
    for dt in (torch.float32, torch.float64):
        m = torch.empty((n,n,B), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 1, kind, 4321)
        out = jpd.diagonalize_batched(m, layout=1); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); [jpd.diagonalize_batched(m, layout=1, out=out) for _ in range(5)]; e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5; es = 4 if dt==torch.float32 else 8
        by=jpd.algorithmic_bytes(n, es)
        print(f"n={n} {dt}: {ms:.3f} ms {B/ms*1e3:.3e} mat/s  {B*by/ms/1e6:.0f} GB/s = {B*by/ms/1e6/6553*100:.1f}% of HBM roofline; iters {out[2].float().mean().item():.2f}", flush=True)
        del m, out
