# dev tool: time tier-2 sizes (run twice: JPD_TIER2_LEGACY=1 for the first-generation kernel)
import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
for n, B, dt in ((32, 131072, torch.float64), (24, 131072, torch.float64), (17, 131072, torch.float64), (32, 131072, torch.float32), (16, 262144, torch.float64)):
    m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 1234)
    out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); jpd.diagonalize_batched(m, layout=0, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"legacy={os.environ.get('JPD_TIER2_LEGACY','0')} n={n} {str(dt)[6:]} B={B}: {best:.2f} ms  {B/best*1e3:.3e} mat/s  mean iters {out[2].float().mean().item():.0f} failed {(out[2]==0).sum().item()}", flush=True)
