# dev tool: time tier-3 sizes (JPD_TIER3_LEGACY=1 for the first-generation kernel)
import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
cases = ((128, 2048, torch.float64), (64, 8192, torch.float64), (96, 2048, torch.float64), (128, 2048, torch.float32), (40, 8192, torch.float64))
for n, B, dt in cases:
    m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 1234)
    out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); jpd.diagonalize_batched(m, layout=0, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"legacy={os.environ.get('JPD_TIER3_LEGACY','0')} n={n} {str(dt)[6:]} B={B}: {best:.2f} ms  {B/best*1e3:.3e} mat/s  mean iters {out[2].float().mean().item():.0f} failed {(out[2]==0).sum().item()}", flush=True)
