import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
dt = torch.float64
for n in (36, 40, 44, 48, 52, 68, 72, 80, 88, 96, 100):
    B = 4096 if n <= 64 else 1184
    m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 1234)
    out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); jpd.diagonalize_batched(m, layout=0, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"mode={os.environ.get('JPD_TIER3_LEGACY','-')} n={n} {best:.2f}", flush=True)
