import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
for dt in (torch.float64, torch.float32):
    for n in (129, 144, 160, 176, 192, 224, 256):
        B = 264
        m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 1234)
        out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); jpd.diagonalize_batched(m, layout=0, out=out); e1.record(); torch.cuda.synchronize()
        print(f"block={os.environ.get('JPD_TIER4_BLOCK','0')} {str(dt)[6:]} n={n} {e0.elapsed_time(e1):.2f}", flush=True)
