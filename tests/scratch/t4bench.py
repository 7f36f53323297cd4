# dev tool: time tier-4 sizes (JPD_TIER4_BLOCK=1: block-ordering cluster kernel)
import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
for n, B, dt in ((256, 528, torch.float64), (200, 528, torch.float64), (160, 528, torch.float64), (132, 528, torch.float64), (256, 528, torch.float32), (192, 528, torch.float32)):
    m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 1234)
    out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); jpd.diagonalize_batched(m, layout=0, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"block={os.environ.get('JPD_TIER4_BLOCK','0')} n={n} {str(dt)[6:]} B={B}: {ms:.1f} ms  {B/ms*1e3:.1f} mat/s  mean iters {out[2].float().mean().item():.0f} failed {(out[2]==0).sum().item()}", flush=True)
