import sys, time; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
cases = ((4, 1<<24, 1, torch.float64), (3, 1<<23, 0, torch.float64), (8, 1<<18, 0, torch.float64), (16, 1<<18, 0, torch.float64), (16, 1<<18, 0, torch.float32), (32, 1<<17, 0, torch.float64), (32, 1<<17, 0, torch.float32), (48, 8192, 0, torch.float64), (64, 8192, 0, torch.float64), (64, 8192, 0, torch.float32), (96, 2048, 0, torch.float64), (128, 2048, 0, torch.float64), (128, 2048, 0, torch.float32), (200, 512, 0, torch.float64), (256, 512, 0, torch.float64), (256, 512, 0, torch.float32))
if len(sys.argv) > 1:
    want = set(int(x) for x in sys.argv[1].split(','))
    cases = [c for c in cases if c[0] in want]
for n, B, kind, dt in cases:
    for layout in ((1,0) if n<=6 else (0,)):
        shape = (B,n,n) if layout==0 else (n,n,B)
        m = torch.empty(shape, dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, layout, kind, 4321)
        out = jpd.diagonalize_batched(m, layout=layout); torch.cuda.synchronize()
        reps = 5 if n<=32 else 2
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): jpd.diagonalize_batched(m, layout=layout, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/reps
        es = 8 if dt==torch.float64 else 4
        by = jpd.algorithmic_bytes(n, es)
        print(f"n={n} B={B} layout={layout} {dt}: {ms:.3f} ms  {B/ms*1e3:.3e} mat/s  algGB/s={B*by/ms/1e6:.0f}  iters mean={out[2].float().mean().item():.2f} max={out[2].max().item()} fails={(out[2]==0).sum().item()}", flush=True)
    del m, out
