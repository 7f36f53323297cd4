import sys, time; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
for n, B, dt in ((300, 32, torch.float64), (500, 16, torch.float32), (500, 16, torch.float64), (1000, 4, torch.float32)):
    g = torch.Generator(device='cuda'); g.manual_seed(n)
    a = torch.randn((B, n, n), dtype=dt, device='cuda', generator=g); m = (a + a.transpose(1, 2)).contiguous()
    t0 = time.time(); out = jpd.diagonalize_batched(m); torch.cuda.synchronize(); dt_s = time.time() - t0
    w = torch.linalg.eigvalsh(m.double()).flip(1)
    err = ((out[0].double() - w).abs().max() / w.abs().max()).item()
    print(f"n={n} B={B} {dt}: {dt_s*1e3:.1f} ms total, {dt_s/B*1e3:.1f} ms/matrix, tier {jpd.lib().jpd_kernel_tier(n, 8, 1)}, converged {(out[2]>0).all().item()}, max rel |dlambda| vs eigvalsh {err:.2e}", flush=True)
