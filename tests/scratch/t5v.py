# dev tool: time tier-5 sizes (package library)
import sys; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
for n, B, dt in ((512, 64, torch.float64), (1000, 16, torch.float64), (300, 64, torch.float64), (512, 64, torch.float32)):
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    a = torch.rand((B, n, n), dtype=dt, device='cuda', generator=g) * 2 - 1; m = (a + a.transpose(1, 2)).contiguous()
    out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); jpd.diagonalize_batched(m, layout=0, out=out); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    A = torch.triu(m[0]) + torch.triu(m[0], 1).T
    ev, vec = out[0][0], out[1][0]
    res = (vec @ A - ev[:, None] * vec).abs().max().item(); orth = (vec @ vec.T - torch.eye(n, dtype=dt, device='cuda')).abs().max().item()
    ref = torch.linalg.eigvalsh(A.double()).flip(0)
    print(f"n={n} {str(dt)[6:]} B={B}: {best:.1f} ms  {B/best*1e3:.1f} mat/s  failed {(out[2]==0).sum().item()} resid {res:.1e} orth {orth:.1e} eval err {(ev.double()-ref).abs().max().item():.1e}", flush=True)
