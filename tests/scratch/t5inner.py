# dev tool: tier 5 (block Jacobi) with a cap on the inner sweeps of the 64x64 pivot solves (JPD_BLOCK_INNER_SWEEPS)
import sys, os; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
for n, B in ((512, 64), (1000, 16), (300, 64)):
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    m = torch.rand((B, n, n), dtype=torch.float64, device='cuda', generator=g) * 2 - 1
    m = (torch.triu(m) + torch.triu(m, 1).transpose(1, 2)).contiguous()
    out = jpd.diagonalize_batched(m); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = jpd.diagonalize_batched(m); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ref = torch.linalg.eigvalsh(m[:2]).flip(-1)
    ev, vec, it = out
    res = (vec[:2] @ m[:2] - ev[:2, :, None] * vec[:2]).abs().max().item()
    orth = (vec[:2] @ vec[:2].transpose(1, 2) - torch.eye(n, device='cuda', dtype=torch.float64)).abs().max().item()
    print(f"inner={os.environ.get('JPD_BLOCK_INNER_SWEEPS','50'):>2s} n={n} B={B}: {ms:8.2f} ms {B/ms*1e3:8.1f} /s  eval err {(ev[:2]-ref).abs().max().item():.1e} resid {res:.1e} orth {orth:.1e} failed {(it==0).sum().item()} iters {it.float().mean().item():.0f}", flush=True)
