# dev tool: bit-repeatability soak of the tier-4 kernel (mbarrier / bulk-copy exchange): many matrices, repeated runs
import sys; sys.path.insert(0, '.')
import torch, jacobi_pd_b200 as jpd
bad = 0
for n, dt, B in ((256, torch.float64, 528), (177, torch.float64, 600), (131, torch.float32, 900), (253, torch.float32, 528), (200, torch.float64, 400)):
    m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 4242 + n)
    first = [t.clone() for t in jpd.diagonalize_batched(m, layout=0)]
    torch.cuda.synchronize()
    for rep in range(6):
        out = jpd.diagonalize_batched(m, layout=0); torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(first, out))
        bad += 0 if same else 1
    A = torch.triu(m) + torch.triu(m, 1).transpose(1, 2)
    ev, vec, it = first
    res = (torch.bmm(vec, A) - ev[:, :, None] * vec).abs().amax().item()
    orth = (torch.bmm(vec, vec.transpose(1, 2)) - torch.eye(n, dtype=dt, device='cuda')).abs().amax().item()
    print(f"n={n} {str(dt)[6:]} B={B}: failed {(it == 0).sum().item()} max resid {res:.2e} max orth {orth:.2e} repeatable {bad == 0}", flush=True)
print('SOAK', 'OK' if bad == 0 else f'MISMATCHES {bad}')
