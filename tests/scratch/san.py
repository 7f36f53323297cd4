# dev tool for compute-sanitizer runs: tiny batches through every kernel tier
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import jacobi_pd_b200 as jpd
sizes = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [4, 7, 16, 17, 32, 33, 64, 100, 130]
for n in sizes:
    B = 5 if n <= 64 else 2
    for dt in (torch.float64,):
        m = torch.empty((B, n, n), dtype=dt, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 77)
        ev, vec, it = jpd.diagonalize_batched(m)
        ev2, _, _ = jpd.diagonalize_batched(m, calc_evecs=False, sort_criteria=3)
        torch.cuda.synchronize()
        print(n, dt, it.tolist(), float(ev.abs().max()))
print("san done")
