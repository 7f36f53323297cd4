"""Helpers shared by the -m gpu parity tests: run the CUDA path through the C ABI."""
from __future__ import annotations

import numpy as np


def run_gpu(mat_aos: np.ndarray, sort=1, calc=True, sweeps=50, layout=0, host_api=False):
    """mat_aos (B,n,n) numpy -> (evals (B,n), evecs (B,n,n) or None, iters (B,)) numpy, AoS view."""
    import torch
    import jacobi_pd_b200 as jpd
    B, n, _ = mat_aos.shape
    src = np.ascontiguousarray(mat_aos if layout == 0 else mat_aos.transpose(1, 2, 0))
    if host_api:
        es, vs = jpd.api.out_shapes(n, B, layout)
        evals = np.empty(es, src.dtype)
        evecs = np.empty(vs, src.dtype) if calc else None
        iters = np.empty(B, np.int32)
        jpd.diagonalize_batched_host(src, evals, evecs, iters, n, B, layout, sort, calc, sweeps)
    else:
        d = torch.from_numpy(src).cuda()
        ev, vec, it = jpd.diagonalize_batched(d, layout=layout, sort_criteria=sort, calc_evecs=calc,
                                              max_num_sweeps=sweeps)
        torch.cuda.synchronize()
        evals = ev.cpu().numpy()
        evecs = vec.cpu().numpy() if calc else None
        iters = it.cpu().numpy()
    if layout == 1:
        evals = np.ascontiguousarray(evals.T)
        if calc:
            evecs = np.ascontiguousarray(evecs.transpose(2, 0, 1))
    return evals, evecs, iters
