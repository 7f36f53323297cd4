# dev tool (context only): torch.linalg.eigh (cuSOLVER batched paths) on the same inputs
import sys, time; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
for n, B in ((3, 1 << 20), (4, 1 << 20), (16, 1 << 16), (32, 1 << 15), (128, 512), (256, 128)):
    m = torch.empty((B, n, n), dtype=torch.float64, device='cuda'); jpd.fill_synthetic(m, n, B, 0, 0, 4321)
    out = jpd.diagonalize_batched(m); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); [jpd.diagonalize_batched(m, out=out) for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / 3
    Bt = min(B, 2048)                       # cusolverDnXsyevBatched rejects very large batches
    mt = m[:Bt].contiguous()
    try:
        w, v = torch.linalg.eigh(mt); torch.cuda.synchronize()
        e0.record(); w, v = torch.linalg.eigh(mt); e1.record(); torch.cuda.synchronize()
        theirs = e0.elapsed_time(e1) * (B / Bt)
        err = (w.flip(1) - out[0][:Bt]).abs().max().item()
        print(f"n={n} B={B}: jpd {ours:.3f} ms ({B/ours*1e3:.3e}/s)  torch.linalg.eigh (batch {Bt}, scaled) {theirs:.3f} ms ({B/theirs*1e3:.3e}/s)  x{theirs/ours:.1f}  max|dl|={err:.2e}", flush=True)
    except Exception as ex:
        print(f"n={n}: torch.linalg.eigh failed: {str(ex)[:80]}", flush=True)
