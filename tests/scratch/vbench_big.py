# dev tool: time the large-n path: vbench_big.py n B [lib]
import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
import jacobi_pd_b200 as jpd
n=int(sys.argv[1]); B=int(sys.argv[2])
rng=np.random.default_rng(1); A=rng.uniform(-1,1,(B,n,n)); A=np.triu(A)+np.triu(A,1).swapaxes(1,2)
d=torch.from_numpy(np.ascontiguousarray(A)).cuda()
out=jpd.diagonalize_batched(d); torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); jpd.diagonalize_batched(d,out=out); e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)
ev=out[0].cpu().numpy(); lap=np.linalg.eigvalsh(A)[:,::-1]
print(f"n={n} B={B}: {ms:.1f} ms -> {B/ms*1e3:.2f} matrices/s, n_iters {out[2].cpu().numpy()[:4]}, max eval err {np.abs(ev-lap).max()/np.abs(lap).max():.2e}", flush=True)
