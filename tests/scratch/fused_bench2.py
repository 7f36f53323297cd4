import sys; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
from jacobi_pd_b200 import _lib
B = 1 << 24
l = _lib.lib()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(reps)]; e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
# the bench workload: S from the synthetic generator (kind 1) and the matching correlation matrices are not available,
# so draw corr and build S with the adapter-free path: all_eigenpairs output vs explicit
corr = torch.rand((9, B), dtype=torch.float64, device='cuda') * 2 - 1
lam = torch.empty((B,), dtype=torch.float64, device='cuda'); q = torch.empty((4, B), dtype=torch.float64, device='cuda'); it = torch.empty((B,), dtype=torch.int32, device='cuda')
ev = torch.empty((4, B), dtype=torch.float64, device='cuda'); vec = torch.empty((4, 4, B), dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
ms = t(lambda: l.jpd_quaternion_from_correlation_f64(corr.data_ptr(), lam.data_ptr(), q.data_ptr(), it.data_ptr(), B, 1, 0, 50, st))
print(f"fused leading (prealloc): {ms:.3f} ms {B/ms*1e3:.3e}/s iters {it.float().mean().item():.2f}")
ms = t(lambda: l.jpd_quaternion_from_correlation_f64(corr.data_ptr(), ev.data_ptr(), vec.data_ptr(), it.data_ptr(), B, 1, 1, 50, st))
print(f"fused all     (prealloc): {ms:.3f} ms {B/ms*1e3:.3e}/s iters {it.float().mean().item():.2f}")
# explicit S from corr (torch ops), then the plain kernel
xx,xy,xz,yx,yy,yz,zx,zy,zz = [corr[i] for i in range(9)]
S = torch.zeros((4,4,B), dtype=torch.float64, device='cuda')
S[0,0]=(xx+yy)+zz; S[0,1]=yz-zy; S[0,2]=zx-xz; S[0,3]=xy-yx; S[1,1]=(xx-yy)-zz; S[1,2]=xy+yx; S[1,3]=zx+xz; S[2,2]=(yy-xx)-zz; S[2,3]=yz+zy; S[3,3]=(zz-xx)-yy
out = jpd.diagonalize_batched(S, layout=1)
ms = t(lambda: jpd.diagonalize_batched(S, layout=1, out=out))
print(f"plain on explicit S:      {ms:.3f} ms {B/ms*1e3:.3e}/s iters {out[2].float().mean().item():.2f}  same evals: {torch.equal(out[0], ev)}")
