# dev tool: time the round-1 L2-workspace kernel (library built from commit d468534) against the
# current block-Jacobi tier at n = 512 / 300, raw ctypes
import sys, ctypes, torch
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
def run(path, n, B):
    l = ctypes.CDLL(path)
    l.jpd_diagonalize_batched_f64.argtypes=[c_vp,c_vp,c_vp,c_vp,c_int,c_ll,c_int,c_int,c_int,c_int,c_vp]
    g=torch.Generator(device='cuda'); g.manual_seed(5)
    m=torch.rand((B,n,n),dtype=torch.float64,device='cuda',generator=g)*2-1
    m=(torch.triu(m)+torch.triu(m,1).transpose(1,2)).contiguous()
    ev=torch.empty((B,n),dtype=torch.float64,device='cuda'); vec=torch.empty((B,n,n),dtype=torch.float64,device='cuda'); it=torch.empty(B,dtype=torch.int32,device='cuda')
    st=torch.cuda.current_stream().cuda_stream
    f=lambda: l.jpd_diagonalize_batched_f64(m.data_ptr(),ev.data_ptr(),vec.data_ptr(),it.data_ptr(),n,B,0,1,1,50,st)
    assert f()==0; torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); assert f()==0; e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)
    ref=torch.linalg.eigvalsh(m[:2]).flip(-1)
    return ms, (ev[:2]-ref).abs().max().item(), int((it==0).sum())
for n,B in ((512,64),(300,64)):
    for path in ('tests/scratch/variants/lib_r2pre_tier5.so','jacobi_pd_b200/libjpd.so'):
        ms,err,fail=run(path,n,B)
        print(f"{path.split('/')[-1]:24s} n={n} B={B}: {ms:9.2f} ms  {B/ms*1e3:8.1f} matrices/s  max|eval-eigvalsh| {err:.1e} failed {fail}", flush=True)
