# dev tool: time variant builds of libjpd.so (raw ctypes, no package) on n=3/n=4 SoA fp64
import sys, ctypes, torch
libs = sys.argv[1:]
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
def run(path, n, B, kind):
    l = ctypes.CDLL(path)
    l.jpd_fill_synthetic_f64.argtypes=[c_vp,c_int,c_ll,c_int,c_int,ctypes.c_ulonglong,c_ll,c_vp]
    l.jpd_diagonalize_batched_f64.argtypes=[c_vp,c_vp,c_vp,c_vp,c_int,c_ll,c_int,c_int,c_int,c_int,c_vp]
    m=torch.empty((n,n,B),dtype=torch.float64,device='cuda'); ev=torch.empty((n,B),dtype=torch.float64,device='cuda'); vec=torch.empty((n,n,B),dtype=torch.float64,device='cuda'); it=torch.empty(B,dtype=torch.int32,device='cuda')
    st=torch.cuda.current_stream().cuda_stream
    assert l.jpd_fill_synthetic_f64(m.data_ptr(),n,B,1,kind,4321,0,st)==0
    f=lambda: l.jpd_diagonalize_batched_f64(m.data_ptr(),ev.data_ptr(),vec.data_ptr(),it.data_ptr(),n,B,1,1,1,50,st)
    assert f()==0; torch.cuda.synchronize()
    best=1e9
    for _ in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); [f() for _ in range(5)]; e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1)/5)
    return best, it.float().mean().item()
for path in libs:
    r4=run(path,4,1<<24,1); r3=run(path,3,1<<23,0)
    print(f"{path.split('/')[-1]:28s} n4: {r4[0]:.3f} ms ({(1<<24)/r4[0]*1e3:.3e} mat/s, {(1<<24)*244/r4[0]/1e6/6553*100:.1f}% hbm) it={r4[1]:.2f} | n3: {r3[0]:.3f} ms ({(1<<23)/r3[0]*1e3:.3e}, {(1<<23)*148/r3[0]/1e6/6553*100:.1f}% hbm) it={r3[1]:.2f}", flush=True)
