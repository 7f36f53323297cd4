# dev tool: time variant builds of libjpd.so (raw ctypes) at a given n, AoS fp64: vbench_n.py n B lib...
import sys, ctypes, torch
n=int(sys.argv[1]); B=int(sys.argv[2]); libs=sys.argv[3:]
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
for path in libs:
    l = ctypes.CDLL(path)
    l.jpd_fill_synthetic_f64.argtypes=[c_vp,c_int,c_ll,c_int,c_int,ctypes.c_ulonglong,c_ll,c_vp]
    l.jpd_diagonalize_batched_f64.argtypes=[c_vp,c_vp,c_vp,c_vp,c_int,c_ll,c_int,c_int,c_int,c_int,c_vp]
    m=torch.empty((B,n,n),dtype=torch.float64,device='cuda'); ev=torch.empty((B,n),dtype=torch.float64,device='cuda'); vec=torch.empty((B,n,n),dtype=torch.float64,device='cuda'); it=torch.empty(B,dtype=torch.int32,device='cuda')
    st=torch.cuda.current_stream().cuda_stream
    assert l.jpd_fill_synthetic_f64(m.data_ptr(),n,B,0,0,4321,0,st)==0
    sweeps = int(sys.argv[0] and 50)
    f=lambda: l.jpd_diagonalize_batched_f64(m.data_ptr(),ev.data_ptr(),vec.data_ptr(),it.data_ptr(),n,B,0,1,1,10,st)
    assert f()==0; torch.cuda.synchronize()
    best=1e9
    for _ in range(2):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    print(f"{path.split('/')[-1]:28s} n={n} B={B}: {best:.3f} ms (max 10 sweeps) -> {best/10/ (n+(n&1)) *1e3:.2f} us/step  it0={it[0].item()}", flush=True)
