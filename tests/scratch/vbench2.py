# dev tool: time exp_run() of experimental kernels against the product library
import sys, ctypes, torch, numpy as np
sys.path.insert(0, '.')
import jacobi_pd_b200 as jpd
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
for path in sys.argv[1:]:
    l = ctypes.CDLL(path); l.exp_run.argtypes=[c_vp,c_vp,c_vp,c_vp,c_int,c_ll,c_vp]
    res=[]
    for n,B,kind in ((4,1<<24,1),(3,1<<23,0)):
        m=torch.empty((n,n,B),dtype=torch.float64,device='cuda'); jpd.fill_synthetic(m,n,B,1,kind,4321)
        ev=torch.empty((n,B),dtype=torch.float64,device='cuda'); vec=torch.empty((n,n,B),dtype=torch.float64,device='cuda'); it=torch.empty(B,dtype=torch.int32,device='cuda')
        st=torch.cuda.current_stream().cuda_stream
        f=lambda: l.exp_run(m.data_ptr(),ev.data_ptr(),vec.data_ptr(),it.data_ptr(),n,B,st)
        assert f()==0; torch.cuda.synchronize()
        ref=jpd.diagonalize_batched(m,layout=1); torch.cuda.synchronize()
        ok = torch.equal(ref[0],ev) and torch.equal(ref[1],vec) and torch.equal(ref[2],it)
        best=1e9
        for _ in range(3):
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record(); [f() for _ in range(5)]; e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1)/5)
        by=jpd.algorithmic_bytes(n,8)
        res.append(f"n{n}: {best:.3f} ms {B*by/best/1e6/6553*100:.1f}% hbm same_as_product={ok}")
        del m,ev,vec,it,ref
    print(path.split('/')[-1], ' | '.join(res), flush=True)
