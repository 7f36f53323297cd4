import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, time
import jacobi_pd_b200 as jpd
from oracle import binding as ob
import parity
from gpu_util import run_gpu
for n,B,dt in ((256,2,np.float64),(129,3,np.float64),(200,5,np.float64),(255,2,np.float64),(130,2,np.float32),(256,40,np.float64)):
    A = ob.fill_synthetic(n, B, seed=9000+n, dtype=dt)
    t0=time.time(); ref = ob.oracle_diagonalize(A, 1); t1=time.time()
    out = run_gpu(A, 1)
    print(n,B,dt.__name__, "tier", jpd.lib().jpd_kernel_tier(n,8,1), parity.full_check(A, out, ref, 1), "iters", out[2][:3], "ref iters", ref[2][:3], f"cpu {t1-t0:.1f}s", flush=True)
    out2 = run_gpu(A, 2, calc=False)
    parity.check_evals(out2[0], ob.oracle_diagonalize(A,2,False)[0])
    out3 = run_gpu(A, 1, layout=1)
    for x,y in zip(out,out3): assert np.array_equal(x,y)
print("cluster ok")
