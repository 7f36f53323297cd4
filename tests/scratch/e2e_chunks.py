import sys, os, time; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
n=4; B=1<<24
h_mat=torch.empty((B,n,n),dtype=torch.float64).pin_memory(); h_ev=torch.empty((B,n),dtype=torch.float64).pin_memory(); h_vec=torch.empty((B,n,n),dtype=torch.float64).pin_memory(); h_it=torch.empty((B,),dtype=torch.int32).pin_memory()
tmp=torch.empty((B,n,n),dtype=torch.float64,device='cuda'); jpd.fill_synthetic(tmp,n,B,0,1,4321); h_mat.copy_(tmp); del tmp; torch.cuda.synchronize()
for _ in range(2): jpd.diagonalize_batched_host(h_mat,h_ev,h_vec,h_it,n,B,0,1,True,50)
t0=time.perf_counter()
for _ in range(4): jpd.diagonalize_batched_host(h_mat,h_ev,h_vec,h_it,n,B,0,1,True,50)
dt=(time.perf_counter()-t0)/4
print("chunk MB", os.environ.get("JPD_HOST_CHUNK_MB","64"), f"{dt*1e3:.1f} ms  {B/dt:.3e} mat/s", flush=True)
