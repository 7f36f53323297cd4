import sys; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
n=int(sys.argv[1]); B=int(sys.argv[2]); layout=int(sys.argv[3]); kind=int(sys.argv[4]) if len(sys.argv)>4 else 0
dt = torch.float32 if (len(sys.argv)>5 and sys.argv[5]=='f32') else torch.float64
shape=(B,n,n) if layout==0 else (n,n,B)
m=torch.empty(shape,dtype=dt,device='cuda'); jpd.fill_synthetic(m,n,B,layout,kind,4321)
out=jpd.diagonalize_batched(m,layout=layout); torch.cuda.synchronize()
for _ in range(3): jpd.diagonalize_batched(m,layout=layout,out=out)
torch.cuda.synchronize(); print("done", out[2].float().mean().item())
