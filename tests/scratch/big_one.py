import sys; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
n=int(sys.argv[1]); B=int(sys.argv[2])
g=torch.Generator(device='cuda'); g.manual_seed(5)
m=torch.rand((B,n,n),dtype=torch.float64,device='cuda',generator=g)*2-1
m=(torch.triu(m)+torch.triu(m,1).transpose(1,2)).contiguous()
out=jpd.diagonalize_batched(m); torch.cuda.synchronize(); print("done", out[2][:4].tolist())
