import sys; sys.path.insert(0,'.')
import torch, jacobi_pd_b200 as jpd
l = jpd.lib(); scratch = torch.zeros(16, dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for es in (8,4):
  for thr, bps in ((256,4),(512,2),(1024,2)):
    blocks = 148*bps; iters = 20000 if es==8 else 40000
    l.jpd_fma_probe(es, blocks, thr, 100, scratch.data_ptr(), st); torch.cuda.synchronize()
    best=1e9
    for _ in range(3):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); n=l.jpd_fma_probe(es, blocks, thr, iters, scratch.data_ptr(), st); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    print(f"elem={es} threads={thr} blocks={blocks}: {n/best/1e9*1e3:.1f} GFMA/s = {2*n/best/1e9:.2f} TFLOP/s ({best:.2f} ms)")
