# dev tool: PCIe ceiling (pinned, concurrent H2D + D2H) vs the e2e arm at several chunk sizes
import sys, os, time, subprocess; sys.path.insert(0,'.')
import torch
n=4; B=1<<24
h_in=torch.empty((B,n,n),dtype=torch.float64).pin_memory(); h_out=torch.empty((B,21),dtype=torch.float64).pin_memory()
d_in=torch.empty_like(h_in,device='cuda'); d_out=torch.empty((B,21),dtype=torch.float64,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for mode in ("h2d","d2h","both"):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(3):
        if mode in ("h2d","both"):
            with torch.cuda.stream(s1): d_in.copy_(h_in,non_blocking=True)
        if mode in ("d2h","both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
    print(mode, f"{dt*1e3:.1f} ms", f"h2d {h_in.numel()*8/dt/1e9:.1f} GB/s" if mode!="d2h" else "", f"d2h {h_out.numel()*8/dt/1e9:.1f} GB/s" if mode!="h2d" else "", f"-> {B/dt:.3e} mat/s bound" if mode=="both" else "", flush=True)
