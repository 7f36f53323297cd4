# dev tool: time variant builds of libjpd.so (raw ctypes) on tier-5 sizes
import sys, ctypes, torch
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
for path in sys.argv[1:]:
    l = ctypes.CDLL(path)
    for n, B in ((512, 64), (1000, 16), (300, 64)):
        run = l.jpd_diagonalize_batched_f64
        run.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp]
        g = torch.Generator(device='cuda'); g.manual_seed(5)
        a = torch.rand((B, n, n), dtype=torch.float64, device='cuda', generator=g) * 2 - 1; m = (a + a.transpose(1, 2)).contiguous(); st = torch.cuda.current_stream().cuda_stream
        ev = torch.empty((B, n), dtype=torch.float64, device='cuda'); vec = torch.empty((B, n, n), dtype=torch.float64, device='cuda'); it = torch.empty(B, dtype=torch.int32, device='cuda')
        f = lambda: run(m.data_ptr(), ev.data_ptr(), vec.data_ptr(), it.data_ptr(), n, B, 0, 1, 1, 50, st)
        assert f() == 0; torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        A = torch.triu(m[0]) + torch.triu(m[0], 1).T
        res = (vec[0] @ A - ev[0][:, None] * vec[0]).abs().max().item()
        print(f"{path.split('/')[-1]:14s} n={n:4d} B={B}: {best:8.2f} ms  {B/best*1e3:8.1f} mat/s failed {(it==0).sum().item()} resid0 {res:.1e}", flush=True)
