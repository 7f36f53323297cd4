# dev tool: time variant builds of libjpd.so (raw ctypes) on tier-3 sizes (and n = 512: tier 5's 64x64 pivot solves)
import sys, ctypes, torch
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
cases = ((128, 2048, 'f64'), (112, 2048, 'f64'), (100, 2048, 'f64'), (64, 8192, 'f64'), (56, 8192, 'f64'), (48, 8192, 'f64'),
         (128, 2048, 'f32'), (96, 2048, 'f32'), (80, 4096, 'f32'), (64, 8192, 'f32'), (40, 8192, 'f32'), (512, 16, 'f64'))
for path in sys.argv[1:]:
    l = ctypes.CDLL(path)
    for n, B, suf in cases:
        dt = torch.float64 if suf == 'f64' else torch.float32
        fill = getattr(l, 'jpd_fill_synthetic_' + suf); run = getattr(l, 'jpd_diagonalize_batched_' + suf)
        fill.argtypes = [c_vp, c_int, c_ll, c_int, c_int, ctypes.c_ulonglong, c_ll, c_vp]
        run.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp]
        if n <= 256:
            m = torch.empty((B, n, n), dtype=dt, device='cuda'); st = torch.cuda.current_stream().cuda_stream
            assert fill(m.data_ptr(), n, B, 0, 0, 1234, 0, st) == 0
        else:
            g = torch.Generator(device='cuda'); g.manual_seed(5)
            a = torch.rand((B, n, n), dtype=dt, device='cuda', generator=g) * 2 - 1; m = (a + a.transpose(1, 2)).contiguous(); st = torch.cuda.current_stream().cuda_stream
        ev = torch.empty((B, n), dtype=dt, device='cuda'); vec = torch.empty((B, n, n), dtype=dt, device='cuda'); it = torch.empty(B, dtype=torch.int32, device='cuda')
        f = lambda: run(m.data_ptr(), ev.data_ptr(), vec.data_ptr(), it.data_ptr(), n, B, 0, 1, 1, 50, st)
        assert f() == 0; torch.cuda.synchronize()
        best = 1e9
        for _ in range(2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        A = torch.triu(m[0]) + torch.triu(m[0], 1).T
        res = (vec[0] @ A - ev[0][:, None] * vec[0]).abs().max().item()
        print(f"{path.split('/')[-1]:18s} n={n:4d} {suf} B={B}: {best:8.2f} ms  iters {it.float().mean().item():.0f} failed {(it==0).sum().item()} resid0 {res:.1e}", flush=True)
