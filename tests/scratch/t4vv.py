# dev tool: time variant builds of libjpd.so (raw ctypes) on tier-4 sizes, interleaved
import sys, ctypes, torch
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
libs = [ctypes.CDLL(p) for p in sys.argv[1:]]
for suf, dt in (('f64', torch.float64), ('f32', torch.float32)):
    for n in (129, 192, 256):
        B = 264
        res = []
        for rep in range(2):
            for path, l in zip(sys.argv[1:], libs):
                fill = getattr(l, 'jpd_fill_synthetic_' + suf); run = getattr(l, 'jpd_diagonalize_batched_' + suf)
                fill.argtypes = [c_vp, c_int, c_ll, c_int, c_int, ctypes.c_ulonglong, c_ll, c_vp]
                run.argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp]
                m = torch.empty((B, n, n), dtype=dt, device='cuda'); st = torch.cuda.current_stream().cuda_stream
                assert fill(m.data_ptr(), n, B, 0, 0, 1234, 0, st) == 0
                ev = torch.empty((B, n), dtype=dt, device='cuda'); vec = torch.empty((B, n, n), dtype=dt, device='cuda'); it = torch.empty(B, dtype=torch.int32, device='cuda')
                f = lambda: run(m.data_ptr(), ev.data_ptr(), vec.data_ptr(), it.data_ptr(), n, B, 0, 1, 1, 50, st)
                assert f() == 0; torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize()
                res.append((path.split('/')[-1], e0.elapsed_time(e1), (it == 0).sum().item()))
        print(suf, n, ' '.join(f"{p}:{t:.2f}" for p, t, _ in res), 'failed', sum(x for _, _, x in res), flush=True)
