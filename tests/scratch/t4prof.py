# dev tool: per-phase clock64 breakdown of the tier-4 block-ordering kernel (instrumented variant builds)
import sys, ctypes, torch
c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p
import os
n = 256; B = 264; CALC = int(os.environ.get("CALC", "1")); F32 = int(os.environ.get("F32", "0")); DT = torch.float32 if F32 else torch.float64; SUF = "f32" if F32 else "f64"
ref = None
for path in sys.argv[1:]:
    l = ctypes.CDLL(path)
    getattr(l, "jpd_fill_synthetic_" + SUF).argtypes = [c_vp, c_int, c_ll, c_int, c_int, ctypes.c_ulonglong, c_ll, c_vp]
    getattr(l, "jpd_diagonalize_batched_" + SUF).argtypes = [c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp]
    m = torch.empty((B, n, n), dtype=DT, device='cuda'); ev = torch.empty((B, n), dtype=DT, device='cuda')
    vec = torch.empty((B, n, n), dtype=DT, device='cuda'); it = torch.empty(B, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    assert getattr(l, "jpd_fill_synthetic_" + SUF)(m.data_ptr(), n, B, 0, 0, 1234, 0, st) == 0
    f = lambda: getattr(l, "jpd_diagonalize_batched_" + SUF)(m.data_ptr(), ev.data_ptr(), vec.data_ptr(), it.data_ptr(), n, B, 0, 1, CALC, 50, st)
    assert f() == 0; torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    if ref is None: ref = ev.clone()
    err = (ev - ref).abs().max().item()
    A = m[0]; A = torch.triu(A) + torch.triu(A, 1).T
    res = (vec[0] @ A - ev[0][:, None] * vec[0]).abs().max().item()
    print(f"{path.split('/')[-1]}: {e0.elapsed_time(e1):.2f} ms  iters {it.float().mean().item():.0f} failed {(it==0).sum().item()} max|ev-ref| {err:.2e} resid0 {res:.2e}", flush=True)
    if hasattr(l, 'jpd_debug_cl4_prof'):
        buf = (ctypes.c_longlong * (4 * 16 * 8 + 16))()
        l.jpd_debug_cl4_prof(buf)
        steps = 64 * 10
        print("  rank warp | A: front exch [back bar2 | or tiles evecs] B: ...   (cycles per block step, 10 sweeps assumed)")
        for r in range(4):
            for w in (0, 1, 2, 5, 15):
                a = [buf[(r * 16 + w) * 8 + k] / steps for k in range(8)]
                print(f"  {r} {w:2d} | " + " ".join(f"{x:6.0f}" for x in a[:4]) + " | " + " ".join(f"{x:6.0f}" for x in a[4:]))
        print("  diag(A) thread 0 of cluster 0..: loads, stage1, scale1, stage2, scale2 (cycles per step; all clusters add up: divide by clusters) ", [round(buf[512 + k] / max(buf[517], 1)) for k in range(5)], buf[517])
