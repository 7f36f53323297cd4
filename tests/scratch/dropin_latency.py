import subprocess, time, os
d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "oracle", "_ref")
for n in (3, 32):
    for exe in ("test_jacobi_dropin", "test_jacobi_ref"):
        t0 = time.time(); r = subprocess.run([os.path.join(d, exe), str(n), "2000", "0.1", "10", "1", "2", "5"], capture_output=True, text=True); dt = time.time() - t0
        print(f"{exe} n={n}: {dt:.3f} s for 4000 Diagonalize calls (+ generation) -> {dt/4000*1e6:.1f} us/call  [{r.stdout.strip().splitlines()[-1]}]")
