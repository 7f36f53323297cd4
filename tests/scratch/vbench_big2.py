# dev tool: sweep JPD_BLOCK_INNER_SWEEPS over sizes: vbench_big2.py "n:B,n:B" "s1,s2"
import sys, os, subprocess
cases=[tuple(map(int,c.split(':'))) for c in sys.argv[1].split(',')]
for s in sys.argv[2].split(','):
    for n,B in cases:
        env=dict(os.environ, JPD_BLOCK_INNER_SWEEPS=s)
        r=subprocess.run([sys.executable,'tests/scratch/vbench_big.py',str(n),str(B)],env=env,capture_output=True,text=True,timeout=300)
        print(f"inner={s}:", (r.stdout.strip() or r.stderr[-300:]), flush=True)
