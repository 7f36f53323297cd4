# dev tool: instrumented copy of csrc (clock64 per phase of the tier-4 block-ordering kernel) -> tests/scratch/variants/lib_<name>.so
import os, shutil, subprocess, sys
root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..', '..'))
name = sys.argv[1]
src = f'{root}/tests/scratch/variants/src_{name}'
shutil.rmtree(src, ignore_errors=True)
shutil.copytree(f'{root}/jacobi_pd_b200/csrc', src, ignore=shutil.ignore_patterns('build', 'ptxas.log'))
p = f'{src}/jpd_cluster4.cuh'; s = open(p).read()
def rep(old, new):
    global s
    assert old in s, old
    s = s.replace(old, new)
rep('''// mbarrier / bulk-copy helpers (PTX ISA: mbarrier, cp.async.bulk)''', '''__device__ long long g_cl4_prof[4 * 16 * 8 + 16];
// mbarrier / bulk-copy helpers (PTX ISA: mbarrier, cp.async.bulk)''')
rep('''__device__ __forceinline__ void cl4_step(const Cl4Ctx<T>& cx, T (&e)[EVECS ? 32 : 1], unsigned& phase) {
  cl4_front<T, EVECS, MODE>(cx);
  cl4_exchange_pivots<T>(cx, phase);
  cl4_back<T, EVECS, MODE>(cx, e);
  cluster_barrier();
}''', '''__device__ __forceinline__ void cl4_step(const Cl4Ctx<T>& cx, T (&e)[EVECS ? 32 : 1], unsigned& phase, unsigned long long* acc) {
  const bool rec = (threadIdx.x & 31) == 0 && MODE != 2;
  unsigned long long* a = acc + (threadIdx.x >> 5) * 8 + 4 * (MODE & 1);
  long long t = clock64(), u;
  cl4_front<T, EVECS, MODE>(cx);
  u = clock64(); if (rec) a[0] += u - t; t = u;
  cl4_exchange_pivots<T>(cx, phase);
  u = clock64(); if (rec) a[1] += u - t; t = u;
  cl4_back<T, EVECS, MODE>(cx, e);
  u = clock64(); if (rec) a[2] += u - t; t = u;
  cluster_barrier();
  u = clock64(); if (rec) a[3] += u - t;
}''')
s = s.replace('(cx, e, phase);', '(cx, e, phase, acc);')
rep('''    bool need = false;
    for (;;) {''', '''    bool need = false;
    __shared__ unsigned long long acc[16 * 8];
    if (tid < 128) acc[tid] = 0;
    for (;;) {''')
rep('''    const bool converged = !need && max_num_sweeps > 0;''', '''    if (blockIdx.x < 4 && b == 0 && tid < 128) g_cl4_prof[rank * 128 + tid] = (long long)acc[tid];
    const bool converged = !need && max_num_sweeps > 0;''')
for a in sys.argv[2:]:           # extra textual patches: old=>new files
    old, new = open(a).read().split('\n=>\n')
    rep(old.rstrip('\n'), new.rstrip('\n'))
open(p, 'w').write(s)
p = f'{src}/jpd_tier4.cu'; t = open(p).read()
t += '''
extern "C" __attribute__((visibility("default"))) int jpd_debug_cl4_prof(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, jpd::g_cl4_prof, sizeof(long long) * (4 * 16 * 8 + 16));
}
'''
open(p, 'w').write(t)
r = subprocess.run(['make', '-C', src, f'OUT={root}/tests/scratch/variants/lib_{name}.so', f'OBJ={root}/tests/scratch/variants/obj_{name}', f'EXTRA=-I{root}/include'], capture_output=True, text=True)
print(r.stdout[-300:], r.stderr[-2000:] if r.returncode else '')
