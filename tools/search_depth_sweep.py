import os, sys, time
sys.path.insert(0, '/root/repo')
from pde_controller_b200 import search, synth
n = 8192
src = synth.make_batch(n, "heat", seed=0)
probs = [src.get_problem(i)[0] for i in range(n)]
search.search_controls(probs[:64], workers=4)
ref = None
for w in (2, 4, 8, 16):
    t0 = time.perf_counter(); out = search.search_controls(probs, workers=w); dt = time.perf_counter() - t0
    r = [o["robustness"] for o in out]
    ref = ref or r
    print(f"{os.environ.get('PDEB_LIB','default').split('/')[-1]} threads={w}: {n/dt:.0f} programs/s same={r == ref}", flush=True)
