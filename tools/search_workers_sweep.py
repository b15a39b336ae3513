import sys, time
sys.path.insert(0, '/root/repo')
from pde_controller_b200 import search, synth
n = 2048
src = synth.make_batch(n, "heat", seed=0)
probs = [src.get_problem(i)[0] for i in range(n)]
search.search_controls(probs[:64], workers=4)
for w in (4, 8, 16, 32, 64):
    t0 = time.perf_counter(); out = search.search_controls(probs, workers=w); dt = time.perf_counter() - t0
    print(f"workers={w}: {n/dt:.0f} programs/s, kernel_ms/program {sum(o['kernel_ms'] for o in out)/n:.2f}")
srcw = synth.make_batch(512, "wave", seed=0)
pw = [srcw.get_problem(i)[0] for i in range(512)]
t0 = time.perf_counter(); out = search.search_controls(pw, workers=16); dt = time.perf_counter() - t0
print(f"wave workers=16: {512/dt:.0f} programs/s")
