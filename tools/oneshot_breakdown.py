import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from pde_controller_b200 import native, synth
n = 65536
b = synth.make_batch(n, "heat", seed=0, uniform_knots=True)
native.evaluate_batch(b)
for _ in range(3):
    t0 = time.perf_counter(); o = native.evaluate_batch(b, timing=True); dt = time.perf_counter() - t0
    tm = o["timing"]
    print(f"wall {1e3*dt:.2f} ms, total_ms(events) {tm.total_ms:.2f}, kernel {tm.kernel_ms:.2f}, h2d {tm.h2d_ms:.2f}, launches {tm.launches}")
# python-side pieces
t0 = time.perf_counter(); cb = b.as_c(); t1 = time.perf_counter(); r = np.empty(n); s = np.empty(n, np.int32); t2 = time.perf_counter()
print(f"as_c {1e3*(t1-t0):.3f} ms, empties {1e3*(t2-t1):.3f} ms")
t0 = time.perf_counter(); c = b.slab.copy(); t1 = time.perf_counter(); print(f"numpy copy of the {b.slab.nbytes/1e6:.1f} MB slab: {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); c[:] = b.slab; t1 = time.perf_counter(); print(f"second copy (pages mapped): {1e3*(t1-t0):.2f} ms")
