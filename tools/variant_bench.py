"""Time alternative builds of libpdeb200.so against each other on the GPU box (development tool).

  python tools/variant_bench.py build  NAME=DEFINE[,DEFINE...] ...   # here (nvcc cross-compiles)
  python tools/variant_bench.py run [--workloads heat,wave] [--batch 65536]   # under gpurun

`run` loads every pde_controller_b200/variants/lib*.so in a fresh process, compares its results on a
4096-sample batch with those of the default library (the first one run; parity against the oracle is the job
of tests/, the only place besides bench.py's CPU leg that may execute oracle/), then times the resident eval
with CUDA events.
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "pde_controller_b200", "variants")


def do_build(specs):
    from pde_controller_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, defs = spec.partition("=")
        defines = [d for d in defs.split(",") if d]
        out = os.path.join(VDIR, f"lib{name}.so")
        build.build(force=True, defines=defines, out=out)
        print("built", out, defines)


def one(workloads, batch, precision):
    import numpy as np
    from pde_controller_b200 import native, synth
    res = {}
    for wl in workloads:
        small = synth.make_batch(4096, wl, seed=5)
        got = native.evaluate_batch(small, precision=precision)
        # reference results: those of the first library run (the default build), kept in a scratch file
        ref_path = os.path.join(os.environ.get("PDEB_VARIANT_REF_DIR", "/tmp"), f"variant_ref_{wl}_{precision}.npz")
        if os.path.exists(ref_path):
            z = np.load(ref_path)
            ref, st = z["robustness"], z["status"]
        else:
            ref, st = got["robustness"].copy(), got["status"].copy()
            np.savez(ref_path, robustness=ref, status=st)
        scale = np.maximum(np.abs(ref), 1e-3 * (np.abs(small.g0) + 1 + np.abs(small.knot_y).max(axis=0) * small.L / np.minimum(small.E_in, small.E_out)))
        err = float((np.abs(got["robustness"] - ref) / scale).max())
        b = synth.make_batch(batch, wl, seed=0)
        r = native.ResidentBatch(b)
        for _ in range(3):
            r.eval(precision, timing=True)
        ms = sorted(r.eval(precision, timing=True).kernel_ms for _ in range(15))
        res[wl] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "evals_per_s": batch / (ms[len(ms) // 2] * 1e-3),
                   "max_rel_err": err, "status_ok": bool((got["status"] == st).all())}
        r.close()
    print(json.dumps(res))


def decomp(workloads, batch, precision):
    """Cost split of the fused kernel on uniform synthetic inputs (all trajectories the same length):
    solve only / + piecewise-linear input / + one clause pass per step / + two."""
    import numpy as np
    from pde_controller_b200 import abi, native, synth
    res = {}
    for wl in workloads:
        out = {}
        for case in ("solve", "solve+pwl", "solve+pwl+1clause", "solve+pwl+2clauses"):
            b = synth.make_batch(batch, wl, seed=0)
            T = 10.0 if wl == "heat" else 1.5
            b.T[:] = T
            nk = int(b.n_knots.max())
            b.n_knots[:] = 1 if case == "solve" else nk
            b.knot_t[:nk, :] = (np.arange(nk)[:, None] * (T / (nk - 1)))
            npred = 2 if case.endswith("2clauses") else 1
            b.n_pred[:] = npred
            b.pred_tlo[:, :] = T if "clause" not in case else 0.0
            b.pred_thi[:, :] = T
            b.n_tree[:] = 2 * npred - 1
            b.tree_op[0, :] = abi.TREE_LEAF; b.tree_lhs[0, :] = 0
            if npred == 2:
                b.tree_op[1, :] = abi.TREE_LEAF; b.tree_lhs[1, :] = 1
                b.tree_op[2, :] = abi.TREE_AND; b.tree_lhs[2, :] = 0; b.tree_rhs[2, :] = 1
            r = native.ResidentBatch(b)
            for _ in range(3):
                r.eval(precision, timing=True)
            ms = sorted(r.eval(precision, timing=True).kernel_ms for _ in range(9))
            steps = T / float(b.dt[0])
            # SMSP-cycles per warp-step at 1965 MHz, 592 SMSPs
            out[case] = {"ms": round(ms[len(ms) // 2], 4),
                         "smsp_cycles_per_warp_step": round(ms[len(ms) // 2] * 1e-3 * 1.965e9 * 592 / (batch / 32 * steps), 1)}
            r.close()
        res[wl] = out
    print(json.dumps(res))


def controls(workloads, m, precision):
    """Controls family: one specification, m candidate controls; kernel time from the library's CUDA events."""
    import numpy as np
    from pde_controller_b200 import native, synth
    res = {}
    for wl in workloads:
        spec = synth.make_batch(8, wl, seed=0)
        s = 3
        T = float(spec.T[s])
        nk = int(spec.n_knots[s])
        knot_t = np.linspace(0.0, T, nk)
        lo, hi = (0.0, 1e6) if wl == "heat" else (-5e3, 5e3)
        y = np.random.Generator(np.random.PCG64(0)).uniform(lo, hi, size=(nk, m))
        ev = native.ControlsEvaluator(spec, knot_t, capacity=m, sample=s)
        ev.eval(y, want_robustness=False)
        ms = sorted(ev.eval(None, m=m, want_robustness=False)["kernel_ms"] for _ in range(7))
        ev.close()
        steps = T / float(spec.dt[s])
        res[wl] = {"m": m, "ms": round(ms[len(ms) // 2], 4), "evals_per_s": m / (ms[len(ms) // 2] * 1e-3), "steps": steps,
                   "n_pred": int(spec.n_pred[s])}
    print(json.dumps(res))


def do_run(argv):
    workloads, batch, precision, mode = "heat,wave", 65536, 64, "one"
    for i, a in enumerate(argv):
        if a == "--decomp":
            mode = "decomp"
        if a == "--controls":
            mode = "controls"
        if a == "--workloads":
            workloads = argv[i + 1]
        if a == "--batch":
            batch = int(argv[i + 1])
        if a == "--precision":
            precision = int(argv[i + 1])
    libs = [os.path.join(ROOT, "pde_controller_b200", "libpdeb200.so")] + sorted(glob.glob(os.path.join(VDIR, "lib*.so")))
    import tempfile
    ref_dir = tempfile.mkdtemp(prefix="pdeb_variant_ref_")
    for lib in libs:
        env = dict(os.environ, PDEB_LIB=lib, PDEB_VARIANT_REF_DIR=ref_dir)
        p = subprocess.run([sys.executable, __file__, mode, workloads, str(batch), str(precision)], env=env,
                           capture_output=True, text=True, timeout=600)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:]
        print(os.path.basename(lib), line, flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        do_build(sys.argv[2:])
    elif sys.argv[1] == "run":
        do_run(sys.argv[2:])
    elif sys.argv[1] == "controls":
        controls(sys.argv[2].split(","), int(sys.argv[3]), int(sys.argv[4]))
    elif sys.argv[1] == "decomp":
        decomp(sys.argv[2].split(","), int(sys.argv[3]), int(sys.argv[4]))
    elif sys.argv[1] == "one":
        one(sys.argv[2].split(","), int(sys.argv[3]), int(sys.argv[4]))
