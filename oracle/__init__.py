"""CPU oracle of the utility-evaluation path.  TEST INFRASTRUCTURE ONLY (see pde_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  PARITY UNPINNED: the reference's femformal arithmetic is not in its
snapshot; see pde_oracle.c for what the restatement follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpde_oracle.so")
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp the C restatement into oracle/libpde_oracle.so."""
    src = os.path.join(_HERE, "pde_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "pdeb200.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    # -ffp-contract=off: the oracle states its arithmetic operation by operation
    cmd = ["gcc", "-O3", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11",
           "-o", _SO, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.pdeo_eval_batch.restype = C.c_int
        _lib.pdeo_steps.restype = C.c_int64
        _lib.pdeo_trajectory.restype = C.c_int
        _lib.pdeo_max_threads.restype = C.c_int
    return _lib


def eval_batch(batch, conv=None, nthreads=0, want_clauses=False):
    """Oracle over a ProblemBatch -> (robustness f64[n], status i32[n][, clause f64[pmax,n]])."""
    L = lib()
    cb = batch.as_c()
    rob = np.empty(batch.n, dtype=np.float64)
    st = np.empty(batch.n, dtype=np.int32)
    cl = np.empty((batch.pmax, batch.stride), dtype=np.float64) if want_clauses else None
    rc = L.pdeo_eval_batch(C.byref(cb), C.byref(conv) if conv is not None else None,
                           rob.ctypes.data_as(C.c_void_p),
                           cl.ctypes.data_as(C.c_void_p) if cl is not None else None,
                           st.ctypes.data_as(C.c_void_p), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError(f"pdeo_eval_batch failed: {rc}")
    if want_clauses:
        return rob, st, cl[:, : batch.n]
    return rob, st


def steps(batch, s=0, conv=None):
    cb = batch.as_c()
    return int(lib().pdeo_steps(C.byref(cb), C.c_int64(s), C.byref(conv) if conv is not None else None))


def trajectory(batch, s=0, conv=None, want_velocity=False):
    """Full trajectory of sample s: (d[S+1, N+1], robustness, clause[np]) (+ velocity for waves)."""
    L = lib()
    cb = batch.as_c()
    S = steps(batch, s, conv)
    if S < 0:
        raise ValueError("bad sample")
    N = int(batch.n_elem[s])
    traj = np.zeros((S + 1, N + 1), dtype=np.float64)
    vel = np.zeros((S + 1, N + 1), dtype=np.float64) if want_velocity else None
    rob = C.c_double(0.0)
    cl = np.zeros(8, dtype=np.float64)
    st = L.pdeo_trajectory(C.byref(cb), C.c_int64(s), C.byref(conv) if conv is not None else None,
                           traj.ctypes.data_as(C.c_void_p),
                           vel.ctypes.data_as(C.c_void_p) if vel is not None else None,
                           C.c_int64(S + 1), C.byref(rob), cl.ctypes.data_as(C.c_void_p))
    out = (traj, rob.value, cl[: int(batch.n_pred[s])].copy(), st)
    if want_velocity:
        return out + (vel,)
    return out


def max_threads():
    return int(lib().pdeo_max_threads())
