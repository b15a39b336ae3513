"""Device-code logic on the CPU: pdeb_core.h compiled by g++ (one-lane warps) against the oracle.

This is scaffolding for a GPU-less box; the -m gpu tests run the same comparison through the
real kernels and the C ABI.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pde_controller_b200 import default_conventions, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL64 = 1e-9   # relative, floor of the scale at 1e-3 of the field magnitude (see tests/parity.py)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "host_emu", "emu.cpp")
    so = os.path.join(ROOT, "tests", "host_emu", "libemu.so")
    deps = [src, os.path.join(ROOT, "pde_controller_b200", "csrc", "pdeb_core.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src], check=True)
    return C.CDLL(so)


def run_emu(emu, b, conv, precision=64):
    cb = b.as_c()
    rob = np.zeros(b.n)
    st = np.zeros(b.n, np.int32)
    cl = np.zeros((b.pmax, b.stride))
    emu.emu_eval_batch(C.byref(cb), C.byref(conv), precision, rob.ctypes.data_as(C.c_void_p),
                       cl.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return rob, st, cl


@pytest.mark.parametrize("wl", ["heat", "wave"])
@pytest.mark.parametrize("kw", [{}, dict(theta=1.0, newmark_beta=0.3025, newmark_gamma=0.6), dict(mass_heat=1, mass_wave=0),
                                dict(region_rule=1, window_rule=1, steps_rule=1), dict(eq_rule=1, eq_eps=0.25, material_sampling=2)])
def test_device_code_matches_oracle(emu, oracle_lib, wl, kw):
    from parity import rel_err
    conv = default_conventions(**kw)
    b = synth.make_batch(384, wl, seed=21)
    # exercise the rare branches too: degenerate interval, window at the horizon, zero control
    b.pred_xhi[0, :16] = b.pred_xlo[0, :16]
    b.pred_thi[0, 16:32] = b.T[16:32]
    b.n_knots[32:48] = 0
    ref, ref_st, ref_cl = oracle_lib.eval_batch(b, conv, want_clauses=True)
    rob, st, cl = run_emu(emu, b, conv)
    assert (st == ref_st).all()
    assert rel_err(rob, ref, b).max() <= TOL64


def test_device_code_fp32_mode(emu, oracle_lib):
    from parity import rel_err
    conv = default_conventions()
    for wl, tol in (("heat", 2e-3), ("wave", 5e-2)):
        b = synth.make_batch(256, wl, seed=22)
        ref, ref_st = oracle_lib.eval_batch(b, conv)
        rob, st, _ = run_emu(emu, b, conv, 32)
        assert (st == ref_st).all()
        assert rel_err(rob, ref, b).max() <= tol


@pytest.mark.parametrize("wl", ["heat", "wave"])
@pytest.mark.parametrize("kw", [{}, dict(region_rule=1)])
def test_device_code_overlapping_windows_in_element_intervals(emu, oracle_lib, wl, kw):
    """Same stress set as the GPU test: several clauses open at once, intervals inside one element, dense knots."""
    from parity import rel_err
    from test_gpu_parity import _stress_batch
    conv = default_conventions(**kw)
    b = _stress_batch(wl, 256, 31)
    ref, ref_st, ref_cl = oracle_lib.eval_batch(b, conv, want_clauses=True)
    rob, st, cl = run_emu(emu, b, conv)
    assert (st == ref_st).all()
    assert rel_err(rob, ref, b).max() <= TOL64


@pytest.mark.parametrize("wl", ["heat", "wave"])
def test_device_code_eight_clauses(emu, oracle_lib, wl):
    from parity import rel_err
    from test_gpu_parity import _wide_batch
    conv = default_conventions()
    b = _wide_batch(wl, 192, 41)
    ref, ref_st, ref_cl = oracle_lib.eval_batch(b, conv, want_clauses=True)
    rob, st, cl = run_emu(emu, b, conv)
    assert (st == ref_st).all()
    assert rel_err(rob, ref, b).max() <= TOL64
    live = (np.arange(b.pmax)[:, None] < b.n_pred) & (ref_st == 0)
    assert np.allclose(cl[:, :b.n][live], ref_cl[:, :b.n][live], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("wl", ["heat", "wave"])
@pytest.mark.parametrize("bad", [np.inf, np.nan])
def test_device_code_flags_a_state_that_turns_nonfinite_inside_a_window(emu, oracle_lib, wl, bad):
    """inf / NaN at the LAST control knot: the trajectory is finite for most of the run and turns non-finite inside
    clause 0's window; the running maxima drop NaN margins, the sticky state test must not (advisor, round 1)."""
    conv = default_conventions()
    b = synth.make_batch(128, wl, seed=21)
    hit = np.arange(b.n) % 2 == 0
    for s in np.nonzero(hit)[0]:
        b.knot_y[b.n_knots[s] - 1, s] = bad
        b.pred_tlo[0, s], b.pred_thi[0, s] = 0.0, b.T[s]
    ref, ref_st = oracle_lib.eval_batch(b, conv)
    rob, st, _ = run_emu(emu, b, conv)
    assert (ref_st[hit] == 2).all() and (st == ref_st).all() and (rob[hit] == -100.0).all()
    from parity import rel_err
    assert rel_err(rob, ref, b)[~hit].max() <= TOL64
