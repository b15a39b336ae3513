"""Size-independent properties at BASELINE.json's full sizes (no oracle needed, so they run where the oracle would take
minutes: the refined mesh, a million candidate controls).

Scaling symmetry: the deviation w = u - g0 is LINEAR in the boundary input, and every clause margin is
+-(w - (p - g0)) reduced by min / max, which are positively homogeneous.  Multiplying the input and the profile
deviation (slope m and b - g0) by alpha therefore multiplies every clause and the formula robustness by alpha; with
alpha = 2 all floating-point operations scale exactly, so for waves (g0 = 0, b' = 2 b exact) the results must be
BITWISE doubled.  For heat b' = g0 + 2 (b - g0) rounds once, hence a 1e-12 tolerance there.

Shift: adding delta to the intercept of a ">" / "<" clause moves that clause's robustness by -+delta.
"""
import numpy as np
import pytest

from pde_controller_b200 import abi, native, synth

pytestmark = pytest.mark.gpu


def doubled(b):
    d = type(b)(b.n, b.kmax, b.pmax, b.tmax, uniform_knots=b.uniform_knots)
    d.slab[:] = b.slab
    d.knot_y[:] = 2.0 * b.knot_y
    d.pred_m[:] = 2.0 * b.pred_m
    live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]
    d.pred_b[:] = np.where(live, b.g0[None, :] + 2.0 * (b.pred_b - b.g0[None, :]), 0.0)
    return d


@pytest.mark.parametrize("wl,n", [("heat", 65536), ("wave", 65536), ("heat_refined", 2048)])
def test_scaling_symmetry_full_size(wl, n):
    b = synth.make_batch(n, wl, seed=0, uniform_knots=True)
    r1 = native.evaluate_batch(b, want_clauses=True)
    r2 = native.evaluate_batch(doubled(b), want_clauses=True)
    assert (r1["status"] == 0).all() and (r2["status"] == 0).all()
    if wl == "wave":
        assert np.array_equal(r2["robustness"], 2.0 * r1["robustness"])           # g0 = 0: exact in binary arithmetic
        live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]                      # unused clause slots hold -100
        assert np.array_equal(r2["clause"][live], 2.0 * r1["clause"][live])
    else:
        scale = np.maximum(np.abs(r1["robustness"]), 1e-3 * (np.abs(b.g0) + 1.0))
        assert (np.abs(r2["robustness"] - 2.0 * r1["robustness"]) <= 1e-10 * scale).all()


def test_intercept_shift_moves_inequality_clauses_exactly_by_delta():
    b = synth.make_batch(65536, "heat", seed=1)
    delta = 0.5
    r1 = native.evaluate_batch(b, want_clauses=True)
    c = type(b)(b.n, b.kmax, b.pmax, b.tmax)
    c.slab[:] = b.slab
    live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]
    c.pred_b[:] = np.where(live, b.pred_b + delta, 0.0)
    r2 = native.evaluate_batch(c, want_clauses=True)
    for j in range(b.pmax):
        m = live[j] & (r1["status"] == 0)
        gt = m & (b.pred_rel[j] == abi.REL_CODE[">"])
        lt = m & (b.pred_rel[j] == abi.REL_CODE["<"])
        assert np.abs(r2["clause"][j][gt] - (r1["clause"][j][gt] - delta)).max() <= 1e-9
        assert np.abs(r2["clause"][j][lt] - (r1["clause"][j][lt] + delta)).max() <= 1e-9


def test_controls_scaling_symmetry_one_million_candidates():
    spec = synth.make_batch(8, "wave", seed=0)
    s, m = 3, 1 << 20
    nk = int(spec.n_knots[s])
    knot_t = np.linspace(0.0, float(spec.T[s]), nk)
    y = np.random.Generator(np.random.PCG64(2)).uniform(-5e3, 5e3, size=(nk, m))
    ev = native.ControlsEvaluator(spec, knot_t, capacity=m, sample=s)
    r1 = ev.eval(y)
    ev.close()
    spec2 = synth.make_batch(8, "wave", seed=0)
    spec2.pred_m[:] *= 2.0
    spec2.pred_b[:] *= 2.0
    ev = native.ControlsEvaluator(spec2, knot_t, capacity=m, sample=s)
    r2 = ev.eval(2.0 * y)
    ev.close()
    assert (r1["status"] == 0).all()
    assert np.array_equal(r2["robustness"], 2.0 * r1["robustness"])
    assert r2["best_index"] == r1["best_index"] == int(np.argmax(r1["robustness"]))
