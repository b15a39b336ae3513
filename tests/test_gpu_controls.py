"""Controls family (one specification, many candidate controls) against the oracle, through the C ABI."""
import numpy as np
import pytest

from parity import TOL_FP64, rel_err
from pde_controller_b200 import ProblemBatch, default_conventions, native, synth

pytestmark = pytest.mark.gpu


replicate = synth.replicate_spec


def candidates(rng, wl, nk, m):
    lo, hi = (0.0, 1e6) if wl == "heat" else (-5e3, 5e3)
    return rng.uniform(lo, hi, size=(nk, m))


@pytest.mark.parametrize("wl", ["heat", "wave"])
@pytest.mark.parametrize("kw", [{}, dict(region_rule=1, window_rule=1, steps_rule=1), dict(eq_rule=1, eq_eps=0.25, mass_heat=1, mass_wave=0)])
def test_controls_match_oracle(oracle_lib, wl, kw):
    conv = default_conventions(**kw)
    spec = synth.make_batch(24, wl, seed=31)
    spec.pred_xhi[0, :4] = spec.pred_xlo[0, :4]          # intervals without a node: interpolated end points
    rng = np.random.Generator(np.random.PCG64(5))
    for s in range(spec.n):
        T = float(spec.T[s])
        nk = 6 if wl == "wave" else int(round(T)) + 1
        knot_t = np.linspace(0.0, T, nk)
        y = candidates(rng, wl, nk, 192)
        ev = native.ControlsEvaluator(spec, knot_t, capacity=256, sample=s, conv=conv)
        got = ev.eval(y, want_clauses=True)
        ev.close()
        rep = replicate(spec, s, knot_t, y)
        ref, ref_st, ref_cl = oracle_lib.eval_batch(rep, conv, want_clauses=True)
        assert (got["status"] == ref_st).all()
        assert rel_err(got["robustness"], ref, rep).max() <= TOL_FP64
        ok = ref_st == 0
        assert (got["robustness"][~ok] == -100.0).all()
        npred = int(spec.n_pred[s])
        assert rel_err(got["clause"][:npred][:, ok].ravel(), ref_cl[:npred][:, ok].ravel(), _tile(rep, ok, npred)).max() <= TOL_FP64
        if ok.any():
            want = int(np.argmax(np.where(ok, ref, -np.inf)))
            assert got["best_index"] >= 0
            assert abs(got["best_value"] - ref[got["best_index"]]) <= TOL_FP64 * max(abs(ref[want]), 1.0) or got["best_index"] == want
            assert got["best_value"] >= ref[want] - 1e-9 * max(abs(ref[want]), 1e-3 * 300)


class _tile:
    """Field scale of the clause matrix [npred, ok samples] flattened (the parity metric needs per-entry scales)."""
    def __init__(self, rep, ok, npred):
        self.g0 = np.tile(rep.g0[ok], npred)
        self.knot_y = np.tile(rep.knot_y[:, ok], (1, npred))
        self.L = np.tile(rep.L[ok], npred)
        self.E_in = np.tile(rep.E_in[ok], npred)
        self.E_out = np.tile(rep.E_out[ok], npred)


def test_controls_equal_batch_path():
    """Same arithmetic as the batch kernels up to rounding: one spec, 4096 controls, both routes."""
    spec = synth.make_batch(4, "heat", seed=3)
    knot_t = np.linspace(0.0, float(spec.T[1]), 9)
    y = candidates(np.random.Generator(np.random.PCG64(9)), "heat", 9, 4096)
    ev = native.ControlsEvaluator(spec, knot_t, capacity=4096, sample=1)
    got = ev.eval(y)
    again = ev.eval(None, m=4096)                         # values already on the device
    ev.close()
    rep = replicate(spec, 1, knot_t, y)
    ref = native.evaluate_batch(rep)
    assert (got["status"] == ref["status"]).all()
    assert rel_err(got["robustness"], ref["robustness"], rep).max() <= TOL_FP64
    assert np.array_equal(got["robustness"], again["robustness"])
    assert got["best_index"] == int(np.argmax(got["robustness"]))


@pytest.mark.parametrize("kw", [{}, dict(newmark_beta=0.3025, newmark_gamma=0.6)])
def test_wave_controls_equal_batch_path(oracle_lib, kw):
    """Waves: the controls family steps in momentum form under the default conventions (two state vectors) and in the
    a-form otherwise; both must agree with the batch kernels (a-form) and with the oracle."""
    conv = default_conventions(**kw)
    spec = synth.make_batch(6, "wave", seed=4)
    rng = np.random.Generator(np.random.PCG64(11))
    for s in (0, 3, 5):
        knot_t = np.linspace(0.0, float(spec.T[s]), 5)
        y = candidates(rng, "wave", 5, 2048)
        ev = native.ControlsEvaluator(spec, knot_t, capacity=2048, sample=s, conv=conv)
        got = ev.eval(y)
        ev.close()
        rep = replicate(spec, s, knot_t, y)
        ref = native.evaluate_batch(rep, conv=conv)
        assert (got["status"] == ref["status"]).all()
        assert rel_err(got["robustness"], ref["robustness"], rep).max() <= TOL_FP64
        orc, orc_st = oracle_lib.eval_batch(rep, conv)
        assert (got["status"] == orc_st).all()
        assert rel_err(got["robustness"], orc, rep).max() <= TOL_FP64


def test_bad_spec_fails_every_candidate():
    spec = synth.make_batch(2, "heat", seed=3)
    spec.E_in[0] = -1.0
    ev = native.ControlsEvaluator(spec, [0.0, 1.0], capacity=16, sample=0)
    got = ev.eval(np.zeros((2, 16)))
    ev.close()
    assert (got["status"] != 0).all() and (got["robustness"] == -100.0).all() and got["best_index"] == -1


def test_search_improves_monotonically():
    from pde_controller_b200 import search
    from pde_controller_b200.problem import Predicate, Problem
    p = Problem("heat", 30, 100.0, 0.05, 6.0, 300.0, 30.0, 60.0, 3.9e-6 * 4.7e8, 4.1e-6 * 3.8e8, 9.0e5, 1.2e6,
                [Predicate(60, 90, ">", 0.0, 303.0, "G", 4.0, 6.0), Predicate(10, 40, "<", 0.0, 301.0, "G", 0.0, 6.0)])
    out = search.search_control(p, population=4096, generations=5, seed=1)
    h = out["history"]
    assert all(b >= a for a, b in zip(h, h[1:]))
    assert np.isfinite(out["robustness"]) and out["knot_y"].shape == out["knot_t"].shape
    assert h[-1] > h[0] - 1e-12
