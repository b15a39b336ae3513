"""Analytical known-answer tests that pin the oracle (the reference ships none: SURVEY.md §4).

Every expected value below comes from a closed form of the DISCRETE scheme (exact up to
rounding), not from another run of the code under test.
"""
import math

import numpy as np
import pytest

from pde_controller_b200 import Predicate, Problem, ProblemBatch, default_conventions
from oracle import pde_oracle_np as onp


def heat_problem(N=30, L=100.0, dt=0.05, T=4.0, g0=300.0, x_a=30.0, x_b=60.0, rho_in=1710.0, rho_out=1864.0,
                 E_in=1.5e6, E_out=8e5, preds=None):
    preds = preds or [Predicate(0.0, L, ">", 0.0, 0.0, "G", 0.0, T)]
    return Problem("heat", N, L, dt, T, g0, x_a, x_b, rho_in, rho_out, E_in, E_out, preds)


def wave_problem(N=20, L=1e5, dt=0.0025, T=1.0, x_a=3e4, x_b=6e4, rho_in=8.5e-6, rho_out=8e-6, E_in=1e8,
                 E_out=2e8, preds=None):
    preds = preds or [Predicate(0.0, L, ">", 0.0, 0.0, "G", 0.0, T)]
    return Problem("wave", N, L, dt, T, 0.0, x_a, x_b, rho_in, rho_out, E_in, E_out, preds)


def traj_of(oracle_lib, p, control=None, conv=None, vel=False):
    b = ProblemBatch.from_problems([p], [control])
    return oracle_lib.trajectory(b, 0, conv, want_velocity=vel)


def const_control(p, y):
    return (np.array([0.0, p.T]), np.array([y, y]))


# (i) steady state under a constant flux: u_i = g0 + q * sum_{e<i} h/E_e, exact at the nodes
def test_heat_steady_state_constant_flux(oracle_lib):
    p = heat_problem(T=2000.0, dt=5.0)
    q = 4.0e5
    conv = default_conventions(theta=1.0)   # backward Euler: monotone decay to the steady state
    traj, rob, cl, st = traj_of(oracle_lib, p, const_control(p, q), conv)
    assert st == 0
    h = p.L / p.n_elem
    E = np.array([p.E_out if ((e + .5) * h < p.x_a or (e + .5) * h > p.x_b) else p.E_in for e in range(p.n_elem)])
    expect = p.g0 + q * np.concatenate([[0.0], np.cumsum(h / E)])
    np.testing.assert_allclose(traj[-1], expect, rtol=1e-12)


# (i') the same closed form pins the material-sampling convention (C1) on interfaces that do NOT sit on nodes:
#      element e takes the material found at its midpoint / left node / right node, with the strict outer test
#      "x < x_a or x > x_b" of the reference's lambdas (one_d_heat/examples.jsonl:1)
@pytest.mark.parametrize("sampling", [0, 1, 2])
def test_heat_steady_state_pins_material_sampling(oracle_lib, sampling):
    p = heat_problem(T=3000.0, dt=5.0, x_a=31.0, x_b=58.5)
    q = 2.5e5
    conv = default_conventions(theta=1.0, material_sampling=sampling)
    traj, rob, cl, st = traj_of(oracle_lib, p, const_control(p, q), conv)
    assert st == 0
    h = p.L / p.n_elem
    xs = {0: lambda e: (e + 0.5) * h, 1: lambda e: e * h, 2: lambda e: (e + 1) * h}[sampling]
    E = np.array([p.E_out if (xs(e) < p.x_a or xs(e) > p.x_b) else p.E_in for e in range(p.n_elem)])
    expect = p.g0 + q * np.concatenate([[0.0], np.cumsum(h / E)])
    np.testing.assert_allclose(traj[-1], expect, rtol=1e-12)
    # the three conventions really differ on this rod (each moves one of the two interface elements)
    layouts = []
    for pick in (lambda e: (e + 0.5) * h, lambda e: e * h, lambda e: (e + 1) * h):
        layouts.append(tuple(p.E_out if (pick(e) < p.x_a or pick(e) > p.x_b) else p.E_in for e in range(p.n_elem)))
    assert len(set(layouts)) == 3


def _modes(N):
    th = (2 * np.arange(1, N + 1) - 1) * math.pi / (2 * N)
    i = np.arange(1, N + 1)
    return th, np.sin(np.outer(i, th))       # phi[i-1, j]


# (ii) single material, step flux: modal closed form with the theta-method amplification factor
#      rho_j = (1 - (1-th) dt lam_j) / (1 + th dt lam_j), lam_j the DISCRETE eigenvalue
@pytest.mark.parametrize("theta", [0.5, 1.0])
def test_heat_modal_amplification(oracle_lib, theta):
    N, L, rho, E, q = 30, 120.0, 2000.0, 1.0e6, 3.0e5
    p = heat_problem(N=N, L=L, T=6.0, rho_in=rho, rho_out=rho, E_in=E, E_out=E)
    traj, _, _, st = traj_of(oracle_lib, p, const_control(p, q), default_conventions(theta=theta))
    assert st == 0
    h = L / N
    th, phi = _modes(N)
    lam = 6 * E / (rho * h * h) * (1 - np.cos(th)) / (2 + np.cos(th))
    # consistent mass on the free dofs
    M = np.zeros((N, N))
    for i in range(N):
        M[i, i] = rho * h * (2 / 3 if i < N - 1 else 1 / 3)
        if i + 1 < N:
            M[i, i + 1] = M[i + 1, i] = rho * h / 6
    mj = np.einsum("ij,ik,kj->j", phi, M, phi)
    css = q * phi[-1] / (mj * lam)
    amp = (1 - (1 - theta) * p.dt * lam) / (1 + theta * p.dt * lam)
    S = traj.shape[0] - 1
    for k in (0, 1, 7, S // 2, S):
        expect = p.g0 + phi @ (css * (1 - amp ** k))
        np.testing.assert_allclose(traj[k, 1:], expect, rtol=0, atol=1e-9 * np.abs(expect).max())


# (iii) Newmark beta=1/4, gamma=1/2, lumped mass, step force: each mode follows
#       c_k = c_ss (1 - cos(k phi_j)),  tan(phi_j / 2) = w_j dt / 2   (period elongation formula)
def test_wave_modal_period_elongation(oracle_lib):
    N, L, rho, E, F = 20, 1.0e5, 8e-6, 2e8, 3000.0
    p = wave_problem(N=N, L=L, T=1.0, rho_in=rho, rho_out=rho, E_in=E, E_out=E)
    traj, _, _, st = traj_of(oracle_lib, p, const_control(p, F))
    assert st == 0
    h = L / N
    th, phi = _modes(N)
    lam = 2 * E / (rho * h * h) * (1 - np.cos(th))
    w = np.sqrt(lam)
    m = np.full(N, rho * h)
    m[-1] *= 0.5
    mj = (phi * phi * m[:, None]).sum(0)
    css = F * phi[-1] / (mj * lam)
    ang = 2 * np.arctan(w * p.dt / 2)
    S = traj.shape[0] - 1
    for k in (0, 1, 13, S // 2, S):
        expect = phi @ (css * (1 - np.cos(k * ang)))
        np.testing.assert_allclose(traj[k, 1:], expect, rtol=0, atol=1e-9 * np.abs(traj).max())


def _modal_mass_and_lambda(N, h, rho, E, lumped):
    """Discrete eigenvalues and modal masses of the Dirichlet-free rod for both mass conventions (C2)."""
    th, phi = _modes(N)
    if lumped:
        lam = 2 * E / (rho * h * h) * (1 - np.cos(th))
        m = np.full(N, rho * h)
        m[-1] *= 0.5
        mj = (phi * phi * m[:, None]).sum(0)
    else:
        lam = 6 * E / (rho * h * h) * (1 - np.cos(th)) / (2 + np.cos(th))
        M = np.zeros((N, N))
        for i in range(N):
            M[i, i] = rho * h * (2 / 3 if i < N - 1 else 1 / 3)
            if i + 1 < N:
                M[i, i + 1] = M[i + 1, i] = rho * h / 6
        mj = np.einsum("ij,ik,kj->j", phi, M, phi)
    return phi, lam, mj


# (ii') / (iii') the NON-default mass conventions, same closed forms: heat with lumped mass, waves with consistent mass
def test_heat_modal_amplification_lumped_mass(oracle_lib):
    N, L, rho, E, q = 30, 120.0, 2000.0, 1.0e6, 3.0e5
    p = heat_problem(N=N, L=L, T=6.0, rho_in=rho, rho_out=rho, E_in=E, E_out=E)
    traj, _, _, st = traj_of(oracle_lib, p, const_control(p, q), default_conventions(mass_heat=1))
    assert st == 0
    phi, lam, mj = _modal_mass_and_lambda(N, L / N, rho, E, lumped=True)
    css = q * phi[-1] / (mj * lam)
    amp = (1 - 0.5 * p.dt * lam) / (1 + 0.5 * p.dt * lam)
    S = traj.shape[0] - 1
    for k in (0, 1, 7, S // 2, S):
        expect = p.g0 + phi @ (css * (1 - amp ** k))
        np.testing.assert_allclose(traj[k, 1:], expect, rtol=0, atol=1e-9 * np.abs(expect).max())


def test_wave_modal_period_elongation_consistent_mass(oracle_lib):
    N, L, rho, E, F = 20, 1.0e5, 8e-6, 2e8, 3000.0
    p = wave_problem(N=N, L=L, T=1.0, rho_in=rho, rho_out=rho, E_in=E, E_out=E)
    traj, _, _, st = traj_of(oracle_lib, p, const_control(p, F), default_conventions(mass_wave=0))
    assert st == 0
    phi, lam, mj = _modal_mass_and_lambda(N, L / N, rho, E, lumped=False)
    css = F * phi[-1] / (mj * lam)
    ang = 2 * np.arctan(np.sqrt(lam) * p.dt / 2)
    S = traj.shape[0] - 1
    for k in (0, 1, 13, S // 2, S):
        expect = phi @ (css * (1 - np.cos(k * ang)))
        np.testing.assert_allclose(traj[k, 1:], expect, rtol=0, atol=1e-9 * np.abs(traj).max())


# (iii'') general Newmark (beta, gamma), e.g. with numerical damping: the discrete modes decouple the scheme exactly, so the
#        rod must equal the superposition of N scalar oscillators, each stepped by the three-line Newmark recurrence
@pytest.mark.parametrize("beta,gamma", [(0.3025, 0.6), (0.25, 0.5), (0.5, 0.9)])
def test_wave_general_newmark_equals_modal_superposition(oracle_lib, beta, gamma):
    N, L, rho, E = 20, 1.0e5, 8e-6, 2e8
    p = wave_problem(N=N, L=L, T=0.6, rho_in=rho, rho_out=rho, E_in=E, E_out=E)
    knots = (np.array([0.0, 0.1, 0.35, 0.6]), np.array([0.0, 2500.0, -1500.0, 500.0]))
    traj, _, _, st = traj_of(oracle_lib, p, knots, default_conventions(newmark_beta=beta, newmark_gamma=gamma))
    assert st == 0
    phi, lam, mj = _modal_mass_and_lambda(N, L / N, rho, E, lumped=True)
    dt = p.dt
    S = traj.shape[0] - 1
    f = np.interp(np.arange(S + 1) * dt, knots[0], knots[1])
    q = np.zeros(N); v = np.zeros(N)
    a = f[0] * phi[-1] / mj - lam * q
    out = [phi @ q]
    for k in range(S):
        qp = q + dt * v + (0.5 - beta) * dt * dt * a
        vp = v + (1 - gamma) * dt * a
        an = (f[k + 1] * phi[-1] / mj - lam * qp) / (1 + beta * dt * dt * lam)
        q = qp + beta * dt * dt * an
        v = vp + gamma * dt * an
        a = an
        out.append(phi @ q)
    np.testing.assert_allclose(traj[:, 1:], np.array(out), rtol=0, atol=1e-9 * np.abs(traj).max())


def _dense_two_material(p, lumped):
    """Dense M, K of the free dofs of a two-material rod (midpoint sampling), assembled here element by element."""
    N, h = p.n_elem, p.L / p.n_elem
    M = np.zeros((N + 1, N + 1)); K = np.zeros((N + 1, N + 1))
    for e in range(N):
        xm = (e + 0.5) * h
        out = xm < p.x_a or xm > p.x_b
        rho, E = (p.rho_out, p.E_out) if out else (p.rho_in, p.E_in)
        K[e:e + 2, e:e + 2] += E / h * np.array([[1.0, -1.0], [-1.0, 1.0]])
        M[e:e + 2, e:e + 2] += (rho * h / 2 * np.eye(2)) if lumped else (rho * h / 6 * np.array([[2.0, 1.0], [1.0, 2.0]]))
    return M[1:, 1:], K[1:, 1:]


# (vi) HETEROGENEOUS rods, time-dependent input: the generalised eigenvectors of (K, M) (scipy, dense) decouple both
#      schemes exactly; every mode is then a scalar recurrence.  Independent of the banded assembly and of the
#      tridiagonal solves of the code under test.
def test_two_material_heat_equals_modal_theta_method(oracle_lib):
    from scipy.linalg import eigh
    p = heat_problem(T=5.0, x_a=31.0, x_b=58.5)
    knots = (np.array([0.0, 1.0, 2.5, 5.0]), np.array([0.0, 6e5, 1e5, 4e5]))
    traj, _, _, st = traj_of(oracle_lib, p, knots)
    assert st == 0
    M, K = _dense_two_material(p, lumped=False)
    lam, V = eigh(K, M)                                   # V^T M V = I
    S = traj.shape[0] - 1
    f = np.interp(np.arange(S + 1) * p.dt, knots[0], knots[1])
    q = np.zeros(p.n_elem)
    out = [V @ q]
    for k in range(S):
        load = V[-1] * 0.5 * (f[k] + f[k + 1])            # theta = 1/2
        q = ((1 - 0.5 * p.dt * lam) * q + p.dt * load) / (1 + 0.5 * p.dt * lam)
        out.append(V @ q)
    np.testing.assert_allclose(traj[:, 1:] - p.g0, np.array(out), rtol=0, atol=1e-9 * np.abs(traj - p.g0).max())


def test_two_material_wave_equals_modal_newmark(oracle_lib):
    from scipy.linalg import eigh
    p = wave_problem(T=0.8)
    knots = (np.array([0.0, 0.2, 0.5, 0.8]), np.array([0.0, 3000.0, -2000.0, 1000.0]))
    traj, _, _, st = traj_of(oracle_lib, p, knots)
    assert st == 0
    M, K = _dense_two_material(p, lumped=True)
    lam, V = eigh(K, M)
    dt, beta, gamma = p.dt, 0.25, 0.5
    S = traj.shape[0] - 1
    f = np.interp(np.arange(S + 1) * dt, knots[0], knots[1])
    q = np.zeros(p.n_elem); v = np.zeros(p.n_elem)
    a = f[0] * V[-1] - lam * q
    out = [V @ q]
    for k in range(S):
        qp = q + dt * v + (0.5 - beta) * dt * dt * a
        vp = v + (1 - gamma) * dt * a
        an = (f[k + 1] * V[-1] - lam * qp) / (1 + beta * dt * dt * lam)
        q, v, a = qp + beta * dt * dt * an, vp + gamma * dt * an, an
        out.append(V @ q)
    np.testing.assert_allclose(traj[:, 1:], np.array(out), rtol=0, atol=1e-8 * np.abs(traj).max())


# (iii) energy is conserved exactly once the force is back to zero
def test_wave_energy_conservation(oracle_lib):
    p = wave_problem(T=1.5)
    ctrl = (np.array([0.0, 0.2, 0.4, 1.5]), np.array([0.0, 4000.0, 0.0, 0.0]))
    traj, _, _, st, vel = traj_of(oracle_lib, p, ctrl, vel=True)
    assert st == 0
    M, K, h = onp.assemble(p, onp.V0)
    k0 = int(round(0.4 / p.dt)) + 1
    en = np.array([0.5 * v @ M @ v + 0.5 * d @ K @ d for d, v in zip(traj[k0:], vel[k0:])])
    assert en.min() > 0
    assert (en.max() - en.min()) / en.mean() < 1e-10


# (iii) static limit: a slowly ramped force ends near u(L) = F * sum h/E_e
def test_wave_static_limit(oracle_lib):
    p = wave_problem(T=40.0, dt=0.01)
    F = 2000.0
    traj, _, _, st = traj_of(oracle_lib, p, (np.array([0.0, 39.0, 40.0]), np.array([0.0, F, F])))
    assert st == 0
    h = p.L / p.n_elem
    E = np.array([p.E_out if ((e + .5) * h < p.x_a or (e + .5) * h > p.x_b) else p.E_in for e in range(p.n_elem)])
    assert traj[-1, -1] == pytest.approx(F * np.sum(h / E), rel=2e-2)


# (iv) STL semantics on the oracle's own trajectory, recomputed by hand with slices
def test_stl_semantics_hand_computed(oracle_lib):
    L, T, dt = 100.0, 4.0, 0.05
    preds = [
        Predicate(10.0, 50.0, ">", 0.1, 295.0, "G", 0.5, 1.5),     # nodes 3..15
        Predicate(60.0, 100.0, "<", -0.05, 330.0, "F", 2.0, 4.0),  # nodes 18..30, window ends at the horizon
        Predicate(20.0, 40.0, "=", 0.0, 305.0, "F", 1.0, 3.0),     # nodes 6..12
    ]
    p = heat_problem(L=L, T=T, dt=dt, preds=preds)
    p.tree = [("leaf", 0, 0), ("leaf", 1, 0), ("or", 0, 1), ("leaf", 2, 0), ("and", 2, 3)]
    traj, rob, cl, st = traj_of(oracle_lib, p, (np.array([0.0, 4.0]), np.array([0.0, 8e5])))
    assert st == 0
    x = np.arange(31) * (L / 30)
    m0 = (traj[10:31, 3:16] - (0.1 * x[3:16] + 295.0)).min(axis=1).min()
    m1 = ((-0.05 * x[18:31] + 330.0) - traj[40:81, 18:31]).min(axis=1).max()
    m2 = (-np.abs(traj[20:61, 6:13] - 305.0)).min(axis=1).max()
    assert cl[0] == m0 and cl[1] == m1 and cl[2] == m2
    assert rob == min(max(m0, m1), m2)


def test_stl_window_edge_cases(oracle_lib):
    L, T, dt = 100.0, 2.0, 0.05
    preds = [
        Predicate(0.0, 100.0, ">", 0.0, 299.0, "G", 1.0, 1.0),      # degenerate window on a step: k = 20
        Predicate(0.0, 100.0, ">", 0.0, 299.0, "G", 1.06, 1.08),    # between steps: nearest step (1.07/0.05 = 21.4 -> k = 21)
        Predicate(12.0, 12.0, ">", 0.0, 299.0, "G", 0.0, 2.0),      # no node inside: interpolate at x = 12
        Predicate(0.0, 100.0, ">", 0.0, 299.0, "F", 1.95, 7.0),     # window clipped at the horizon
    ]
    p = heat_problem(L=L, T=T, dt=dt, preds=preds)
    p.tree = [("leaf", 0, 0)]
    traj, rob, cl, st = traj_of(oracle_lib, p, (np.array([0.0, 2.0]), np.array([0.0, -8e5])))
    assert st == 0
    assert cl[0] == (traj[20] - 299.0).min()
    assert cl[1] == (traj[21] - 299.0).min()
    h = L / 30
    xi = (12.0 - 3 * h) / h
    u12 = traj[:, 3] + xi * (traj[:, 4] - traj[:, 3])
    assert cl[2] == pytest.approx((u12 - 299.0).min(), rel=1e-14)
    assert cl[3] == (traj[39:41] - 299.0).min(axis=1).max()


# (iv') the NON-default STL conventions (C7 truncation of the window bounds, C8 end points always included, C9 band
#       rule for "="), recomputed by hand on the oracle's own trajectory
def test_stl_non_default_conventions_hand_computed(oracle_lib):
    L, T, dt = 100.0, 3.0, 0.05
    h = L / 30
    preds = [
        Predicate(11.0, 52.0, ">", 0.0, 299.0, "G", 1.04, 2.03),     # trunc: steps 20..40; nodes 4..15 + end points 11, 52
        Predicate(20.0, 40.0, "=", 0.0, 300.5, "F", 0.52, 1.99),      # trunc: steps 10..39; band eps: eps - |u - p|
    ]
    p = heat_problem(L=L, T=T, dt=dt, preds=preds)
    p.tree = [("leaf", 0, 0), ("leaf", 1, 0), ("and", 0, 1)]
    conv = default_conventions(window_rule=1, region_rule=1, eq_rule=1, eq_eps=0.75)
    traj, rob, cl, st = traj_of(oracle_lib, p, (np.array([0.0, 3.0]), np.array([0.0, 6e5])), conv)
    assert st == 0

    def at(xx):
        e = int(xx // h)
        xi = (xx - e * h) / h
        return traj[:, e] + xi * (traj[:, e + 1] - traj[:, e])

    nodes = traj[:, 4:16] - 299.0
    ends = np.stack([at(11.0) - 299.0, at(52.0) - 299.0], axis=1)
    m0 = np.minimum(nodes.min(axis=1), ends.min(axis=1))[20:41].min()
    inner = np.concatenate([traj[:, 6:13], np.stack([at(20.0), at(40.0)], axis=1)], axis=1)
    m1 = (0.75 - np.abs(inner - 300.5)).min(axis=1)[10:40].max()
    assert cl[0] == pytest.approx(m0, rel=1e-13) and cl[1] == pytest.approx(m1, rel=1e-13)
    assert rob == pytest.approx(min(m0, m1), rel=1e-13)


# (v) second-order convergence in h of the end-point value
def test_mesh_refinement_order_two(oracle_lib):
    errs = []
    ref = None
    for N in (1024, 16, 32, 64):
        p = heat_problem(N=N, x_a=25.0, x_b=62.5, T=4.0, dt=0.02)   # interfaces on element edges for every N
        traj, _, _, st = traj_of(oracle_lib, p, (np.array([0.0, 4.0]), np.array([0.0, 5e5])))
        assert st == 0
        if ref is None:
            ref = traj[-1, -1]
        else:
            errs.append(abs(traj[-1, -1] - ref))
    assert 3.5 < errs[0] / errs[1] < 4.5 and 3.5 < errs[1] / errs[2] < 4.5


def test_c_oracle_matches_numpy_restatement(oracle_lib):
    from pde_controller_b200 import synth
    for wl in ("heat", "wave"):
        b = synth.make_batch(24, wl, seed=3)
        rob, st, cl = oracle_lib.eval_batch(b, want_clauses=True)
        assert (st == 0).all()
        for s in range(b.n):
            p, c = b.get_problem(s)
            r, clause = onp.evaluate(p, c)
            scale = max(abs(r), 1e-3 * (abs(p.g0) + 1.0))
            assert abs(rob[s] - r) <= 1e-10 * scale
            np.testing.assert_allclose(cl[: len(clause), s], clause, rtol=0, atol=1e-10 * max(scale, np.abs(clause).max()))


@pytest.mark.parametrize("wl", ["heat", "wave"])
def test_c_oracle_matches_numpy_restatement_on_stress_shapes(oracle_lib, wl):
    """The shapes the GPU stress tests use (overlapping windows, intervals inside one element, dense non-uniform knots,
    eight clauses with and/or trees): the checker itself is checked there by its independent numpy twin."""
    from test_gpu_parity import _stress_batch, _wide_batch
    for b in (_stress_batch(wl, 20, 31), _wide_batch(wl, 20, 41)):
        rob, st, cl = oracle_lib.eval_batch(b, want_clauses=True)
        assert (st == 0).all()
        for s in range(b.n):
            p, c = b.get_problem(s)
            r, clause = onp.evaluate(p, c)
            scale = max(abs(r), 1e-3 * (abs(p.g0) + 1.0 + np.abs(c[1]).max() * p.L / min(p.E_in, p.E_out)))
            assert abs(rob[s] - r) <= 1e-9 * scale
            np.testing.assert_allclose(cl[: len(clause), s], clause, rtol=0, atol=1e-9 * max(scale, np.abs(clause).max()))


def test_conventions_change_results(oracle_lib):
    """Each declared convention flag is live (moves the result) and is honoured by both restatements."""
    from pde_controller_b200 import synth
    b = synth.make_batch(8, "heat", seed=5)
    base, _ = oracle_lib.eval_batch(b)
    for kw in (dict(mass_heat=1), dict(material_sampling=1), dict(theta=1.0), dict(region_rule=1), dict(eq_rule=1, eq_eps=0.5)):
        rob, st = oracle_lib.eval_batch(b, default_conventions(**kw))
        assert (st == 0).all()
        assert np.any(np.abs(rob - base) > 1e-9), kw
        p, c = b.get_problem(0)
        r, _ = onp.evaluate(p, c, kw)
        assert abs(rob[0] - r) <= 1e-10 * max(abs(r), 1.0), kw


def test_bad_parameters_are_flagged_not_raised(oracle_lib):
    from pde_controller_b200 import synth
    b = synth.make_batch(4, "heat", seed=1)
    b.dt[1] = 0.0
    b.n_elem[2] = 0
    b.knot_y[0, 3] = np.inf
    rob, st = oracle_lib.eval_batch(b)
    assert st[0] == 0 and st[1] == 3 and st[2] == 3 and st[3] == 2
    assert rob[1] == -100.0 and rob[2] == -100.0 and rob[3] == -100.0


def test_uniform_knots_equal_explicit_linspace(oracle_lib):
    """knot_t == NULL (uniform knots, C11) is the same input as explicit linspace(0, T, K) knots, bit for bit."""
    from pde_controller_b200 import synth
    for wl in ("heat", "wave"):
        a = synth.make_batch(96, wl, seed=4)
        b = synth.make_batch(96, wl, seed=4, uniform_knots=True)
        assert b.knot_t is None and b.slab_bytes < a.slab_bytes
        ra, sa = oracle_lib.eval_batch(a)
        rb, sb = oracle_lib.eval_batch(b)
        assert (sa == sb).all() and np.array_equal(ra, rb)
        p, c = b.get_problem(5)
        assert np.array_equal(c[0], a.knot_t[: len(c[0]), 5])

