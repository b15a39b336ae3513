"""Independent numpy restatement of the oracle (dense matrices, np.linalg.solve every step).

TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see pde_oracle.c).  It follows the same declared
algorithm as pde_oracle.c but shares no code with it: dense element scatter instead of bands,
a dense solve per step instead of a prefactored Thomas sweep, and `K @ d` instead of element
fluxes.  It works on `Problem` objects, so it also checks the batch packing.  It is the
"closest in spirit" stand-in for femformal's numpy time loops (cpu_numpy_1core).
Reference call sites it restates: heatlinfem.heatlinfem_mix / mechlinfem.mechlinfem / logic.APCont
at utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3, one_d_wave/examples.jsonl:1-3.
"""
import math

import numpy as np

EPS = 1e-9


def step_count(T, dt, rule=0):
    q = T / dt
    if rule == 1:
        s = math.floor(q + 0.5)
    elif rule == 2:
        s = math.floor(q + EPS)
    else:
        s = math.ceil(q - EPS)
    return max(int(s), 1)


def pwl(control, t):
    if control is None:
        return 0.0
    kt, ky = control
    if len(kt) == 0:
        return 0.0
    if t <= kt[0]:
        return float(ky[0])
    if t >= kt[-1]:
        return float(ky[-1])
    j = int(np.searchsorted(kt, t, side="right")) - 1
    return float(ky[j] + (ky[j + 1] - ky[j]) * ((t - kt[j]) / (kt[j + 1] - kt[j])))


def assemble(p, conv):
    """Dense M, K over all N+1 nodes."""
    N = p.n_elem
    h = p.L / N
    lumped = conv["mass_heat"] if p.kind == "heat" else conv["mass_wave"]
    M = np.zeros((N + 1, N + 1))
    K = np.zeros((N + 1, N + 1))
    for e in range(N):
        ms = conv["material_sampling"]
        xs = e * h if ms == 1 else (e + 1) * h if ms == 2 else (e + 0.5) * h
        outer = xs < p.x_a or xs > p.x_b
        rho = p.rho_out if outer else p.rho_in
        E = p.E_out if outer else p.E_in
        ke = E / h * np.array([[1.0, -1.0], [-1.0, 1.0]])
        if lumped:
            me = rho * h / 2.0 * np.eye(2)
        else:
            me = rho * h / 6.0 * np.array([[2.0, 1.0], [1.0, 2.0]])
        K[e:e + 2, e:e + 2] += ke
        M[e:e + 2, e:e + 2] += me
    return M, K, h


def margin(rel, u, prof, conv):
    if rel == ">":
        return u - prof
    if rel == "<":
        return prof - u
    if conv["eq_rule"] == 1:
        return min(u - (prof - conv["eq_eps"]), (prof + conv["eq_eps"]) - u)
    return -abs(u - prof)


def spatial_margin(p, q, u, x, h, conv):
    N = p.n_elem
    ilo = max(int(math.ceil(q.x_lo / h - EPS)), 0)
    ihi = min(int(math.floor(q.x_hi / h + EPS)), N)
    vals = [margin(q.rel, u[i], q.m * x[i] + q.b, conv) for i in range(ilo, ihi + 1)]
    if not vals or conv["region_rule"] == 1:
        for xe in (q.x_lo, q.x_hi):
            xx = min(max(xe, 0.0), p.L)
            e = min(max(int(math.floor(xx / h)), 0), N - 1)
            xi = (xx - e * h) / h
            ux = u[e] + xi * (u[e + 1] - u[e])
            vals.append(margin(q.rel, ux, q.m * xx + q.b, conv))
    return min(vals)


V0 = dict(material_sampling=0, mass_heat=0, mass_wave=1, steps_rule=0, window_rule=0, region_rule=0,
          eq_rule=0, eq_eps=0.0, theta=0.5, newmark_beta=0.25, newmark_gamma=0.5)


def simulate(p, control=None, conv=None):
    """-> (trajectory d[S+1, N+1], velocity[S+1, N+1] or None)."""
    conv = dict(V0, **(conv or {}))
    N = p.n_elem
    M, K, h = assemble(p, conv)
    S = step_count(p.T, p.dt, conv["steps_rule"])
    dt = p.dt
    fr = slice(1, N + 1)
    Mff, Kff = M[fr, fr], K[fr, fr]
    Kf0 = K[fr, 0]
    g = p.g0

    def F(t):
        f = np.zeros(N)
        f[-1] += pwl(control, t)
        return f - Kf0 * g          # Dirichlet value moved to the right-hand side

    traj = np.zeros((S + 1, N + 1))
    traj[:, 0] = g
    if p.kind == "heat":
        th = conv["theta"]
        d = np.full(N, g, dtype=np.float64)
        v = np.linalg.solve(Mff, F(0.0) - Kff @ d)
        A = Mff + th * dt * Kff
        traj[0, 1:] = d
        for k in range(S):
            dp = d + (1 - th) * dt * v
            v = np.linalg.solve(A, F((k + 1) * dt) - Kff @ dp)
            d = dp + th * dt * v
            traj[k + 1, 1:] = d
        return traj, None
    be, ga = conv["newmark_beta"], conv["newmark_gamma"]
    d = np.zeros(N)
    v = np.zeros(N)
    a = np.linalg.solve(Mff, F(0.0) - Kff @ d)
    A = Mff + be * dt * dt * Kff
    vel = np.zeros((S + 1, N + 1))
    for k in range(S):
        dp = d + dt * v + 0.5 * (1 - 2 * be) * dt * dt * a
        vp = v + (1 - ga) * dt * a
        a = np.linalg.solve(A, F((k + 1) * dt) - Kff @ dp)
        d = dp + be * dt * dt * a
        v = vp + ga * dt * a
        traj[k + 1, 1:] = d
        vel[k + 1, 1:] = v
    return traj, vel


def robustness_of_trajectory(p, traj, conv=None):
    """-> (formula robustness, [clause robustness])."""
    conv = dict(V0, **(conv or {}))
    N = p.n_elem
    h = p.L / N
    x = np.arange(N + 1) * h
    x[N] = p.L
    S = traj.shape[0] - 1
    dt = p.dt
    clause = []
    for q in p.predicates:
        if conv["window_rule"] == 1:
            k0, k1 = int(q.t_lo / dt), int(q.t_hi / dt)
        else:
            k0, k1 = int(math.ceil(q.t_lo / dt - EPS)), int(math.floor(q.t_hi / dt + EPS))
        k0, k1 = max(k0, 0), min(k1, S)
        if k0 > k1:
            k = min(max(int(math.floor(0.5 * (q.t_lo + q.t_hi) / dt + 0.5)), 0), S)
            k0 = k1 = k
        vals = [spatial_margin(p, q, traj[k], x, h, conv) for k in range(k0, k1 + 1)]
        clause.append(min(vals) if q.top == "G" else max(vals))
    tree = p.tree if p.tree else p.default_tree()
    val = []
    for op, l, r in tree:
        if op == "leaf":
            val.append(clause[l])
        elif op == "and":
            val.append(min(val[l], val[r]))
        else:
            val.append(max(val[l], val[r]))
    return val[-1], clause


def evaluate(p, control=None, conv=None):
    traj, _ = simulate(p, control, conv)
    return robustness_of_trajectory(p, traj, conv)
