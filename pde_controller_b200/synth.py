"""Seeded synthetic workload generator.

The reference's 2M-sample corpus is hosted off-tree (README.md:13, :34 of the reference);
the ranges here bracket the 14 in-tree records
(utils/few_shot_prompts/examples/*/examples.jsonl).  Samples are generated in blocks of
BLOCK global indices, each block from its own PCG64 stream keyed by (seed, workload, block),
so any shard [start, start+n) of a workload is reproducible independently of the others.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .problem import ProblemBatch

BLOCK = 4096
KMAX = {"heat": 15, "wave": 5, "heat_refined": 11, "mixed": 15}
PMAX = 3
WORKLOADS = ("heat", "wave", "heat_refined", "mixed")
_WID = {"heat": 1, "wave": 2, "heat_refined": 3}


def _gen_block(workload: str, seed: int, block: int):
    """All fields of one block of BLOCK samples as a dict of arrays ([plane, sample] for lists)."""
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, _WID[workload], block])))
    n = BLOCK
    heat = workload != "wave"
    f = {}
    if heat:
        f["kind"] = np.full(n, abi.KIND_HEAT, np.uint8)
        L = rng.integers(80, 301, n).astype(np.float64)
        g0 = rng.integers(250, 351, n).astype(np.float64)
        if workload == "heat_refined":
            f["n_elem"] = np.full(n, 1024, np.int32)
            T = np.full(n, 10.0)
            dt = T / 10000.0
        else:
            f["n_elem"] = np.full(n, 30, np.int32)
            T = np.round(rng.uniform(5.0, 14.0, n), 4)
            dt = np.full(n, 0.05)
        rho_in = rng.uniform(3e-6, 6e-6, n) * rng.uniform(3e8, 4.8e8, n)
        rho_out = rng.uniform(3e-6, 6e-6, n) * rng.uniform(3e8, 4.8e8, n)
        E_in = rng.uniform(4e5, 1.7e6, n)
        E_out = rng.uniform(4e5, 1.7e6, n)
    else:
        f["kind"] = np.full(n, abi.KIND_WAVE, np.uint8)
        f["n_elem"] = np.full(n, 20, np.int32)
        L = rng.integers(50000, 100001, n).astype(np.float64)
        g0 = np.zeros(n)
        T = np.round(rng.uniform(0.9, 2.0, n), 4)
        dt = np.full(n, 0.0025)
        rho_in = rng.uniform(7.6e-6, 8.5e-6, n)
        rho_out = rng.uniform(7.6e-6, 8.5e-6, n)
        E_in = rng.uniform(1e8, 2.3e8, n)
        E_out = rng.uniform(1e8, 2.3e8, n)
    xa = np.floor(rng.uniform(0, 1, n) * L)
    xb = np.floor(rng.uniform(0, 1, n) * L)
    x_a, x_b = np.minimum(xa, xb), np.maximum(xa, xb)
    x_b = np.where(x_b <= x_a, np.minimum(x_a + 1.0, L), x_b)
    f.update(L=L, dt=dt, T=T, g0=g0, x_a=x_a, x_b=x_b, rho_in=rho_in, rho_out=rho_out, E_in=E_in, E_out=E_out)

    # predicates
    P = rng.integers(1, PMAX + 1, n)
    f["n_pred"] = P.astype(np.int32)
    xl = np.floor(rng.uniform(0, 1, (PMAX, n)) * L)
    xh = np.floor(rng.uniform(0, 1, (PMAX, n)) * L)
    lo, hi = np.minimum(xl, xh), np.maximum(xl, xh)
    hi = np.where(hi <= lo, np.minimum(lo + 1.0, L), hi)
    f["pred_xlo"], f["pred_xhi"] = lo, hi
    if heat:
        m = rng.uniform(-0.5, 0.5, (PMAX, n))
        mid = g0 + rng.uniform(-30.0, 40.0, (PMAX, n))
        b = mid - m * 0.5 * (lo + hi)
    else:
        m = rng.uniform(-5e-5, 5e-5, (PMAX, n))
        b = rng.uniform(-3.0, 3.0, (PMAX, n))
    f["pred_m"], f["pred_b"] = m, b
    r = rng.uniform(0, 1, (PMAX, n))
    f["pred_rel"] = np.where(r < 0.35, abi.REL_GT, np.where(r < 0.70, abi.REL_LT, abi.REL_EQ)).astype(np.uint8)
    f["pred_top"] = rng.integers(0, 2, (PMAX, n)).astype(np.uint8)
    # windows: clause j lives in the j-th of P equal segments of [0, T]
    u0 = rng.uniform(0, 1, (PMAX, n))
    u1 = rng.uniform(0, 1, (PMAX, n))
    seg = T / P
    j = np.arange(PMAX)[:, None]
    a = (j + np.minimum(u0, u1)) * seg
    c = (j + np.maximum(u0, u1)) * seg
    f["pred_tlo"], f["pred_thi"] = np.round(a, 4), np.minimum(np.round(c, 4), T)
    for name in ("pred_xlo", "pred_xhi", "pred_m", "pred_b", "pred_tlo", "pred_thi"):
        f[name] = np.where(j < P, f[name], 0.0)

    # formula tree, children before parents; left-associated; at most one "|" (20 % of the samples)
    tmax = 2 * PMAX - 1
    op = np.zeros((tmax, n), np.uint8)
    lhs = np.zeros((tmax, n), np.int8)
    rhs = np.zeros((tmax, n), np.int8)
    use_or = rng.uniform(0, 1, n) < 0.2
    or_at = rng.integers(1, PMAX, n)          # which join (1-based) becomes the "|"
    # node order: leaf0, leaf1, join1, leaf2, join2, ...
    op[0], lhs[0] = abi.TREE_LEAF, 0
    for q in range(1, PMAX):
        leaf, join = 2 * q - 1, 2 * q
        op[leaf], lhs[leaf] = abi.TREE_LEAF, q
        is_or = use_or & ((or_at == q) | ((P == 2) & (q == 1)))
        op[join] = np.where(is_or, abi.TREE_OR, abi.TREE_AND)
        lhs[join], rhs[join] = (0 if q == 1 else 2 * q - 2), leaf
    f["n_tree"] = (2 * P - 1).astype(np.int32)
    f["tree_op"], f["tree_lhs"], f["tree_rhs"] = op, lhs, rhs

    # control: C11 uniform knots over [0, T]
    kmax = KMAX[workload]
    if heat:
        K = (np.floor(T + 0.5).astype(np.int64) + 1) if workload == "heat" else np.full(n, kmax, np.int64)
        y = rng.uniform(0.0, 1e6, (kmax, n))
    else:
        K = np.full(n, 5, np.int64)
        y = rng.uniform(-5e3, 5e3, (kmax, n))
    jj = np.arange(kmax)[:, None]
    kt = jj * (T / (K - 1))
    kt = np.where(jj == K - 1, T, kt)
    live = jj < K
    f["n_knots"] = K.astype(np.int32)
    f["knot_t"] = np.where(live, kt, 0.0)
    f["knot_y"] = np.where(live, y, 0.0)
    return f


def _copy_into(batch: ProblemBatch, dst_lo: int, f: dict, src_lo: int, cnt: int):
    for name, a in f.items():
        if name == "knot_t" and batch.uniform_knots:
            continue
        dst = getattr(batch, name)
        if a.ndim == 1:
            dst[dst_lo:dst_lo + cnt] = a[src_lo:src_lo + cnt]
        else:
            dst[: a.shape[0], dst_lo:dst_lo + cnt] = a[:, src_lo:src_lo + cnt]


def make_batch(n: int, workload: str = "heat", seed: int = 0, start: int = 0, total: int | None = None,
               buffer=None, uniform_knots: bool = False, kind_slice: tuple | None = None) -> ProblemBatch:
    """Samples [start, start+n) of `workload`.

    mixed: global indices below total//2 are heat, the rest wave (sorted by kind so that warps
    are homogeneous); `total` defaults to start+n.
    kind_slice = (rank, world), mixed only: the n-sample shard of rank `rank` when every rank takes its 1/world slice of
    EACH kind of the global set of world * n samples (heat samples [rank n/2, (rank+1) n/2) and the same wave samples):
    contiguous index ranges of a kind-sorted set would give the first ranks only heat and the last ones only waves,
    and waves cost three times as much (balance mixed batches by cost, not by count).
    """
    if workload not in WORKLOADS:
        raise ValueError(f"unknown workload {workload!r}")
    # the synthetic controls sit on linspace(0, T, K) (C11), so the compact form without knot times describes the same set
    batch = ProblemBatch(n, KMAX[workload], PMAX, buffer=buffer, uniform_knots=uniform_knots)
    if workload == "mixed" and kind_slice is not None:
        rank, world = kind_slice
        if not (0 <= rank < world):
            raise ValueError("kind_slice rank out of range")
        nh = n // 2
        pieces = [("heat", rank * nh, (rank + 1) * nh), ("wave", rank * (n - nh), (rank + 1) * (n - nh))]
    elif workload == "mixed":
        total = start + n if total is None else total
        half = total // 2
        pieces = []
        if start < half:
            pieces.append(("heat", start, min(start + n, half)))
        if start + n > half:
            pieces.append(("wave", max(start, half) - half, start + n - half))
    else:
        pieces = [(workload, start, start + n)]
    dst = 0
    for wl, lo, hi in pieces:
        i = lo
        while i < hi:
            blk, off = divmod(i, BLOCK)
            cnt = min(BLOCK - off, hi - i)
            _copy_into(batch, dst, _gen_block(wl, seed, blk), off, cnt)
            dst += cnt
            i += cnt
    return batch


def replicate_spec(spec: ProblemBatch, s: int, knot_t, knot_y) -> ProblemBatch:
    """The m-sample batch a controls-family call stands for: sample `s` of `spec` once per column of knot_y [nk, m]
    (used by the parity tests and by the CPU arm of the search-mode benchmark)."""
    knot_y = np.asarray(knot_y, dtype=np.float64)
    nk, m = knot_y.shape
    b = ProblemBatch(m, max(nk, 1), spec.pmax, spec.tmax)
    for name in b._names:
        if name in ("knot_t", "knot_y", "n_knots"):
            continue
        src, dst = getattr(spec, name), getattr(b, name)
        if src.ndim == 1:
            dst[:] = src[s]
        else:
            dst[:, :] = src[:, s:s + 1]
    b.n_knots[:] = nk
    b.knot_t[:nk, :] = np.asarray(knot_t, dtype=np.float64)[:, None]
    b.knot_y[:nk, :] = knot_y
    return b

