"""Problem IR of the utility-evaluation path and its packed batch form.

A `Problem` is what one femformal-dialect program describes (call
sites utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3 and
one_d_wave/examples.jsonl:1-3 of the reference): PDE kind, uniform mesh, two-material rod,
Dirichlet value, step, horizon, APCont predicates and the cspec formula tree.  The control
input (piecewise linear in time, applied at x = L) is a separate argument because the
reference synthesises it inside femformal; here it is given.

`ProblemBatch` packs many problems into the structure-of-arrays layout of
include/pdeb200.h inside ONE contiguous host block, so the library can move it with a
single copy.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi


@dataclass
class Predicate:
    """One clause: TOP_[t_lo,t_hi] ( forall x in [x_lo,x_hi]: u(x) - (m x + b) REL 0 )."""
    x_lo: float
    x_hi: float
    rel: str          # ">", "<", "="
    m: float
    b: float
    top: str = "G"    # "G" or "F"
    t_lo: float = 0.0
    t_hi: float = 0.0
    label: str = ""


@dataclass
class Problem:
    kind: str                      # "heat" | "wave"
    n_elem: int
    L: float
    dt: float
    T: float
    g0: float
    x_a: float
    x_b: float
    rho_in: float
    rho_out: float
    E_in: float
    E_out: float
    predicates: List[Predicate] = field(default_factory=list)
    # children-before-parent list of (op, lhs, rhs); op in {"leaf","and","or"}; root is last
    tree: List[Tuple[str, int, int]] = field(default_factory=list)

    def default_tree(self):
        """Left-associated conjunction of all clauses (the only shape in the 14 in-tree records)."""
        out = [("leaf", 0, 0)]
        acc = 0
        for j in range(1, len(self.predicates)):
            out.append(("leaf", j, 0))
            leaf = len(out) - 1
            out.append(("and", acc, leaf))
            acc = len(out) - 1
        return out


def default_knot_times(p: Problem) -> np.ndarray:
    """C11: uniform knots linspace(0, T, K); K = round(T)+1 for heat, 5 for wave."""
    k = int(round(p.T)) + 1 if p.kind == "heat" else 5
    k = max(k, 2)
    return np.linspace(0.0, p.T, k)


def default_knot_counts(kind: np.ndarray, T: np.ndarray) -> np.ndarray:
    """default_knot_times' K for arrays of samples (kind codes of abi.py): round(T)+1 for heat (round half to even, like
    Python's round), 5 for waves, at least 2."""
    with np.errstate(invalid="ignore"):
        k_heat = np.where(np.isfinite(T), np.rint(np.clip(np.nan_to_num(T), -1e9, 1e9)), 0.0).astype(np.int64) + 1
    return np.maximum(np.where(kind == abi.KIND_HEAT, k_heat, 5), 2)


def fill_default_knot_times(batch: "ProblemBatch", samples, counts: Optional[np.ndarray] = None) -> np.ndarray:
    """n_knots / knot_t of `samples` <- the C11 default knots linspace(0, T, K) (vectorised default_knot_times); the knot
    values of those samples are zeroed.  -> K per sample of `samples`.  K is clipped to the batch's kmax."""
    idx = np.asarray(samples, dtype=np.int64)
    if idx.size == 0:
        return np.zeros(0, dtype=np.int64)
    T = batch.T[idx]
    K = default_knot_counts(batch.kind[idx], T) if counts is None else np.asarray(counts, dtype=np.int64)
    K = np.minimum(K, batch.kmax)
    batch.n_knots[idx] = K
    j = np.arange(batch.kmax, dtype=np.float64)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        kt = j * (T / np.maximum(K - 1, 1))[None, :]
    kt = np.where(j == (K - 1)[None, :], T[None, :], kt)
    kt = np.where(j < K[None, :], kt, 0.0)
    if not batch.uniform_knots:
        batch.knot_t[:, idx] = kt
    batch.knot_y[:, idx] = 0.0
    return K


_F8 = np.dtype(np.float64)      # native float64 arrays share this dtype object: an identity test, not a comparison
_F64 = ["L", "dt", "T", "g0", "x_a", "x_b", "rho_in", "rho_out", "E_in", "E_out"]
_PRED_F64 = ["pred_xlo", "pred_xhi", "pred_m", "pred_b", "pred_tlo", "pred_thi"]
_PLANAR = set(["knot_t", "knot_y", "pred_rel", "pred_top", "tree_op", "tree_lhs", "tree_rhs"] + _PRED_F64)
_TREE_OPS = {"leaf": abi.TREE_LEAF, "and": abi.TREE_AND, "or": abi.TREE_OR}


def _align(x, a=64):
    return (x + a - 1) // a * a


class ProblemBatch:
    """SoA batch in one host slab.  Arrays are numpy views; 2-D arrays are [plane, sample]."""

    def __init__(self, n: int, kmax: int, pmax: int, tmax: Optional[int] = None, buffer=None, uniform_knots: bool = False):
        """uniform_knots: the batch carries no knot times; every control is over linspace(0, T, n_knots) (convention
        C11) and the times are formed on the device (`knot_t == NULL` in the C ABI): ~120 fewer bytes per sample."""
        if tmax is None:
            tmax = 2 * pmax - 1
        self.uniform_knots = bool(uniform_knots)
        if not (1 <= pmax <= abi.MAX_PRED) or not (1 <= tmax <= abi.MAX_TREE) or kmax < 1:
            raise ValueError("kmax/pmax/tmax out of range")
        self.n, self.kmax, self.pmax, self.tmax = int(n), int(kmax), int(pmax), int(tmax)
        self.stride = self.n
        layout = []
        off = 0

        def add(name, dtype, planes):
            nonlocal off
            nbytes = np.dtype(dtype).itemsize * max(self.n, 1) * planes
            layout.append((name, dtype, planes, off))
            off = _align(off + nbytes)

        for f in _F64:
            add(f, np.float64, 1)
        if not self.uniform_knots:
            add("knot_t", np.float64, kmax)
        add("knot_y", np.float64, kmax)
        for f in _PRED_F64:
            add(f, np.float64, pmax)
        for f in ("n_elem", "n_knots", "n_pred", "n_tree"):
            add(f, np.int32, 1)
        add("kind", np.uint8, 1)
        add("pred_rel", np.uint8, pmax)
        add("pred_top", np.uint8, pmax)
        add("tree_op", np.uint8, tmax)
        add("tree_lhs", np.int8, tmax)
        add("tree_rhs", np.int8, tmax)
        self.slab_bytes = off
        if buffer is None:
            buffer = np.zeros(off, dtype=np.uint8)
        else:
            buffer = np.frombuffer(buffer, dtype=np.uint8, count=off)
            buffer[:] = 0
        self.slab = buffer
        self.knot_t = None
        self._names = []
        for name, dtype, planes, o in layout:
            cnt = self.n * planes
            arr = np.frombuffer(self.slab, dtype=dtype, count=cnt, offset=o) if cnt else np.zeros(0, dtype)
            if name in _PLANAR:
                arr = arr.reshape(planes, self.n)
            setattr(self, name, arr)
            self._names.append(name)
        self._parent = None  # set on slices to keep the slab alive

    # ---- construction -----------------------------------------------------
    @staticmethod
    def nbytes_for(n, kmax, pmax, tmax=None, uniform_knots=False):
        """Slab size of a batch, e.g. to allocate it in pinned memory first."""
        if tmax is None:
            tmax = 2 * pmax - 1
        off = 0
        n1 = max(n, 1)
        for planes, isz in ([(1, 8)] * len(_F64) + [(kmax, 8)] * (1 if uniform_knots else 2) + [(pmax, 8)] * len(_PRED_F64)
                            + [(1, 4)] * 4 + [(1, 1), (pmax, 1), (pmax, 1), (tmax, 1), (tmax, 1), (tmax, 1)]):
            off = _align(off + isz * n1 * planes)
        return off

    @classmethod
    def from_problems(cls, problems: Sequence[Problem], controls: Optional[Sequence] = None,
                      kmax: Optional[int] = None, pmax: Optional[int] = None) -> "ProblemBatch":
        n = len(problems)
        if controls is None:
            controls = [None] * n
        if len(controls) != n:
            raise ValueError("controls must have one entry per problem")
        ctrl = []
        for c in controls:
            if c is None:
                ctrl.append((np.zeros(0), np.zeros(0)))
            else:
                t, y = c
                if not (type(t) is np.ndarray and type(y) is np.ndarray and t.dtype is _F8 and y.dtype is _F8
                        and t.ndim == 1 and y.ndim == 1):
                    t = np.asarray(t, dtype=np.float64).ravel()
                    y = np.asarray(y, dtype=np.float64).ravel()
                if t.shape != y.shape:
                    raise ValueError("control knot times and values differ in length")
                ctrl.append((t, y))
        if kmax is None:
            kmax = max([len(t) for t, _ in ctrl] + [1])
        if pmax is None:
            pmax = max([len(p.predicates) for p in problems] + [1])
        tmax = max([len(p.tree) if p.tree else 2 * len(p.predicates) - 1 for p in problems] + [1])
        tmax = max(tmax, 2 * pmax - 1)
        b = cls(n, kmax, pmax, tmax)
        if n < 64:
            for s, p in enumerate(problems):
                b.set_problem(s, p, ctrl[s])
            return b
        # long lists: whole columns at a time (the per-sample route costs ~17 us a sample in the interpreter, which
        # would put the batch assembly, not the parser workers, on the critical path of eval_robustness_batch)
        b.kind[:] = [abi.KIND_HEAT if p.kind == "heat" else abi.KIND_WAVE for p in problems]
        b.n_elem[:] = [p.n_elem for p in problems]
        import operator
        for f in _F64:
            getattr(b, f)[:] = np.fromiter(map(operator.attrgetter(f), problems), dtype=np.float64, count=n)
        preds = [p.predicates for p in problems]
        cnt = np.fromiter(map(len, preds), dtype=np.int64, count=n)
        if cnt.max(initial=0) > pmax:
            raise ValueError("too many predicates for this batch")
        b.n_pred[:] = cnt
        flat = [q for ps in preds for q in ps]                       # sample-major, the order a [n, pmax] mask walks
        m = np.arange(pmax)[None, :] < cnt[:, None]
        tot = len(flat)
        for name, attr in (("pred_xlo", "x_lo"), ("pred_xhi", "x_hi"), ("pred_m", "m"), ("pred_b", "b"),
                           ("pred_tlo", "t_lo"), ("pred_thi", "t_hi")):
            getattr(b, name).T[m] = np.fromiter(map(operator.attrgetter(attr), flat), dtype=np.float64, count=tot)
        b.pred_rel.T[m] = np.fromiter((abi.REL_CODE[q.rel] for q in flat), dtype=np.uint8, count=tot)
        b.pred_top.T[m] = np.fromiter((abi.TOP_CODE[q.top] for q in flat), dtype=np.uint8, count=tot)
        trees = [p.tree if p.tree else p.default_tree() for p in problems]
        tcnt = np.fromiter(map(len, trees), dtype=np.int64, count=n)
        if tcnt.max(initial=0) > tmax:
            raise ValueError("formula tree too large for this batch")
        b.n_tree[:] = tcnt
        tm = np.arange(tmax)[None, :] < tcnt[:, None]
        tflat = [t for tr in trees for t in tr]
        ttot = len(tflat)
        b.tree_op.T[tm] = np.fromiter((_TREE_OPS[t[0]] for t in tflat), dtype=np.uint8, count=ttot)
        b.tree_lhs.T[tm] = np.fromiter((t[1] for t in tflat), dtype=np.int8, count=ttot)
        b.tree_rhs.T[tm] = np.fromiter((t[2] for t in tflat), dtype=np.int8, count=ttot)
        kcnt = np.fromiter((len(t) for t, _ in ctrl), dtype=np.int64, count=n)
        b.n_knots[:] = kcnt
        if kcnt.sum() > 0:
            km = np.arange(kmax)[None, :] < kcnt[:, None]
            b.knot_t.T[km] = np.concatenate([t for t, _ in ctrl])
            b.knot_y.T[km] = np.concatenate([y for _, y in ctrl])
        return b

    def set_problem(self, s: int, p: Problem, control=None):
        self.kind[s] = abi.KIND_HEAT if p.kind == "heat" else abi.KIND_WAVE
        self.n_elem[s] = p.n_elem
        for f in _F64:
            getattr(self, f)[s] = getattr(p, f)
        np_ = len(p.predicates)
        if np_ > self.pmax:
            raise ValueError("too many predicates for this batch")
        self.n_pred[s] = np_
        for j, q in enumerate(p.predicates):
            self.pred_xlo[j, s] = q.x_lo
            self.pred_xhi[j, s] = q.x_hi
            self.pred_m[j, s] = q.m
            self.pred_b[j, s] = q.b
            self.pred_tlo[j, s] = q.t_lo
            self.pred_thi[j, s] = q.t_hi
            self.pred_rel[j, s] = abi.REL_CODE[q.rel]
            self.pred_top[j, s] = abi.TOP_CODE[q.top]
        tree = p.tree if p.tree else p.default_tree()
        if len(tree) > self.tmax:
            raise ValueError("formula tree too large for this batch")
        self.n_tree[s] = len(tree)
        for j, (op, l, r) in enumerate(tree):
            self.tree_op[j, s] = _TREE_OPS[op]
            self.tree_lhs[j, s] = l
            self.tree_rhs[j, s] = r
        if control is not None:
            t, y = control
            if len(t) > self.kmax:
                raise ValueError("too many control knots for this batch")
            self.n_knots[s] = len(t)
            if self.uniform_knots:
                if len(t) and not np.array_equal(np.asarray(t, dtype=np.float64), np.linspace(0.0, float(p.T), len(t))):
                    raise ValueError("this batch carries uniform knots only: control times must be linspace(0, T, K)")
            else:
                self.knot_t[: len(t), s] = t
            self.knot_y[: len(t), s] = y
        else:
            self.n_knots[s] = 0

    def get_problem(self, s: int):
        """Inverse of set_problem: -> (Problem, control or None)."""
        preds = []
        for j in range(int(self.n_pred[s])):
            preds.append(Predicate(float(self.pred_xlo[j, s]), float(self.pred_xhi[j, s]),
                                   abi.REL_TEXT[int(self.pred_rel[j, s])], float(self.pred_m[j, s]),
                                   float(self.pred_b[j, s]), abi.TOP_TEXT[int(self.pred_top[j, s])],
                                   float(self.pred_tlo[j, s]), float(self.pred_thi[j, s]), chr(ord("A") + j)))
        names = {v: k for k, v in _TREE_OPS.items()}
        tree = [(names[int(self.tree_op[j, s])], int(self.tree_lhs[j, s]), int(self.tree_rhs[j, s]))
                for j in range(int(self.n_tree[s]))]
        p = Problem("heat" if self.kind[s] == abi.KIND_HEAT else "wave", int(self.n_elem[s]),
                    *[float(getattr(self, f)[s]) for f in _F64], predicates=preds, tree=tree)
        k = int(self.n_knots[s])
        kt = np.linspace(0.0, float(self.T[s]), k) if self.uniform_knots else self.knot_t[:k, s].copy()
        control = (kt, self.knot_y[:k, s].copy()) if k else None
        return p, control

    # ---- views ------------------------------------------------------------
    def shard(self, lo: int, hi: int) -> "ProblemBatch":
        """Zero-copy view of samples [lo, hi) (same plane stride; no slab fast path)."""
        v = object.__new__(ProblemBatch)
        v.n, v.kmax, v.pmax, v.tmax = hi - lo, self.kmax, self.pmax, self.tmax
        v.stride = self.stride
        v.slab, v.slab_bytes = None, 0
        v.uniform_knots, v.knot_t = self.uniform_knots, None
        v._names = list(self._names)
        for name in self._names:
            a = getattr(self, name)
            setattr(v, name, a[..., lo:hi])
        v._parent = self
        return v

    def as_c(self) -> abi.PdebBatch:
        c = abi.PdebBatch()
        c.n, c.stride = self.n, self.stride
        c.kmax, c.pmax, c.tmax = self.kmax, self.pmax, self.tmax
        if self.slab is not None and self.n > 0:
            c.slab_base = self.slab.ctypes.data
            c.slab_bytes = self.slab_bytes
        else:
            c.slab_base, c.slab_bytes = None, 0
        c.knot_t = None                                   # uniform knots unless the batch carries times
        for name in self._names:
            a = getattr(self, name)
            setattr(c, name, a.ctypes.data if a.size else None)
        c._keepalive = self
        return c

    def steps(self, steps_rule: int = 0) -> np.ndarray:
        """C6 step count per sample (host-side helper for sharding by cost and for sorting)."""
        q = self.T / self.dt
        if steps_rule == 1:
            s = np.floor(q + 0.5)
        elif steps_rule == 2:
            s = np.floor(q + 1e-9)
        else:
            s = np.ceil(q - 1e-9)
        return np.maximum(s, 1).astype(np.int64)

    def algorithmic_bytes(self) -> int:
        """105 + 16 K + 50 P + 3 (2P-1) bytes per evaluation, summed over the batch."""
        k = self.n_knots.astype(np.int64)
        p = self.n_pred.astype(np.int64)
        per_knot = 8 if self.uniform_knots else 16      # knot times are not carried when they are uniform
        return int(np.sum(105 + per_knot * k + 50 * p + 3 * (2 * p - 1)))

    def algorithmic_flops(self, convention: str = "executed", wave_form: str = "a"):
        """FP64 work of the batch, steps up to the last observed one (what the kernels simulate).

        convention "survey":   SURVEY.md §8(d) (FMA = 2): heat S (15 n + 3 sum_p n_p), wave S (21 n + 3 sum_p n_p),
                               + 20 n one-off for assembly and factorisation.
        convention "executed": the operations of the algorithm AS IMPLEMENTED here: per step and row
                               heat add + 3 FMA + mul = 8 flops in 5 instructions, wave 20 flops in 11 (incl. the
                               Newmark predictor / corrector and the low part of the diagonal; wave_form "momentum",
                               the controls family under the default conventions: 15 flops in 9); per node of a clause
                               interval and window step 2 FMA = 4 flops in 2 instructions.
        Returns (flops, fp64_instructions); n = free dofs, n_p = nodes inside predicate p's interval."""
        n = self.n_elem.astype(np.float64)
        dt = self.dt
        S = self.steps().astype(np.float64)
        P = self.n_pred.astype(np.int64)
        h = self.L / n
        node_steps = np.zeros(self.n)
        last = np.zeros(self.n)
        for j in range(self.pmax):
            live = j < P
            ilo = np.clip(np.ceil(self.pred_xlo[j] / h - 1e-9), 0, n)
            ihi = np.clip(np.floor(self.pred_xhi[j] / h + 1e-9), 0, n)
            nodes = np.maximum(ihi - ilo + 1, 2)           # an interval without a node costs its two end points
            a0 = np.clip(np.ceil(self.pred_tlo[j] / dt - 1e-9), 0, S)
            a1 = np.clip(np.floor(self.pred_thi[j] / dt + 1e-9), 0, S)
            steps = np.maximum(a1 - a0 + 1, 1)
            node_steps += np.where(live, nodes * steps, 0.0)
            last = np.maximum(last, np.where(live, a1, 0.0))
        S_run = np.maximum(last, 1.0)
        heat = self.kind == abi.KIND_HEAT
        if convention == "survey":
            flops = np.sum(S_run * np.where(heat, 15.0, 21.0) * n + 3.0 * node_steps + 20.0 * n)
            return int(flops), None
        wf, wi = (15.0, 9.0) if wave_form == "momentum" else (20.0, 11.0)
        flops = np.sum(S_run * np.where(heat, 8.0, wf) * n + 4.0 * node_steps)
        instr = np.sum(S_run * np.where(heat, 5.0, wi) * n + 2.0 * node_steps)
        return int(flops), int(instr)


class PackedBatch:
    """Packed TRANSFER form of a ProblemBatch (PdebPackedBatch of include/pdeb200.h): every per-sample list ragged
    (sample after sample, only the entries in use), the counts narrow, all arrays in one 64-byte-aligned block.

    The end-to-end rate of the path is host->device bandwidth over bytes per sample; this form is what a producer that
    cares about it (the synthetic sweep, a dataset loader) writes.  `capacity_for` gives the worst-case block size of a
    shape, so one page-locked buffer serves every batch of that shape; `pack(batch)` refills it."""

    _SCALARS = [("kind", np.uint8), ("n_elem", np.uint16)] + [(f, np.float64) for f in _F64] + \
               [("n_knots", np.uint16), ("n_pred", np.uint8), ("n_tree", np.uint8)]
    _KNOT = [("knot_t", np.float64), ("knot_y", np.float64)]
    _PRED = [(f, np.float64) for f in _PRED_F64] + [("pred_rel", np.uint8), ("pred_top", np.uint8)]
    _TREE = [("tree_op", np.uint8), ("tree_lhs", np.int8), ("tree_rhs", np.int8)]

    @classmethod
    def capacity_for(cls, n, kmax, pmax, tmax=None, uniform_knots=False):
        tmax = 2 * pmax - 1 if tmax is None else tmax
        total = 0
        for _, dt in cls._SCALARS:
            total = _align(total + np.dtype(dt).itemsize * max(n, 1))
        for name, dt in cls._KNOT:
            if name == "knot_t" and uniform_knots:
                continue
            total = _align(total + np.dtype(dt).itemsize * max(n, 1) * kmax)
        for _, dt in cls._PRED:
            total = _align(total + np.dtype(dt).itemsize * max(n, 1) * pmax)
        for _, dt in cls._TREE:
            total = _align(total + np.dtype(dt).itemsize * max(n, 1) * tmax)
        return total

    def __init__(self, n, kmax, pmax, tmax=None, buffer=None, uniform_knots=False):
        self.n, self.kmax, self.pmax = int(n), int(kmax), int(pmax)
        self.tmax = int(2 * pmax - 1 if tmax is None else tmax)
        self.uniform_knots = bool(uniform_knots)
        cap = self.capacity_for(n, kmax, pmax, self.tmax, uniform_knots)
        if buffer is None:
            buffer = np.zeros(cap + 64, dtype=np.uint8)
        a = buffer.ctypes.data
        self._buffer = buffer
        self._slab = buffer[(-a) % 64:][:cap]
        self.nbytes = 0                      # bytes in use after pack(): what crosses PCIe
        self.totals = (0, 0, 0)
        self._views = {}

    def pack(self, b: "ProblemBatch", use_native: Optional[bool] = None):
        """Refill from a plane-major batch of the same shape (lists come out sample-major).

        use_native: True = the library's host-side packer pdeb_pack_batch (C, 2-3x the numpy route), False = numpy,
        None = native when the library can be loaded.  Both produce the same content (tests compare them)."""
        if (b.n, b.kmax, b.pmax, b.tmax) != (self.n, self.kmax, self.pmax, self.tmax) or b.uniform_knots != self.uniform_knots:
            raise ValueError("batch shape differs from this packed buffer")
        if use_native is None or use_native:
            try:
                from . import native as _native
                L = _native.lib()
            except Exception:
                if use_native:
                    raise
                L = None
            if L is not None:
                return self._pack_native(b, L)
        return self._pack_numpy(b)

    def _pack_native(self, b, L):
        cb = b.as_c()
        c = abi.PdebPackedBatch()
        rc = L.pdeb_pack_batch(C.addressof(cb), self._slab.ctypes.data, self._slab.nbytes, C.addressof(c))
        if rc != 0:
            raise ValueError("pdeb_pack_batch: " + (L.pdeb_last_error() or b"").decode())
        base = self._slab.ctypes.data
        self.nbytes = int(c.slab_bytes)
        self.totals = (int(c.total_knots), int(c.total_pred), int(c.total_tree))
        self.enc = {name: int(c.enc[i]) for name, i in abi.PK_INDEX.items() if not (name == "knot_t" and self.uniform_knots)}
        n, (tk, tp, tt) = self.n, self.totals
        sizes = {}
        for name, dt in self._SCALARS:
            sizes[name] = (np.dtype(dt), n)
        for name, dt in self._KNOT:
            sizes[name] = (np.dtype(dt), tk)
        for name, dt in self._PRED:
            sizes[name] = (np.dtype(dt), tp)
        for name, dt in self._TREE:
            sizes[name] = (np.dtype(dt), tt)
        self._views = {}
        for name, (dt, cnt) in sizes.items():
            ptr = getattr(c, name)
            if not ptr:
                continue
            e = self.enc.get(name, abi.ENC_F64)
            if name in abi.PK_INDEX and e == abi.ENC_F32:
                dt = np.dtype(np.float32)
            elif name in abi.PK_INDEX and e == abi.ENC_ONE:
                cnt = 1
            self._views[name] = (int(ptr) - base, dt, cnt)
        self._c = c
        return self

    def _pack_numpy(self, b: "ProblemBatch"):
        n = self.n
        if b.n_elem.max(initial=0) > 65535 or b.n_knots.max(initial=0) > 65535 or b.n_elem.min(initial=0) < 0 or \
                b.n_knots.min(initial=0) < 0 or b.n_pred.min(initial=0) < 0 or b.n_tree.min(initial=0) < 0 or \
                b.n_pred.max(initial=0) > 255 or b.n_tree.max(initial=0) > 255:
            raise ValueError("a count does not fit the packed form's narrow integers")
        nk = np.minimum(b.n_knots[:n], self.kmax)
        npd = np.minimum(b.n_pred[:n], self.pmax)
        nt = np.minimum(b.n_tree[:n], self.tmax)
        masks = {"k": (np.arange(self.kmax)[None, :] < nk[:, None]), "p": (np.arange(self.pmax)[None, :] < npd[:, None]),
                 "t": (np.arange(self.tmax)[None, :] < nt[:, None])}
        off = 0
        self._views = {}

        self.enc = {}

        def put(name, arr):
            nonlocal off
            if name in abi.PK_INDEX:
                # lossless storage: one value for all / float where every value is exactly a float / as is.
                # Compared as BIT PATTERNS, so -0.0, infinities and NaN payloads survive (or force the f64 form).
                bits = arr.view(np.uint64)
                if arr.size > 1 and np.all(bits == bits[0]):
                    self.enc[name], arr = abi.ENC_ONE, arr[:1].copy()
                elif arr.size > 0 and np.array_equal(arr.astype(np.float32).astype(np.float64).view(np.uint64), bits):
                    self.enc[name], arr = abi.ENC_F32, arr.astype(np.float32)
                else:
                    self.enc[name] = abi.ENC_F64
            nb = arr.nbytes
            self._slab[off:off + nb] = arr.view(np.uint8).reshape(-1)
            self._views[name] = (off, arr.dtype, arr.size)
            off = _align(off + max(nb, 1))

        for name, dt in self._SCALARS:
            put(name, np.ascontiguousarray(getattr(b, name)[:n].astype(dt, copy=False)))
        # counts as the image will hold them: the kernels flag counts beyond the planes, so keep the caller's values
        for name, dt in self._KNOT:
            if name == "knot_t" and self.uniform_knots:
                continue
            put(name, np.ascontiguousarray(getattr(b, name)[:, :n].T[masks["k"]]))
        for name, dt in self._PRED:
            put(name, np.ascontiguousarray(getattr(b, name)[:, :n].T[masks["p"]]))
        for name, dt in self._TREE:
            put(name, np.ascontiguousarray(getattr(b, name)[:, :n].T[masks["t"]]))
        self.nbytes = off
        self.totals = (int(nk.sum()), int(npd.sum()), int(nt.sum()))
        self._c = None
        return self

    def unpack(self) -> "ProblemBatch":
        """The plane-major batch this block stands for (numpy restatement of the device-side unpack; used by the tests).
        Unused list slots are zero."""
        b = ProblemBatch(self.n, self.kmax, self.pmax, self.tmax, uniform_knots=self.uniform_knots)
        base = self._slab

        def get(name):
            off, dt, size = self._views[name]
            a = base[off:off + size * np.dtype(dt).itemsize].view(dt)
            e = self.enc.get(name, abi.ENC_F64)
            if e == abi.ENC_F32:
                return a.astype(np.float64)
            if e == abi.ENC_ONE:
                return a[0]                      # broadcast by the assignment
            return a

        for name, _ in self._SCALARS:
            getattr(b, name)[:] = get(name)
        nk = np.minimum(b.n_knots, self.kmax)
        npd = np.minimum(b.n_pred, self.pmax)
        nt = np.minimum(b.n_tree, self.tmax)
        for group, cnt, planes in ((self._KNOT, nk, self.kmax), (self._PRED, npd, self.pmax), (self._TREE, nt, self.tmax)):
            mask = np.arange(planes)[None, :] < cnt[:, None]
            for name, _ in group:
                if name == "knot_t" and self.uniform_knots:
                    continue
                dst = np.zeros((self.n, planes), dtype=getattr(b, name).dtype)
                dst[mask] = get(name)
                getattr(b, name)[:, : self.n] = dst.T
        return b

    def as_c(self) -> abi.PdebPackedBatch:
        """The C view of the block (cached until the next pack(): it is built once per upload otherwise, on the thread
        that feeds the GPU)."""
        if getattr(self, "_c", None) is not None:
            return self._c
        c = abi.PdebPackedBatch()
        c.n, c.kmax, c.pmax, c.tmax = self.n, self.kmax, self.pmax, self.tmax
        c.total_knots, c.total_pred, c.total_tree = self.totals
        base = self._slab.ctypes.data
        c.slab_base, c.slab_bytes = base, self.nbytes
        for name, (off, _, _) in self._views.items():
            setattr(c, name, base + off)
        if self.uniform_knots:
            c.knot_t = None
        for name, e in self.enc.items():
            c.enc[abi.PK_INDEX[name]] = e
        self._c = c
        return c
