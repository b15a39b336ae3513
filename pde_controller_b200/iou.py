"""IoU between two sSTL formulas on the GPU: drop-in for the reference's `eval_IOU`.

The reference (test/PDEcontrol/evaluation/eval/eval_script.py:233-281) compares a predicted formula with the ground truth
by sampling both: every clause is evaluated for 50 x 50 linear test profiles u(t) = m t + b on a 51 x 51 time x position
grid over the common domain (`check_conditions`, :437-487: one Python `eval` per cell), folded over positions and run
through the G / F bit-array state machines (:159-219); the clauses' arrays are combined along the formula's bracket
structure (`parse_sstl` :283-328, `compute_clause_junction` :519-571) and the metric is |pred AND gt| / |pred OR gt|.
It is the second most expensive CPU metric of the harness (~100 s per clause).

Here the cells are evaluated by `pdeb_iou_grids` (csrc/pdeb_iou.cu), one thread per (clause, profile), with Python's
IEEE operations in Python's order; grids are formed by the same numpy calls; the formula structure and the G / F array
orders (F arrays hold their profiles in REVERSE order: they are built by inserting at the front) are restated here.
Results are bit-identical to the reference's (tests/golden/iou_v0.json: outputs of its own eval_IOU).

    eval_IOU(item, sstl_key="sstl", pred_key="predicted_sstl") -> float | "failed"
    eval_IOU_DPO(item)                                           (the DPO key names of the reference)
    eval_IOU_batch(items, ...)                                   -> list, one GPU call for all pairs
"""
from __future__ import annotations

import ctypes as C
import re
from typing import List, Sequence

import numpy as np

from . import abi, native
from .sstl import SstlError, _match_clause

# the reference's hyper-parameters (eval_script.py:235-239).  Its two assignments `m_min, m_max, d_steps = -7, 7, 10` and
# `b_min, b_max, d_steps = -500, 500, 50` share the name d_steps, so BOTH step sizes are formed with 50: the profile grid is
# 50 x 50 (m in steps of 0.28), not 10 x 50 -- restated as the code behaves, not as it reads
M_RANGE = (-7, 7, 50)
B_RANGE = (-500, 500, 50)
DT_STEPS = DX_STEPS = 50
_SPLIT = re.compile(r"(\\land|\\lor)")


def _profile_grid():
    m_min, m_max, m_steps = M_RANGE
    b_min, b_max, b_steps = B_RANGE
    dm, db = (m_max - m_min) / m_steps, (b_max - b_min) / b_steps
    return np.arange(m_min, m_max, dm), np.arange(b_min, b_max, db)


def split_formula(sstl: str):
    """Clause texts, junctions and bracket structure of a formula, by the reference's rules (parse_sstl, :283-328):
    split at \\land / \\lor; a part that starts with '(' opens a group, a part that closes more parentheses than it opens
    closes groups; inside a group (and at top level) junctions are applied left to right, without precedence."""
    clauses: List[str] = []
    junctions: List[str] = []
    tree: list = []
    stack = [tree]
    balance = 0
    for part in _SPLIT.split(sstl):
        part = part.strip()
        if part == "\\land":
            junctions.append("and")
            continue
        if part == "\\lor":
            junctions.append("or")
            continue
        clauses.append(part)
        k = len(clauses) - 1
        head = part[0]                      # an empty part is an error, as in the reference
        if head in "GFU":
            stack[-1].append(k)
        elif head == "(":
            group = [k]
            stack[-1].append(group)
            stack.append(group)
        now = balance + part.count("(") - part.count(")")
        if now < balance:                   # (the reference only ever records a DEcrease of the balance)
            for _ in range(balance - now):
                if len(stack) > 1:
                    stack.pop()
            balance = now
    return clauses, junctions, tree


def parse_clause(text: str):
    """One clause as the reference's parse_clause reads it (:331-382): temporal operator, time interval, position interval
    truncated to integers, the expression in the spelling it has, comparison operator, threshold."""
    m = None
    for i in range(len(text)):              # the reference searches; a group-opening part carries leading parentheses
        if text[i] in "GF":
            m = _match_clause(text, i)
            if m:
                break
    if not m:
        raise SstlError(f"cannot read clause {text[:60]!r}")
    _, c = m
    kind, coef, neg, const = c["form"]
    return {"temporal_op": c["top"], "t0": c["t0"], "t1": c["t1"], "p0": int(c["a"]), "p1": int(c["b"]), "kind": kind, "coef": coef,
            "neg": neg, "c": const, "op": c["op_raw"], "thr": c["thr"]}


def _flatten(node):
    if isinstance(node, int):
        return [node]
    out = []
    for ch in node:
        out.extend(_flatten(ch))
    return out


def combine(tree, junctions, arrays):
    """The reference's compute_clause_junction (:519-571): fold the elements of every bracket level from left to right;
    the junction between two neighbouring elements is the one behind the LAST clause of the left element."""
    flat = _flatten(tree)

    def ev(node):
        if isinstance(node, int):
            return arrays[node]
        parts = [ev(ch) for ch in node]
        res = parts[0]
        for i in range(1, len(parts)):
            j = junctions[flat.index(_flatten(node[i - 1])[-1])]
            if res.shape != parts[i].shape:
                raise ValueError("bit arrays of equal length expected")
            res = (parts[i] & res) if j == "and" else (parts[i] | res)
        return res

    return ev(tree)


class _Pair:
    """Host-side state of one (ground truth, prediction) pair."""

    def __init__(self, gt: str, pred: str):
        self.sides = []
        for text in (pred, gt):                          # the reference parses the prediction first
            clauses, junctions, tree = split_formula(text)
            self.sides.append(([parse_clause(c) for c in clauses], junctions, tree))
        allc = self.sides[1][0] + self.sides[0][0]
        self.t_min, self.t_max = min(c["t0"] for c in allc), max(c["t1"] for c in allc)
        self.x_min, self.x_max = min(c["p0"] for c in allc), max(c["p1"] for c in allc)
        dt, dx = (self.t_max - self.t_min) / DT_STEPS, (self.x_max - self.x_min) / DX_STEPS
        if dt == 0 or dx == 0:
            raise ZeroDivisionError("float modulo")       # the reference's check_steps divides by the step
        # the four grids, by the calls of check_conditions (:452-457)
        self.t_f = np.arange(self.t_min, self.t_max + dt + 1e-7, dt)
        self.x_f = np.arange(self.x_min, self.x_max + dx + 1e-7, dx)
        self.t_b = np.arange(self.t_max, self.t_min - dt - 1e-7, -dt)
        self.x_b = np.arange(self.x_max, self.x_min - dx - 1e-7, -dx)


def _iou_of(pair: _Pair, bits_of_clause, n_prof):
    out = []
    k = 0
    for clauses, junctions, tree in pair.sides:
        arrays = []
        for c in clauses:
            nt = len(pair.t_b) if c["temporal_op"] == "F" else len(pair.t_f)
            a = bits_of_clause[k][:, :nt].astype(bool)
            if c["temporal_op"] == "F":
                a = a[::-1]                               # F arrays are built by inserting at the front: profiles reversed
            arrays.append(np.ascontiguousarray(a).ravel())
            k += 1
        out.append(combine(tree, junctions, arrays))
    pred, gt = out
    if pred.shape != gt.shape:
        raise ValueError("bit arrays of equal length expected")
    union = int((pred | gt).sum())
    return 1 if union == 0 else int((pred & gt).sum()) / union


def eval_IOU_batch(items: Sequence[dict], sstl_key: str = "sstl", pred_key: str = "predicted_sstl", device: int = 0) -> list:
    """IoU of every item's predicted formula against its ground truth, all grids in ONE GPU call.  An item whose formulas
    cannot be parsed yields "failed" (the reference's own sentinel, :248-249); errors the reference raises behind its
    parser (a degenerate domain, arrays of unequal length) are raised for a single item and reported as "failed" here."""
    m_vals, b_vals = _profile_grid()
    pairs: List = []
    for it in items:
        try:
            pairs.append(_Pair(it[sstl_key], it[pred_key]))
        except Exception as e:      # the reference's parser stage has a bare except -> "failed" (:248-249); a degenerate
            pairs.append(e)         # common domain (its check_steps divides by the step) is reported the same way here
    live = [p for p in pairs if isinstance(p, _Pair)]
    results: list = ["failed"] * len(items)
    if not live:
        return results
    tmax = max(max(len(p.t_f), len(p.t_b)) for p in live)
    xmax = max(max(len(p.x_f), len(p.x_b)) for p in live)
    if tmax > 64:
        raise ValueError("time grid longer than 64 points")
    tg = np.zeros((len(live), 2, tmax))
    xg = np.zeros((len(live), 2, xmax))
    ng = np.zeros((len(live), 4), dtype=np.int32)
    recs = []
    for i, p in enumerate(live):
        tg[i, 0, :len(p.t_f)], tg[i, 1, :len(p.t_b)] = p.t_f, p.t_b
        xg[i, 0, :len(p.x_f)], xg[i, 1, :len(p.x_b)] = p.x_f, p.x_b
        ng[i] = (len(p.t_f), len(p.t_b), len(p.x_f), len(p.x_b))
        for clauses, _, _ in p.sides:
            for c in clauses:
                recs.append((c["t0"], c["t1"], float(c["p0"]), float(c["p1"]), c["coef"], c["c"], c["thr"], i,
                             abi.IOU_DIV if c["kind"] == "div" else abi.IOU_MUL, 1 if c["neg"] else 0, abi.IOU_OP[c["op"]],
                             1 if c["temporal_op"] == "F" else 0))
    arr = (abi.PdebIouClause * len(recs))(*recs)
    n_prof = len(m_vals) * len(b_vals)
    bits = np.empty((len(recs), n_prof, tmax), dtype=np.uint8)
    m_vals, b_vals = np.ascontiguousarray(m_vals, dtype=np.float64), np.ascontiguousarray(b_vals, dtype=np.float64)
    rc = native.lib().pdeb_iou_grids(arr, len(recs), tg.ctypes.data, xg.ctypes.data, ng.ctypes.data, len(live), tmax, xmax,
                                     m_vals.ctypes.data, len(m_vals), b_vals.ctypes.data, len(b_vals), bits.ctypes.data, device,
                                     None, None)
    if rc != 0:
        raise native.NativeError(f"pdeb_iou_grids failed ({rc}): {native.lib().pdeb_last_error().decode()}")
    k = 0
    li = 0
    for idx, p in enumerate(pairs):
        if not isinstance(p, _Pair):
            continue
        ncl = sum(len(s[0]) for s in p.sides)
        try:
            results[idx] = _iou_of(p, bits[k:k + ncl], n_prof)
        except ValueError:
            results[idx] = "failed"
        k += ncl
        li += 1
    return results


def eval_IOU(item, sstl_key="sstl", pred_key="predicted_sstl"):
    """Same call as the reference's eval_IOU (:233)."""
    return eval_IOU_batch([item], sstl_key, pred_key)[0]


def eval_IOU_DPO(item, sstl_key="w_sstl", pred_key="predicted_intermediate_sstl"):
    """eval_script.py:229-230."""
    return eval_IOU(item, sstl_key=sstl_key, pred_key=pred_key)
