"""CPU restatement of the reference's IoU metric between two sSTL formulas.  TEST INFRASTRUCTURE ONLY.

Follows test/PDEcontrol/evaluation/eval/eval_script.py of the reference: eval_IOU :233-281 (domains, grids, profiles),
check_conditions :437-487 (cell rule, fold over positions), G_array / F_array :159-219 (bit-array state machines, written
out below as loops over the times of one profile), compute_clause_junction :519-571 and compute_iou :574-582.
PINNED: tests/golden/iou_v0.json holds outputs of the reference's OWN eval_IOU on 16 formula pairs (generated in the build
container by tests/golden/make_iou_golden.py); tests/test_iou.py checks this restatement against every one of them.
Only the formula PARSING is shared with the product (pde_controller_b200.iou.split_formula / parse_clause: host logic);
the cells, the state machines and the combination are restated here independently, in plain numpy / Python loops.
"""
import operator

import numpy as np

OPS = {"<=": operator.le, "<": operator.lt, ">=": operator.ge, ">": operator.gt, "=": operator.eq, "!=": operator.ne}


def clause_bits(c, t_range, x_range, m_vals, b_vals):
    """Bit array of one clause: for every profile (m outer, b inner) the G / F array segment, concatenated as the
    reference's array ends up (G: profiles in order, appended; F: every profile's segment inserted at the FRONT)."""
    t = np.asarray(t_range, dtype=np.float64)
    x = np.asarray(x_range, dtype=np.float64)
    xs = -x if c["neg"] else x
    expr = (xs / c["coef"] if c["kind"] == "div" else c["coef"] * xs) + c["c"]       # two roundings, as Python's eval
    in_t = (t >= c["t0"]) & (t <= c["t1"])
    in_x = (x >= c["p0"]) & (x <= c["p1"])
    whole = []
    for m in m_vals:
        for b in b_vals:
            u = m * t + b
            cond = OPS[c["op"]](u[:, None] - expr[None, :], c["thr"])
            inside = in_t[:, None] & in_x[None, :]
            vac = ~inside.any(axis=1)
            holds = np.where(inside, cond, True).all(axis=1)
            seg = []
            if c["temporal_op"] == "F":
                flag = False
                for k in range(len(t)):                       # the reference walks these times (descending) and inserts at 0
                    if vac[k]:
                        seg.insert(0, True)
                    else:
                        if holds[k]:
                            flag = True
                        seg.insert(0, True if flag else False)
                whole = seg + whole                           # the next profile's segment lands in FRONT of this one
            else:
                flag = False
                for k in range(len(t)):
                    if vac[k]:
                        seg.append(True)
                    elif holds[k]:
                        seg.append(not flag)
                    else:
                        seg = [False] * len(seg)
                        seg.append(False)
                        flag = True
                whole = whole + seg
    return np.array(whole, dtype=bool)


def _leaves(node):
    return [node] if isinstance(node, int) else [k for ch in node for k in _leaves(ch)]


def _combine(node, tree, junctions, arrays):
    """compute_clause_junction: left fold of a bracket level; the junction used between element i and i+1 is the entry of
    the junction list at the flat position of element i's last clause."""
    if isinstance(node, int):
        return arrays[node]
    flat = _leaves(tree)
    acc = _combine(node[0], tree, junctions, arrays)
    for i in range(1, len(node)):
        nxt = _combine(node[i], tree, junctions, arrays)
        if len(nxt) != len(acc):
            raise ValueError("bit arrays of equal length expected")
        which = junctions[flat.index(_leaves(node[i - 1])[-1])]
        acc = (nxt & acc) if which == "and" else (nxt | acc)
    return acc


def eval_iou(gt: str, pred: str):
    """IoU of `pred` against `gt`, or "failed" when a formula cannot be parsed (eval_script.py:248-249)."""
    from pde_controller_b200 import iou as host
    from pde_controller_b200.sstl import SstlError
    try:
        sides = []
        for text in (pred, gt):
            clauses, junctions, tree = host.split_formula(text)
            sides.append(([host.parse_clause(c) for c in clauses], junctions, tree))
    except (SstlError, IndexError, TypeError, ValueError):
        return "failed"
    allc = sides[1][0] + sides[0][0]
    t_min, t_max = min(c["t0"] for c in allc), max(c["t1"] for c in allc)
    x_min, x_max = min(c["p0"] for c in allc), max(c["p1"] for c in allc)
    dt, dx = (t_max - t_min) / host.DT_STEPS, (x_max - x_min) / host.DX_STEPS
    m_vals, b_vals = host._profile_grid()
    out = []
    for clauses, junctions, tree in sides:
        arrays = []
        for c in clauses:
            if c["temporal_op"] == "F":
                tr, xr = np.arange(t_max, t_min - dt - 1e-7, -dt), np.arange(x_max, x_min - dx - 1e-7, -dx)
            else:
                tr, xr = np.arange(t_min, t_max + dt + 1e-7, dt), np.arange(x_min, x_max + dx + 1e-7, dx)
            arrays.append(clause_bits(c, tr, xr, m_vals, b_vals))
        out.append(_combine(tree, tree, junctions, arrays))
    p, g = out
    union = int((p | g).sum())
    return 1 if union == 0 else int((p & g).sum()) / union
