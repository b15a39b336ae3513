"""sSTL LaTeX <-> femformal program cross-check.

Every dataset record of the reference carries the same specification twice: as LaTeX in `sstl`
    G_[[t0, t1]] (\\forall x \\in [a, b]  (u(x) - (m \\cdot x + c) > 0)) \\land F_[[...]] (...) ...
and as the `APCont` / `cspec` lines of `python` (utils/few_shot_prompts/examples/*/examples.jsonl; LaTeX grammar of the
reference: parse_sstl / parse_clause, test/PDEcontrol/evaluation/eval/eval_script.py:283-382).  Before a program is
simulated it is cheap to verify that the two agree - it is the data-quality gate for LLM-written programs, where
`python` and `sstl` come from different generation steps.

    parse_sstl(latex)            -> (predicates, tree)  in the IR of problem.py (same tree form as the cspec parser:
                                    children before parents, `\\land` binds tighter than `\\lor`, parentheses group)
    cross_check(problem, latex)  -> list of human-readable mismatches (empty = the two describe the same clauses)
    to_latex(problem)            -> the LaTeX twin of a Problem (round-trip tests, prompts)

Clause-level failures raise SstlError with the offending text; nothing is evaluated (regular expressions and a
recursive-descent parser only).
"""
from __future__ import annotations

import re
from typing import List, Tuple

from .problem import Predicate, Problem

_NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"
_UNUM = r"(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"          # unsigned
_MUL = r"(?:\\cdot|\\times|\*)"
# header of a clause up to (and including) "u(x)":  TOP_[[t0, t1]] (\forall x \in [a, b] ( u(x)
_HEAD = re.compile(
    r"\s*(?P<top>[GF])_\[\[\s*(?P<t0>" + _NUM + r")\s*,\s*(?P<t1>" + _NUM + r")\s*\]\]\s*"
    r"\(\s*\\forall\s+x\s*\\in\s*\[\s*(?P<a>" + _NUM + r")\s*,\s*(?P<b>" + _NUM + r")\s*\]\s*"
    r"\(\s*u\(x\)\s*")
# the profile inside "- ( ... )", in the five spellings the reference's parse_clause accepts
# (test/PDEcontrol/evaluation/eval/eval_script.py:347-355: pattern1 .. pattern5), each as (slope, intercept) pieces:
#   1  m \cdot x + c            (c may carry its own sign: "+ -1.7281")
#   2  x / (d) + c               slope 1/d; the sign in front of x negates it
#   3  x \cdot m + c
#   4  x \cdot (m) + c
#   5  mx + c                    juxtaposition
_TAIL = r"(?:\s*(?P<sgn>[-+])\s*(?P<c>" + _NUM + r"))?\s*"
_PROFILES = [
    re.compile(r"(?P<m>" + _NUM + r")\s*" + _MUL + r"\s*x" + _TAIL),
    re.compile(r"(?P<xs>[-+]?)\s*x\s*/\s*\(?\s*(?P<d>" + _UNUM + r")\s*\)?" + _TAIL),
    re.compile(r"(?P<xs>[-+]?)\s*x\s*" + _MUL + r"\s*\(?\s*(?P<m>" + _NUM + r")\s*\)?" + _TAIL),
    re.compile(r"(?P<m>" + _NUM + r")\s*x" + _TAIL),
]
_REL = re.compile(r"\s*(?P<rel>\\leq|\\geq|\\neq|<=|>=|!=|<|>|=)\s*(?P<thr>" + _NUM + r")\s*\)\s*\)")
_JUNCTION = re.compile(r"\s*(\\land|\\lor|\\wedge|\\vee|&|\|)")


class SstlError(ValueError):
    pass


def _rel(tok: str) -> str:
    """The relation of the IR: the reference's clause grammar also knows <=, >=, \\leq, \\geq (the same robustness as
    the strict forms) and != / \\neq, which no APCont can state: it is kept as "!=" so that the cross-check reports it."""
    return {"\\leq": "<", "<=": "<", "\\geq": ">", ">=": ">", "\\neq": "!="}.get(tok, tok)


def _match_clause(s: str, i: int):
    """One clause at position i -> (end, dict) or None.  dict: top, t0, t1, a, b (floats), slope, icpt, rel, thr."""
    h = _HEAD.match(s, i)
    if not h:
        return None
    j = h.end()
    slope, icpt = 0.0, 0.0
    form = ("mul", 0.0, False, 0.0)          # (kind, coefficient, x negated, constant): expr(x) as the text spells it
    if j < len(s) and s[j] == "-":
        # u(x) - ( PROFILE )
        k = j + 1
        while k < len(s) and s[k].isspace():
            k += 1
        if k >= len(s) or s[k] != "(":
            return None
        k += 1
        while k < len(s) and s[k].isspace():
            k += 1
        hit = None
        for rx in _PROFILES:
            m = rx.match(s, k)
            if m and m.end() < len(s) and s[m.end()] == ")":
                hit = m
                break
        if hit is None:
            return None
        g = hit.groupdict()
        if g.get("d") is not None:
            d = float(g["d"])
            if d == 0.0:
                raise SstlError("sstl: division by zero in a profile")
            slope = 1.0 / d
        else:
            slope = float(g["m"])
        if g.get("xs") == "-":
            slope = -slope
        if g.get("c") is not None:
            icpt = float(g["c"]) * (-1.0 if g["sgn"] == "-" else 1.0)
        if g.get("d") is not None:
            form = ("div", float(g["d"]), g.get("xs") == "-", icpt)
        else:
            form = ("mul", float(g["m"]), g.get("xs") == "-", icpt)
        j = hit.end() + 1
    r = _REL.match(s, j)
    if not r:
        return None
    out = {"top": h.group("top"), "t0": float(h.group("t0")), "t1": float(h.group("t1")), "a": float(h.group("a")),
           "b": float(h.group("b")), "slope": slope, "icpt": icpt, "rel": _rel(r.group("rel")), "thr": float(r.group("thr")),
           "form": form, "op_raw": {"\\leq": "<=", "\\geq": ">=", "\\neq": "!="}.get(r.group("rel"), r.group("rel")),
           "text": s[i:r.end()].strip()}
    return r.end(), out


class _Parser:
    """or := and (lor and)* ; and := unary (land unary)* ; unary := clause | '(' or ')'"""

    def __init__(self, text: str, truncate_positions: bool = False):
        self.s, self.i = text, 0
        self.truncate_positions = truncate_positions
        self.preds: List[Predicate] = []
        self.tree: List[Tuple[str, int, int]] = []

    def ws(self):
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1

    def junction(self):
        m = _JUNCTION.match(self.s, self.i)
        if not m:
            return None
        return "and" if m.group(1) in ("\\land", "\\wedge", "&") else "or"

    def take_junction(self):
        m = _JUNCTION.match(self.s, self.i)
        self.i = m.end()

    def p_or(self) -> int:
        e = self.p_and()
        while self.junction() == "or":
            self.take_junction()
            r = self.p_and()
            self.tree.append(("or", e, r))
            e = len(self.tree) - 1
        return e

    def p_and(self) -> int:
        e = self.p_unary()
        while self.junction() == "and":
            self.take_junction()
            r = self.p_unary()
            self.tree.append(("and", e, r))
            e = len(self.tree) - 1
        return e

    def p_unary(self) -> int:
        self.ws()
        hit = _match_clause(self.s, self.i)
        if hit:
            self.i, c = hit
            t0, t1, a, b = c["t0"], c["t1"], c["a"], c["b"]
            if self.truncate_positions:
                # the reference reads the position interval as int(float(.)) (eval_script.py:341): truncation towards zero
                a, b = float(int(a)), float(int(b))
            if t1 < t0 or b < a:
                raise SstlError(f"sstl: reversed interval in {c['text']!r}")
            label = chr(ord("A") + len(self.preds)) if len(self.preds) < 26 else f"P{len(self.preds)}"
            # u - (m x + c) REL thr   <=>   u - (m x + c + thr) REL 0
            self.preds.append(Predicate(a, b, c["rel"], c["slope"], c["icpt"] + c["thr"], c["top"], t0, t1, label))
            self.tree.append(("leaf", len(self.preds) - 1, 0))
            return len(self.tree) - 1
        if self.i < len(self.s) and self.s[self.i] == "(":
            self.i += 1
            e = self.p_or()
            self.ws()
            if self.i >= len(self.s) or self.s[self.i] != ")":
                raise SstlError("sstl: missing ')'")
            self.i += 1
            return e
        raise SstlError(f"sstl: cannot read a clause at {self.s[self.i:self.i + 60]!r}")


def parse_sstl(latex: str, truncate_positions: bool = False):
    """-> (predicates, tree); labels A, B, C ... in order of appearance, as the reference's programs name them.

    truncate_positions: read the ends of every position interval as int(float(.)), the rule of the reference's
    parse_clause (eval_script.py:341); off by default because the programs carry the untruncated numbers."""
    if not isinstance(latex, str) or not latex.strip():
        raise SstlError("sstl is empty")
    text = latex.replace("\\left", "").replace("\\right", "").replace("$", "")
    p = _Parser(text, truncate_positions)
    root = p.p_or()
    p.ws()
    if p.i != len(p.s):
        raise SstlError(f"sstl: trailing text {p.s[p.i:p.i + 40]!r}")
    assert root == len(p.tree) - 1
    return p.preds, p.tree


def _canon_tree(tree, n_leaf_map=None):
    """Nested tuple form of a children-before-parent tree, with and/or operands sorted (both are commutative and
    associative, so the comparison is insensitive to the order the two notations list the clauses in)."""
    def build(j):
        op, l, r = tree[j]
        if op == "leaf":
            return ("leaf", l)
        parts = []
        for child in (build(l), build(r)):
            if child[0] == op:
                parts.extend(child[1])              # flatten (a & b) & c
            else:
                parts.append(child)
        return (op, tuple(sorted(parts, key=repr)))
    return build(len(tree) - 1)


def cross_check(problem: Problem, latex: str, rtol: float = 1e-9, atol: float = 1e-12,
                truncate_positions: bool = False) -> List[str]:
    """Mismatches between the program's clauses / formula and the LaTeX specification (empty list = consistent).
    truncate_positions compares the position intervals as the reference's parse_clause reads them: int(float(.)) on
    both sides."""
    try:
        preds, tree = parse_sstl(latex, truncate_positions)
    except SstlError as e:
        return [str(e)]
    out = []
    if len(preds) != len(problem.predicates):
        return [f"clause count: program {len(problem.predicates)}, sstl {len(preds)}"]

    def close(x, y):
        return abs(x - y) <= atol + rtol * max(abs(x), abs(y))

    for j, (p, q) in enumerate(zip(problem.predicates, preds)):
        name = p.label or chr(ord("A") + j)
        if p.top != q.top:
            out.append(f"{name}: temporal operator {p.top} vs {q.top}")
        if p.rel != q.rel:
            out.append(f"{name}: relation {p.rel!r} vs {q.rel!r}")
        pxl, pxh = (float(int(p.x_lo)), float(int(p.x_hi))) if truncate_positions else (p.x_lo, p.x_hi)
        for what, x, y in (("t_lo", p.t_lo, q.t_lo), ("t_hi", p.t_hi, q.t_hi), ("x_lo", pxl, q.x_lo), ("x_hi", pxh, q.x_hi),
                           ("slope", p.m, q.m), ("intercept", p.b, q.b)):
            if not close(float(x), float(y)):
                out.append(f"{name}: {what} {x!r} vs {y!r}")
    ptree = problem.tree if problem.tree else problem.default_tree()
    if _canon_tree(ptree) != _canon_tree(tree):
        out.append("formula structure differs")
    return out


def to_latex(problem: Problem) -> str:
    """LaTeX twin of the problem's specification, in the reference's canonical spelling."""
    def clause(p: Predicate) -> str:
        return (f"{p.top}_[[{p.t_lo!r}, {p.t_hi!r}]] (\\forall x \\in [{p.x_lo!r}, {p.x_hi!r}]  "
                f"(u(x) - ({p.m!r} \\cdot x + {p.b!r}) {p.rel} 0))")

    tree = problem.tree if problem.tree else problem.default_tree()

    def render(j, parent=None):
        op, l, r = tree[j]
        if op == "leaf":
            return clause(problem.predicates[l])
        junction = "\\land" if op == "and" else "\\lor"
        s = f"{render(l, op)} {junction} {render(r, op)}"
        return f"({s})" if parent is not None and parent != op else s

    return render(len(tree) - 1)


def check_records(items, python_key: str = "python", stl_key: str = "sstl"):
    """Data-quality gate over dataset records (the keys are the harness's own: python_key / stl_key,
    train/validate.py:28-64 of the reference): -> [(index, [mismatch, ...])] for the records whose program and LaTeX
    specification disagree or whose program does not parse."""
    from .frontend import ParseError, parse_program
    bad = []
    for i, it in enumerate(items):
        try:
            problem = parse_program(it[python_key])
        except (ParseError, KeyError) as e:
            bad.append((i, [f"program: {e}"]))
            continue
        issues = cross_check(problem, it.get(stl_key, ""))
        if issues:
            bad.append((i, issues))
    return bad

