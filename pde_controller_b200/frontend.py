"""Front-end: femformal-dialect program text -> `Problem`, without exec.

The reference hands the generated program to a Python-2 interpreter that executes it inside
femformal (eval_robustness_wrapper.py:13-16; the dialect is visible in
utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3 and one_d_wave/examples.jsonl:1-3,
SURVEY.md Appendix A).  Programs come from an LLM, so here the text is parsed with `ast` and
interpreted by a whitelist: numeric arithmetic, lists / dicts / strings, lambdas (evaluated
symbolically only at mesh points), and the handful of calls the dialect uses
(np.linspace / np.zeros, heatlinfem.heatlinfem_mix, mechlinfem.mechlinfem, logic.APCont).
Anything else raises ParseError, which the batched entry points turn into status PARSE / -100.
"""
from __future__ import annotations

import ast
import math
import re
import textwrap
import warnings
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import abi
from .problem import Predicate, Problem


class ParseError(ValueError):
    pass


# ---------------------------------------------------------------------------
# restricted expression interpreter
# ---------------------------------------------------------------------------
class _Module:
    def __init__(self, name):
        self.name = name


class _Lambda:
    def __init__(self, node: ast.Lambda, env: dict):
        if node.args.vararg or node.args.kwarg or node.args.kwonlyargs or len(node.args.args) != 1:
            raise ParseError("only one-argument lambdas are part of the dialect")
        self.arg = node.args.args[0].arg
        self.body = node.body
        self.env = env

    def __call__(self, x):
        return _eval(self.body, self.env, {self.arg: x})


class _Affine:
    """c1 * x + c0, propagated through + - * / so that profile coefficients come out exactly as written."""

    def __init__(self, c1, c0):
        self.c1, self.c0 = float(c1), float(c0)

    @staticmethod
    def lift(v):
        if isinstance(v, _Affine):
            return v
        if isinstance(v, bool) or not isinstance(v, _NUM):
            raise ParseError("profile is not arithmetic on numbers")
        return _Affine(0.0, v)

    def __add__(self, o):
        o = _Affine.lift(o)
        return _Affine(self.c1 + o.c1, self.c0 + o.c0)

    __radd__ = __add__

    def __sub__(self, o):
        o = _Affine.lift(o)
        return _Affine(self.c1 - o.c1, self.c0 - o.c0)

    def __rsub__(self, o):
        return _Affine.lift(o) - self

    def __mul__(self, o):
        o = _Affine.lift(o)
        if self.c1 != 0.0 and o.c1 != 0.0:
            raise ParseError("profile is not affine in x")
        if o.c1 == 0.0:
            return _Affine(self.c1 * o.c0, self.c0 * o.c0)
        return _Affine(o.c1 * self.c0, o.c0 * self.c0)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = _Affine.lift(o)
        if o.c1 != 0.0 or o.c0 == 0.0:
            raise ParseError("profile is not affine in x")
        return _Affine(self.c1 / o.c0, self.c0 / o.c0)

    def __neg__(self):
        return _Affine(-self.c1, -self.c0)

    def __pos__(self):
        return self


class _System:
    def __init__(self, kind, args):
        self.kind, self.args = kind, args


class _APCont:
    def __init__(self, A, r, p, dp):
        self.A, self.r, self.p, self.dp = A, r, p, dp


_BINOPS = {
    ast.Add: lambda a, b: a + b, ast.Sub: lambda a, b: a - b, ast.Mult: lambda a, b: a * b,
    ast.Div: lambda a, b: a / b, ast.Pow: lambda a, b: a ** b, ast.FloorDiv: lambda a, b: a // b,
    ast.Mod: lambda a, b: a % b,
}
_CMPOPS = {
    ast.Lt: lambda a, b: a < b, ast.LtE: lambda a, b: a <= b, ast.Gt: lambda a, b: a > b,
    ast.GtE: lambda a, b: a >= b, ast.Eq: lambda a, b: a == b, ast.NotEq: lambda a, b: a != b,
}
_NUM = (int, float, np.integer, np.floating)
_SAFE_FUNCS = {"float": float, "int": int, "abs": abs, "min": min, "max": max, "round": round, "len": len}
_MODULES = {"np": "numpy", "numpy": "numpy", "math": "math", "logic": "logic", "heatlinfem": "heatlinfem",
            "mechlinfem": "mechlinfem"}


def _dotted(node) -> Optional[str]:
    parts = []
    while isinstance(node, ast.Attribute):
        parts.append(node.attr)
        node = node.value
    if isinstance(node, ast.Name):
        parts.append(node.id)
        return ".".join(reversed(parts))
    return None


def _num(v, what="value"):
    if isinstance(v, bool) or not isinstance(v, _NUM):
        raise ParseError(f"{what} is not a number")
    return float(v)


def _call(name: str, args: list, kwargs: dict):
    base, _, fn = name.rpartition(".")
    mod = base.split(".")[-1] if base else ""
    if mod in ("np", "numpy"):
        if fn == "linspace" and 2 <= len(args) <= 3 and not kwargs:
            n = int(args[2]) if len(args) == 3 else 50
            if not (2 <= n <= 1_000_000):
                raise ParseError("linspace size out of range")
            return np.linspace(_num(args[0]), _num(args[1]), n)
        if fn in ("zeros", "ones") and len(args) == 1:
            n = int(args[0])
            if not (1 <= n <= 1_000_000):
                raise ParseError("array size out of range")
            return np.zeros(n) if fn == "zeros" else np.ones(n)
        if fn == "array" and len(args) == 1:
            return np.asarray(args[0], dtype=np.float64)
        if fn in ("sqrt", "exp", "log", "sin", "cos", "abs") and len(args) == 1:
            return float(getattr(np, fn)(_num(args[0])))
    if mod == "math" and fn in ("sqrt", "exp", "log", "sin", "cos", "fabs", "floor", "ceil") and len(args) == 1:
        return float(getattr(math, fn)(_num(args[0])))
    if fn in ("heatlinfem_mix", "heatlinfem") and (mod == "heatlinfem" or not mod):
        return _System("heat", _bind(args, kwargs, ["xpart", "rho", "E", "g", "f_nodal", "dt"]))
    if fn == "mechlinfem" and (mod == "mechlinfem" or not mod):
        return _System("wave", _bind(args, kwargs, ["xpart", "rho", "E", "g", "f_nodal", "dt"]))
    if fn == "APCont":
        a = _bind(args, kwargs, ["A", "r", "p", "dp"], optional=("dp",))
        return _APCont(a["A"], a["r"], a["p"], a.get("dp"))
    if not base and fn in _SAFE_FUNCS:
        return _SAFE_FUNCS[fn](*args)
    raise ParseError(f"call to {name!r} is not part of the dialect")


def _bind(args, kwargs, names, optional=()):
    if len(args) > len(names):
        raise ParseError("too many arguments")
    out = dict(zip(names, args))
    for k, v in kwargs.items():
        if k not in names or k in out:
            raise ParseError(f"unexpected argument {k!r}")
        out[k] = v
    for n in names:
        if n not in out and n not in optional:
            raise ParseError(f"missing argument {n!r}")
    return out


def _eval(node, env: dict, local: Optional[dict] = None):
    if isinstance(node, ast.Constant):
        if isinstance(node.value, (int, float, str, bool)) or node.value is None:
            return node.value
        raise ParseError("unsupported constant")
    if isinstance(node, ast.Name):
        if local and node.id in local:
            return local[node.id]
        if node.id in env:
            return env[node.id]
        if node.id in _MODULES:
            return _Module(node.id)
        if node.id in ("None", "True", "False"):
            return {"None": None, "True": True, "False": False}[node.id]
        raise ParseError(f"name {node.id!r} is not defined")
    if isinstance(node, ast.UnaryOp):
        v = _eval(node.operand, env, local)
        if isinstance(node.op, ast.USub):
            return -v
        if isinstance(node.op, ast.UAdd):
            return +v
        if isinstance(node.op, ast.Not):
            return not v
        raise ParseError("unsupported unary operator")
    if isinstance(node, ast.BinOp):
        op = _BINOPS.get(type(node.op))
        if op is None:
            raise ParseError("unsupported binary operator")
        a, b = _eval(node.left, env, local), _eval(node.right, env, local)
        if isinstance(a, _Affine) or isinstance(b, _Affine):
            if isinstance(node.op, (ast.Add, ast.Sub, ast.Mult, ast.Div)) and not isinstance(b if isinstance(a, _Affine) else a, (np.ndarray, bool, str, list, dict, type(None))):
                try:
                    return op(a, b)
                except TypeError:
                    raise ParseError("profile is not affine in x")
            raise ParseError("profile is not affine in x")
        if not isinstance(a, _NUM + (np.ndarray,)) or not isinstance(b, _NUM + (np.ndarray,)):
            raise ParseError("arithmetic on non-numbers")
        if isinstance(node.op, ast.Pow) and (abs(float(np.max(np.abs(b)))) > 64):
            raise ParseError("exponent too large")
        try:
            return op(a, b)
        except ZeroDivisionError:
            raise ParseError("division by zero")
    if isinstance(node, ast.BoolOp):
        vals = [_eval(v, env, local) for v in node.values]
        return any(vals) if isinstance(node.op, ast.Or) else all(vals)
    if isinstance(node, ast.Compare):
        left = _eval(node.left, env, local)
        for op, comp in zip(node.ops, node.comparators):
            right = _eval(comp, env, local)
            f = _CMPOPS.get(type(op))
            if f is None:
                raise ParseError("unsupported comparison")
            if not f(left, right):
                return False
            left = right
        return True
    if isinstance(node, ast.IfExp):
        return _eval(node.body if _eval(node.test, env, local) else node.orelse, env, local)
    if isinstance(node, (ast.List, ast.Tuple)):
        return [_eval(e, env, local) for e in node.elts]
    if isinstance(node, ast.Dict):
        return {_eval(k, env, local): _eval(v, env, local) for k, v in zip(node.keys, node.values)}
    if isinstance(node, ast.Lambda):
        return _Lambda(node, env)
    if isinstance(node, ast.Subscript):
        v = _eval(node.value, env, local)
        i = _eval(node.slice, env, local)
        try:
            return v[i]
        except Exception:
            raise ParseError("bad subscript")
    if isinstance(node, ast.Attribute):
        name = _dotted(node)
        if name in ("np.pi", "numpy.pi", "math.pi"):
            return math.pi
        if name in ("np.inf", "numpy.inf", "math.inf"):
            return math.inf
        raise ParseError(f"attribute {name!r} is not part of the dialect")
    if isinstance(node, ast.Call):
        name = _dotted(node.func)
        if name is None:
            raise ParseError("unsupported call")
        args = [_eval(a, env, local) for a in node.args]
        kwargs = {k.arg: _eval(k.value, env, local) for k in node.keywords if k.arg}
        return _call(name, args, kwargs)
    raise ParseError(f"unsupported syntax: {type(node).__name__}")


# ---------------------------------------------------------------------------
# cspec: "((F_[t0, t1] (A)) & (G_[t2, t3] (B)))"
# ---------------------------------------------------------------------------
_TOK = re.compile(r"\s*(?:(?P<top>[GF])_\[\s*(?P<a>[-+0-9.eE]+)\s*,\s*(?P<b>[-+0-9.eE]+)\s*\]|(?P<par>[()])|(?P<op>&&?|\|\|?)"
                  r"|(?P<id>[A-Za-z_][A-Za-z_0-9]*))")


def _tokenize(spec: str):
    pos, out = 0, []
    spec = spec.strip()
    while pos < len(spec):
        m = _TOK.match(spec, pos)
        if not m or m.end() == pos:
            raise ParseError(f"cspec: cannot read {spec[pos:pos + 12]!r}")
        pos = m.end()
        if m.group("top"):
            try:
                out.append(("T", m.group("top"), float(m.group("a")), float(m.group("b"))))
            except ValueError:
                raise ParseError("cspec: bad window bound")
        elif m.group("par"):
            out.append((m.group("par"),))
        elif m.group("op"):
            out.append(("&",) if m.group("op")[0] == "&" else ("|",))
        else:
            out.append(("L", m.group("id")))
    return out


class _SpecParser:
    """or := and ('|' and)* ; and := unary ('&' unary)* ; unary := T unary | '(' or ')' | LABEL"""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def parse(self):
        e = self.p_or()
        if self.peek() is not None:
            raise ParseError("cspec: trailing tokens")
        return e

    def p_or(self):
        e = self.p_and()
        while self.peek() == ("|",):
            self.i += 1
            e = ("or", e, self.p_and())
        return e

    def p_and(self):
        e = self.p_unary()
        while self.peek() == ("&",):
            self.i += 1
            e = ("and", e, self.p_unary())
        return e

    def p_unary(self):
        tok = self.peek()
        if tok is None:
            raise ParseError("cspec: unexpected end")
        if tok[0] == "T":
            self.i += 1
            return ("T", tok[1], tok[2], tok[3], self.p_unary())
        if tok == ("(",):
            self.i += 1
            e = self.p_or()
            if self.peek() != (")",):
                raise ParseError("cspec: missing ')'")
            self.i += 1
            return e
        if tok[0] == "L":
            self.i += 1
            return ("L", tok[1])
        raise ParseError("cspec: unexpected token")


def _push_temporal(top, a, b, e):
    """Rewrite TOP_[a,b](e) into a boolean tree of single-label clauses where that is exact."""
    if e[0] == "L":
        return ("C", top, a, b, e[1])
    if e[0] == "T":
        raise ParseError("cspec: nested temporal operators are not supported")
    if (e[0] == "and" and top == "G") or (e[0] == "or" and top == "F"):   # min/min and max/max commute
        return (e[0], _push_temporal(top, a, b, e[1]), _push_temporal(top, a, b, e[2]))
    raise ParseError("cspec: G over '|' / F over '&' is not a tree of single-region clauses")


def _normalise(e):
    if e[0] == "L":
        return ("C", "G", 0.0, 0.0, e[1])          # bare region label: the predicate at t = 0
    if e[0] == "T":
        return _push_temporal(e[1], e[2], e[3], e[4])
    return (e[0], _normalise(e[1]), _normalise(e[2]))


def parse_cspec(spec: str) -> Tuple[List[Tuple[str, float, float, str]], List[Tuple[str, int, int]]]:
    """-> (clauses [(top, t_lo, t_hi, label)], postfix tree [(op, lhs, rhs)])."""
    if not isinstance(spec, str) or not spec.strip():
        raise ParseError("cspec is empty")
    e = _normalise(_SpecParser(_tokenize(spec)).parse())
    clauses, tree = [], []

    def emit(n):
        if n[0] == "C":
            if not (n[3] >= n[2]) or n[2] < 0:
                raise ParseError("cspec: bad window")
            clauses.append((n[1], n[2], n[3], n[4]))
            tree.append(("leaf", len(clauses) - 1, 0))
        else:
            l = emit(n[1])
            r = emit(n[2])
            tree.append((n[0], l, r))
        return len(tree) - 1

    emit(e)
    if len(clauses) > abi.MAX_PRED:
        raise ParseError("cspec: too many clauses")
    return clauses, tree


# ---------------------------------------------------------------------------
# program -> Problem
# ---------------------------------------------------------------------------
def _material(rho, E, xpart) -> Tuple[float, float, float, float, float, float]:
    """-> (x_a, x_b, rho_in, rho_out, E_in, E_out) from the two callables (or constants)."""
    L = float(xpart[-1])

    def canon(f):
        # lambda x: OUT if x < a or x > b else IN
        if not isinstance(f, _Lambda) or not isinstance(f.body, ast.IfExp):
            return None
        t = f.body.test
        if not (isinstance(t, ast.BoolOp) and isinstance(t.op, ast.Or) and len(t.values) == 2):
            return None
        c0, c1 = t.values
        ok = all(isinstance(c, ast.Compare) and len(c.ops) == 1 and isinstance(c.left, ast.Name) and c.left.id == f.arg
                 for c in (c0, c1))
        if not ok or not isinstance(c0.ops[0], ast.Lt) or not isinstance(c1.ops[0], ast.Gt):
            return None
        try:
            a = _num(_eval(c0.comparators[0], f.env))
            b = _num(_eval(c1.comparators[0], f.env))
            vo = _num(_eval(f.body.body, f.env, {f.arg: 0.0}))
            vi = _num(_eval(f.body.orelse, f.env, {f.arg: 0.0}))
        except ParseError:
            return None
        return a, b, vi, vo

    cr, cE = canon(rho), canon(E)
    if cr and cE and cr[0] == cE[0] and cr[1] == cE[1]:
        return cr[0], cr[1], cr[2], cr[3], cE[2], cE[3]

    # general callables / constants: sample at the element midpoints and fit "outer, inner, outer"
    def sample(f):
        mids = 0.5 * (xpart[:-1] + xpart[1:])
        if isinstance(f, _Lambda):
            return np.array([_num(f(float(x)), "material value") for x in mids])
        return np.full(len(mids), _num(f, "material value"))

    r, e = sample(rho), sample(E)
    pair = np.stack([r, e], 1)
    change = np.nonzero(np.any(pair[1:] != pair[:-1], axis=1))[0] + 1
    if len(change) == 0:
        return 0.0, L, float(r[0]), float(r[0]), float(e[0]), float(e[0])
    if len(change) == 1:      # two segments: call the right one "inner" reaching to the end of the rod
        k = change[0]
        return float(xpart[k]), L, float(r[-1]), float(r[0]), float(e[-1]), float(e[0])
    if len(change) == 2 and np.all(pair[0] == pair[-1]):
        k0, k1 = change
        return float(xpart[k0]), float(xpart[k1]), float(r[k0]), float(r[0]), float(e[k0]), float(e[0])
    raise ParseError("material is not a two-material rod with one inner segment")


def _affine(p: _Lambda, what: str) -> Tuple[float, float]:
    if not isinstance(p, _Lambda):
        raise ParseError(f"{what}: profile must be a lambda")
    try:
        v = p(_Affine(1.0, 0.0))
    except TypeError:
        raise ParseError(f"{what}: profile is not affine in x")
    v = _Affine.lift(v)
    if not (math.isfinite(v.c1) and math.isfinite(v.c0)):
        raise ParseError(f"{what}: profile is not finite")
    return v.c1, v.c0


# ---------------------------------------------------------------------------
# fast path: the canonical template
# ---------------------------------------------------------------------------
# Every program of the reference's records (and of render_program) is the SAME dozen lines with different numbers
# (SURVEY.md Appendix A).  Interpreting one through `ast` costs ~1 ms; matching the template with one regular expression
# and converting its numbers costs ~40 us.  The match is deliberately strict (anchored, fixed spelling of every line,
# thresholds of rho and E textually identical, labels A, B, C... in order, a conjunction spelt flat or left-nested): anything else is NOT an error here, it simply falls through to the general interpreter below, so the two
# routes cannot disagree on what they both accept (tests compare them on every record, on synthetic sets and under fuzz).
# a number: written so that every string has ONE way to match (digits are never shared between two quantifiers), or a
# near-miss of the template would make the matcher backtrack exponentially; generated text is untrusted input
_RE_NUM = r"[-+]?(?:\d+(?:\.\d*)?|\.\d+)(?:[eE][-+]?\d+)?"
_VAL = rf"({_RE_NUM})(?:\s*\*\s*({_RE_NUM}))?"

def _fast_regex():
    num = _RE_NUM
    val = lambda a, b: rf"(?P<{a}>{num})(?:\*(?P<{b}>{num}))?"
    pat = (
        rf"(?:from femformal\.core\.fem import (?P<imod>heatlinfem|mechlinfem) as (?P=imod)\n+)?"
        rf"N = (?P<N>\d+)\n+"
        rf"L = (?P<L>{num})\n+"
        rf"rho = lambda x: {val('ro1', 'ro2')} if x < (?P<xa>{num}) or x > (?P<xb>{num}) else {val('ri1', 'ri2')}\n+"
        rf"E = lambda x: {val('eo1', 'eo2')} if x < (?P=xa) or x > (?P=xb) else {val('ei1', 'ei2')}\n+"
        rf"xpart = np\.linspace\(0, L, N \+ 1\)\n+"
        rf"g = \[(?P<g0>{num}), None\]\n+"
        rf"f_nodal = np\.zeros\(N \+ 1\)\n+"
        rf"dt = (?P<dt>{num})\n+"
        rf"T = (?P<T>{num})\n+"
        rf"(?P<sys>fosys = heatlinfem\.heatlinfem_mix|sosys = mechlinfem\.mechlinfem)\(xpart, rho, E, g, f_nodal, dt\)\n+"
        rf"(?P<apcs>(?:apc\d = logic\.APCont\(\[{num}, {num}\], \"[<>=]\", lambda x: {num} \* x \+ {num}, lambda x: {num}\)\n+){{1,8}})"
        rf"cregions = \{{(?P<regs>[^}}\n]*)\}}\n+"
        rf"cspec = \"(?P<spec>[^\"\n]*)\"\s*"
    )
    return re.compile(pat)


_FAST = _fast_regex()
# an integer literal with a leading zero ("02298576") is a SyntaxError for Python (2 and 3 alike; "0123.5", "1e-06" and
# "00" are fine): the template matcher must not accept what the interpreter of the dialect rejects
_ZERO_DIGIT = re.compile(r"[^\w.]0\d")
_LEADING_ZERO_INT = re.compile(r"(?<![\w.])(?<![eE][-+])0\d*[1-9]\d*(?![\d.eE])")
_FAST_APC = re.compile(rf"apc(\d) = logic\.APCont\(\[({_RE_NUM}), ({_RE_NUM})\], \"([<>=])\", lambda x: ({_RE_NUM}) \* x \+ ({_RE_NUM}), lambda x: {_RE_NUM}\)")
_FAST_JOIN = re.compile(r"\)\) ([&|]) \(")
_FAST_CLAUSE = re.compile(rf"\(([GF])_\[({_RE_NUM}), ({_RE_NUM})\] \(([A-H])\)\)")


def _parse_fast(src: str) -> Optional[Problem]:
    """The canonical template, or None (then the general interpreter decides)."""
    if len(src) > 8192:
        return None
    m = _FAST.fullmatch(src)
    if m is None:
        return None
    # leading-zero integers: the lookbehind scan costs more than the template match, so it only runs on texts that hold
    # a zero followed by a digit after a non-digit at all (a cheap necessary condition)
    if _ZERO_DIGIT.search(src) and _LEADING_ZERO_INT.search(src):
        return None
    heat = m["sys"].startswith("fosys")
    if m["imod"] is not None and (m["imod"] == "heatlinfem") != heat:
        return None
    f = float
    N = int(m["N"])
    L, dt, T, g0, x_a, x_b = f(m["L"]), f(m["dt"]), f(m["T"]), f(m["g0"]), f(m["xa"]), f(m["xb"])
    prod = lambda a, b: f(m[a]) * f(m[b]) if m[b] is not None else f(m[a])
    rho_out, rho_in, E_out, E_in = prod("ro1", "ro2"), prod("ri1", "ri2"), prod("eo1", "eo2"), prod("ei1", "ei2")
    # the checks the general route makes on these values; when one fails, let IT word the error
    if not (2 <= N <= abi.MAX_ELEM) or not (L > 0 and math.isfinite(L)) or not (dt > 0 and T > 0 and math.isfinite(dt) and math.isfinite(T)):
        return None
    if not all(v > 0 and math.isfinite(v) for v in (rho_in, rho_out, E_in, E_out)) or not math.isfinite(g0):
        return None
    apcs = _FAST_APC.findall(m["apcs"])
    P = len(apcs)
    labels = "ABCDEFGH"[:P]
    if [int(a[0]) for a in apcs] != list(range(P)):
        return None
    if m["regs"].strip() != ", ".join(f'"{labels[j]}" : apc{j}' for j in range(P)):
        return None
    # cspec: (C0) | ((C0) & (C1)) | (((C0) & (C1)) & (C2)) ... with the labels in order
    spec = m["spec"]
    cl = _FAST_CLAUSE.findall(spec)
    if len(cl) != P or "".join(c[3] for c in cl) != labels:
        return None
    parts = [f"({c[0]}_[{c[1]}, {c[2]}] ({c[3]}))" for c in cl]
    # the joins, in order; the structure is then CHECKED by rebuilding the text: left-nested with those operators
    # ((A op B) op C), or flat with "&" only -- the spelling of the reference's records: ((A) & (B) & (C)), ((A))
    ops = _FAST_JOIN.findall(spec)
    if len(ops) != P - 1:
        return None
    nested = parts[0]
    for q, op in zip(parts[1:], ops):
        nested = f"({nested} {op} {q})"
    flat = "(" + " & ".join(parts) + ")"
    if spec != nested and not (spec == flat and all(op == "&" for op in ops)):
        return None
    preds, tree = [], []
    for j, (a, c) in enumerate(zip(apcs, cl)):
        x_lo, x_hi, t_lo, t_hi = f(a[1]), f(a[2]), f(c[1]), f(c[2])
        if x_hi < x_lo or not (t_hi >= t_lo) or t_lo < 0 or not all(map(math.isfinite, (x_lo, x_hi, f(a[4]), f(a[5])))):
            return None
        preds.append(Predicate(x_lo, x_hi, a[3], f(a[4]), f(a[5]), c[0], t_lo, t_hi, labels[j]))
        tree.append(("leaf", j, 0))
        if j >= 1:
            tree.append(("and" if ops[j - 1] == "&" else "or", 0 if j == 1 else len(tree) - 2, len(tree) - 1))
    return Problem("heat" if heat else "wave", N, L, dt, T, g0, x_a, x_b, rho_in, rho_out, E_in, E_out, preds, tree)


def parse_program(text: str, fast: Optional[bool] = None) -> Problem:
    """Program text -> Problem.  The ONLY exception this raises is ParseError: whatever a malformed program makes the
    interpreter below throw (TypeError from arithmetic on odd operands inside a material lambda, an unhashable
    relation, RecursionError, OverflowError ...) is converted, so that batch callers can label the sample -100
    (simulate_gt.py:59 of the reference) and go on.  See _parse_program for the dialect."""
    try:
        return _parse_program(text, fast)
    except ParseError:
        raise
    except (RecursionError, MemoryError) as e:
        raise ParseError(f"program too deep or too large: {type(e).__name__}")
    except Exception as e:
        raise ParseError(f"cannot interpret program: {e!r}")


def _parse_program(text: str, fast: Optional[bool] = None) -> Problem:
    """Interpret one femformal-dialect program.  Raises ParseError on anything outside the dialect.

    fast: None = try the canonical-template matcher first, then the general interpreter; False = general interpreter
    only; True = template matcher only (ParseError if the text is not the template; used by the tests)."""
    if not isinstance(text, str):
        raise ParseError("program is not text")
    src = text.replace("\x00", "") if "\x00" in text else text
    # common leading whitespace is removed only where there is any (dedent is a quarter of the template route's time)
    src = (textwrap.dedent(src) if (src[:1] in " \t" or "\n " in src or "\n\t" in src) else src).strip()
    if fast is None or fast:
        p = _parse_fast(src)
        if p is not None:
            return p
        if fast:
            raise ParseError("not the canonical template")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)     # generated text: "invalid decimal literal" and the like
            mod = ast.parse(src)
    except (SyntaxError, ValueError, RecursionError, MemoryError) as e:
        raise ParseError(f"not parseable: {e}")
    env: Dict[str, object] = {}
    try:
        for stmt in mod.body:
            if isinstance(stmt, (ast.Import, ast.ImportFrom, ast.Pass)):
                continue
            if isinstance(stmt, ast.Expr):
                continue                     # docstrings, stray prints: no effect on the problem
            if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1:
                tgt = stmt.targets[0]
                if isinstance(tgt, ast.Name):
                    env[tgt.id] = _eval(stmt.value, env)
                    continue
                if isinstance(tgt, ast.Tuple) and all(isinstance(t, ast.Name) for t in tgt.elts):
                    vals = _eval(stmt.value, env)
                    if not isinstance(vals, list) or len(vals) != len(tgt.elts):
                        raise ParseError("tuple assignment mismatch")
                    for t, v in zip(tgt.elts, vals):
                        env[t.id] = v
                    continue
            raise ParseError(f"statement not part of the dialect: {type(stmt).__name__}")
    except ParseError:
        raise
    except RecursionError:
        raise ParseError("expression too deep")
    except Exception as e:               # arithmetic on odd operands etc.
        raise ParseError(f"cannot evaluate: {e!r}")

    systems = [v for k, v in env.items() if isinstance(v, _System)]
    if len(systems) != 1:
        raise ParseError("program must build exactly one system (fosys / sosys)")
    sysm = systems[0]
    a = sysm.args
    xpart = a["xpart"]
    if not isinstance(xpart, np.ndarray) or xpart.ndim != 1 or len(xpart) < 3:
        raise ParseError("xpart must be a 1-D mesh")
    N = len(xpart) - 1
    h = np.diff(xpart)
    if xpart[0] != 0.0 or not np.all(h > 0) or np.max(np.abs(h - h.mean())) > 1e-9 * h.mean():
        raise ParseError("mesh must be uniform and start at 0")
    if N > abi.MAX_ELEM:
        raise ParseError("mesh too fine for this build")
    L = float(xpart[-1])
    g = a["g"]
    if not isinstance(g, list) or len(g) != 2 or g[1] is not None or g[0] is None:
        raise ParseError("g must be [g0, None]")
    g0 = _num(g[0], "g0")
    f_nodal = a["f_nodal"]
    if not isinstance(f_nodal, np.ndarray) or f_nodal.shape != xpart.shape or np.any(f_nodal != 0.0):
        raise ParseError("f_nodal must be zeros(N + 1)")
    dt = _num(a["dt"], "dt")
    if "T" not in env:
        raise ParseError("horizon T is missing")
    T = _num(env["T"], "T")
    if not (dt > 0 and T > 0 and math.isfinite(dt) and math.isfinite(T)):
        raise ParseError("dt and T must be positive")
    x_a, x_b, rho_in, rho_out, E_in, E_out = _material(a["rho"], a["E"], xpart)
    for v in (rho_in, rho_out, E_in, E_out):
        if not (v > 0 and math.isfinite(v)):
            raise ParseError("material values must be positive")

    regions = env.get("cregions")
    spec = env.get("cspec")
    if not isinstance(regions, dict) or not regions:
        raise ParseError("cregions is missing")
    clauses, tree = parse_cspec(spec)
    preds = []
    for top, t_lo, t_hi, label in clauses:
        ap = regions.get(label)
        if not isinstance(ap, _APCont):
            raise ParseError(f"cspec names unknown region {label!r}")
        A = ap.A
        if not isinstance(A, list) or len(A) != 2:
            raise ParseError("APCont interval must be [x_lo, x_hi]")
        x_lo, x_hi = _num(A[0], "x_lo"), _num(A[1], "x_hi")
        if x_hi < x_lo:
            raise ParseError("APCont interval is reversed")
        if ap.r not in abi.REL_CODE:
            raise ParseError(f"unknown relation {ap.r!r}")
        m, b = _affine(ap.p, f"region {label}")
        preds.append(Predicate(x_lo, x_hi, ap.r, m, b, top, t_lo, t_hi, label))
    return Problem(sysm.kind, N, L, dt, T, g0, x_a, x_b, rho_in, rho_out, E_in, E_out, preds, tree)


# ---------------------------------------------------------------------------
# Problem -> program text (round trips through parse_program; used by tests and data tools)
# ---------------------------------------------------------------------------
def render_program(p: Problem) -> str:
    heat = p.kind == "heat"
    mod, fn, sysname = ("heatlinfem", "heatlinfem_mix", "fosys") if heat else ("mechlinfem", "mechlinfem", "sosys")
    labels = []
    lines = [
        f"from femformal.core.fem import {mod} as {mod}", "",
        f"N = {p.n_elem}", f"L = {p.L!r}",
        f"rho = lambda x: {p.rho_out!r} if x < {p.x_a!r} or x > {p.x_b!r} else {p.rho_in!r}",
        f"E = lambda x: {p.E_out!r} if x < {p.x_a!r} or x > {p.x_b!r} else {p.E_in!r}",
        "xpart = np.linspace(0, L, N + 1)", f"g = [{p.g0!r}, None]", "f_nodal = np.zeros(N + 1)", f"dt = {p.dt!r}", "",
        f"T = {p.T!r}", f"{sysname} = {mod}.{fn}(xpart, rho, E, g, f_nodal, dt)", "",
    ]
    for j, q in enumerate(p.predicates):
        lab = chr(ord("A") + j)
        labels.append(lab)
        lines.append(f'apc{j} = logic.APCont([{q.x_lo!r}, {q.x_hi!r}], "{q.rel}", lambda x: {q.m!r} * x + {q.b!r}, '
                     f"lambda x: {q.m!r})")
    lines.append("cregions = {" + ", ".join(f'"{l}" : apc{j}' for j, l in enumerate(labels)) + "}")
    tree = p.tree if p.tree else p.default_tree()
    txt = []
    for op, l, r in tree:
        if op == "leaf":
            q = p.predicates[l]
            txt.append(f"({q.top}_[{q.t_lo!r}, {q.t_hi!r}] ({labels[l]}))")
        else:
            txt.append(f"({txt[l]} {'&' if op == 'and' else '|'} {txt[r]})")
    lines.append(f'cspec = "{txt[-1]}"')
    return "\n".join(lines) + "\n"


def parse_native(texts, pmax: int = abi.MAX_PRED, nthreads: int = 0):
    """The library's native matcher of the canonical template (pdeb_parse_programs, csrc/pdeb_frontend.cpp) over a list
    of program texts: -> (ProblemBatch of len(texts) samples with the accepted programs' numbers in place, accepted mask).

    ~1 us per program on every host thread instead of ~85 us (the Python template matcher) or ~1 ms (the interpreter), and
    no Python object per program.  A text it declines is NOT an error: the caller parses that one with parse_program.
    What it accepts it reads exactly as _parse_fast does (tests/test_frontend_native.py).  Host-only: no CUDA involved."""
    import ctypes as C
    from . import native
    from .problem import ProblemBatch
    L = native.lib()
    n = len(texts)
    batch = ProblemBatch(n, 1, pmax, 2 * pmax - 1)
    accepted = np.zeros(n, dtype=np.uint8)
    if n == 0:
        return batch, accepted.astype(bool)
    # ONE byte buffer for the whole list (a join and an encode at memcpy speed) and a pointer table formed in numpy; the
    # per-text route below (one bytes object and one ctypes conversion per program) cost more than the parse itself
    blob, lens = None, None
    try:
        if n and isinstance(texts[0], (bytes, bytearray)):
            # a caller that holds its programs as bytes (a file read in binary mode, a socket): no encode at all
            lens = np.fromiter(map(len, texts), dtype=np.int64, count=n)
            blob = b"\0".join(texts)
            joined = None
        else:
            joined = "\0".join(texts)
        if joined is None:
            pass
        elif joined.isascii():
            lens = np.fromiter(map(len, texts), dtype=np.int64, count=n)          # characters == bytes
            if int(lens.sum()) + n - 1 == len(joined):
                blob = joined.encode("ascii")
    except TypeError:                                                              # an entry that is not a str
        blob = None
    if blob is not None:
        starts = np.zeros(n, dtype=np.int64)
        np.cumsum(lens[:-1] + 1, out=starts[1:])
        base = C.cast(C.c_char_p(blob), C.c_void_p).value or 0
        table = (starts + base).astype(np.uint64)
        ptrs = table.ctypes.data
    else:
        enc = [t.encode("utf-8", "surrogatepass") if isinstance(t, str) else (bytes(t) if isinstance(t, (bytes, bytearray)) else b"")
               for t in texts]
        keep = (C.c_char_p * n)(*enc)
        ptrs = C.cast(keep, C.c_void_p).value
        lens = np.fromiter(map(len, enc), dtype=np.int64, count=n)
    cb = batch.as_c()
    rc = L.pdeb_parse_programs(ptrs, lens.ctypes.data, n, C.addressof(cb), 0, accepted.ctypes.data, int(nthreads))
    if rc != 0:
        raise native.NativeError(f"pdeb_parse_programs failed ({rc})")
    return batch, accepted.astype(bool)


def _parse_chunk(texts):
    """Worker of parse_many: never raises; (problem, None) or (None, message) per text."""
    out = []
    for t in texts:
        try:
            out.append((parse_program(t), None))
        except Exception as e:          # parse_program raises ParseError only; belt and braces for the worker pool
            out.append((None, str(e)))
    return out


def parse_many(texts, workers: Optional[int] = None, min_parallel: int = 2000):
    """parse_program over a list, in worker processes for long lists.

    One program takes ~1 ms of interpreter time (half of it Python's own `ast.parse`), five orders of magnitude more than
    its evaluation on the GPU, so for the sweep sizes of the reference (a 2 M-sample corpus) the front-end, not the
    kernels, sets the wall time: the list is cut into chunks and parsed by `workers` processes (default: the CPUs of
    the box, at most 32).  The workers are forked (no re-import of the caller's main module, which "spawn" would run
    again) and only ever run this pure-Python parser: they never touch CUDA, the same contract as a forked data-loader
    worker.  Returns a list of (Problem, None) / (None, message)."""
    texts = list(texts)
    n = len(texts)
    if workers is None:
        import os
        workers = min(os.cpu_count() or 1, 32)
    if n < min_parallel or workers <= 1:
        return _parse_chunk(texts)
    import concurrent.futures as cf
    import multiprocessing as mp
    chunk = max(256, -(-n // (workers * 8)))
    parts = [texts[i:i + chunk] for i in range(0, n, chunk)]
    try:
        ctx = mp.get_context("fork")
        with cf.ProcessPoolExecutor(max_workers=workers, mp_context=ctx) as ex:
            res = list(ex.map(_parse_chunk, parts))
    except (OSError, RuntimeError, cf.process.BrokenProcessPool):
        return _parse_chunk(texts)                # a box that cannot start workers: same result, one process
    return [x for part in res for x in part]

