"""Front-end: dialect text -> Problem, no exec (Appendix A for the grammar)."""
import numpy as np
import pytest

from pde_controller_b200 import Predicate, Problem, ProblemBatch, synth
from pde_controller_b200.frontend import ParseError, parse_cspec, parse_program, render_program

HEAT = '''
from femformal.core.fem import heatlinfem as heatlinfem

N = 30
L = 140
rho = lambda x: 4.2e-06*455000000 if x < 21 or x > 77 else 5.1e-06*401000000
E = lambda x: 905000 if x < 21 or x > 77 else 1420000
xpart = np.linspace(0, L, N + 1)
g = [305, None]
f_nodal = np.zeros(N + 1)
dt = .05

T = 9
fosys = heatlinfem.heatlinfem_mix(xpart, rho, E, g, f_nodal, dt)

apc0 = logic.APCont([11, 40], "=", lambda x: 0.25 * x + 290.5, lambda x: 0.25)
apc1 = logic.APCont([90.0, 121.0], "<", lambda x: -0.031 * x + -2.5e+2 + 560, lambda x: -0.031)
cregions = {"A" : apc0, "B" : apc1}
cspec = "((F_[0.5, 2.25] (A)) & (G_[4.125, 7.5] (B)))"
'''


def test_parse_canonical_heat_program():
    p = parse_program(HEAT)
    assert (p.kind, p.n_elem, p.L, p.dt, p.T, p.g0) == ("heat", 30, 140.0, 0.05, 9.0, 305.0)
    assert (p.x_a, p.x_b) == (21.0, 77.0)
    assert p.rho_out == 4.2e-06 * 455000000 and p.rho_in == 5.1e-06 * 401000000
    assert (p.E_out, p.E_in) == (905000.0, 1420000.0)
    a, b = p.predicates
    assert (a.x_lo, a.x_hi, a.rel, a.m, a.b, a.top, a.t_lo, a.t_hi) == (11.0, 40.0, "=", 0.25, 290.5, "F", 0.5, 2.25)
    assert (b.rel, b.top) == ("<", "G") and b.m == pytest.approx(-0.031) and b.b == pytest.approx(310.0)
    assert p.tree == [("leaf", 0, 0), ("leaf", 1, 0), ("and", 0, 1)]


def test_wave_and_or_tree():
    txt = HEAT.replace("heatlinfem as heatlinfem", "mechlinfem as mechlinfem") \
              .replace("fosys = heatlinfem.heatlinfem_mix", "sosys = mechlinfem.mechlinfem") \
              .replace('cspec = "((F_[0.5, 2.25] (A)) & (G_[4.125, 7.5] (B)))"',
                       'cspec = "(((G_[0.5, 2.25] (A)) & (G_[1, 2] (B))) | (F_[4.125, 7.5] (B)))"')
    p = parse_program(txt)
    assert p.kind == "wave" and len(p.predicates) == 3
    assert p.tree == [("leaf", 0, 0), ("leaf", 1, 0), ("and", 0, 1), ("leaf", 2, 0), ("or", 2, 3)]
    assert [q.label for q in p.predicates] == ["A", "B", "B"]


def test_render_round_trip_on_synthetic_samples():
    for wl in ("heat", "wave"):
        b = synth.make_batch(64, wl, seed=4)
        for s in range(b.n):
            p, _ = b.get_problem(s)
            q = parse_program(render_program(p))
            for pq in q.predicates:
                pq.label = ""
            for pp in p.predicates:
                pp.label = ""
            assert q == p


def test_general_material_callables_are_fitted():
    txt = HEAT.replace("rho = lambda x: 4.2e-06*455000000 if x < 21 or x > 77 else 5.1e-06*401000000", "rho = lambda x: 2000.0") \
              .replace("E = lambda x: 905000 if x < 21 or x > 77 else 1420000", "E = lambda x: 1e6 if x <= 70 else 2e6")
    p = parse_program(txt)
    assert p.rho_in == p.rho_out == 2000.0
    assert (p.E_out, p.E_in) == (1e6, 2e6) and p.x_a == pytest.approx(70.0) and p.x_b == 140.0


@pytest.mark.parametrize("mutate", [
    lambda t: t.replace("T = 9", "T = __import__('os').getpid()"),
    lambda t: t.replace("T = 9", "T = open('/etc/passwd').read()"),
    lambda t: t.replace("g = [305, None]", "g = [305, 300]"),
    lambda t: t.replace("f_nodal = np.zeros(N + 1)", "f_nodal = np.ones(N + 1)"),
    lambda t: t.replace('"="', '">="'),
    lambda t: t.replace("lambda x: 0.25 * x + 290.5,", "lambda x: 0.25 * x * x + 290.5,"),
    lambda t: t.replace("(A)) &", "(Z)) &"),
    lambda t: t.replace("T = 9", ""),
    lambda t: t.replace("dt = .05", "dt = 0"),
    lambda t: t.replace("xpart = np.linspace(0, L, N + 1)", "xpart = np.array([0, 1, 5, 140])"),
    lambda t: t[: len(t) // 2],
    lambda t: "this is not a program",
    lambda t: t.replace("E = lambda x: 905000 if x < 21 or x > 77 else 1420000",
                        "E = lambda x: 1e6 if x < 20 else (2e6 if x < 60 else 3e6)"),
    lambda t: t.replace("N = 30", "N = 30\nwhile True: pass"),
    lambda t: t.replace("N = 30", "N = 10**9"),
    # exceptions other than ParseError raised BEHIND the guarded interpreter loop (advisor, round 1):
    # TypeError inside the material lambda when it is sampled, unhashable relation in APCont
    lambda t: t.replace("rho = lambda x: 4.2e-06*455000000 if", "rho = lambda x: -'a' if"),
    lambda t: t.replace('logic.APCont([11, 40], "="', "logic.APCont([11, 40], [1]"),
    lambda t: t.replace("L = 140", "L = 1e308 * 1e308"),
    lambda t: t.replace("g = [305, None]", "g = [{}, None]"),
])
def test_malformed_or_hostile_programs_raise_parse_error(mutate):
    with pytest.raises(ParseError):
        parse_program(mutate(HEAT))


def test_parse_many_turns_every_failure_into_a_message():
    """One malformed program must not abort the batch (contract of eval_robustness_batch / simulate_gt)."""
    from pde_controller_b200.frontend import parse_many
    bad = HEAT.replace("rho = lambda x: 4.2e-06*455000000 if", "rho = lambda x: -'a' if")
    out = parse_many([HEAT, bad, HEAT.replace('logic.APCont([11, 40], "="', "logic.APCont([11, 40], [1]"), "", None])
    assert out[0][0] is not None and out[0][1] is None
    for prob, msg in out[1:]:
        assert prob is None and isinstance(msg, str) and msg


def test_stray_statements_are_never_executed(tmp_path):
    """Expression statements (prints, stray calls) are skipped, not run: the text is never exec'd."""
    marker = tmp_path / "pwned"
    p = parse_program(HEAT.replace("N = 30", f"N = 30\nimport os\nos.system('touch {marker}')\nprint('hi')"))
    assert p.n_elem == 30 and not marker.exists()


def test_cspec_grammar():
    cl, tr = parse_cspec("((G_[0.1, 0.2] (A)))")
    assert cl == [("G", 0.1, 0.2, "A")] and tr == [("leaf", 0, 0)]
    cl, tr = parse_cspec("G_[0,1] ((A) & (B))")          # G distributes over & exactly
    assert [c[3] for c in cl] == ["A", "B"] and tr[-1] == ("and", 0, 1)
    cl, tr = parse_cspec("(A) & F_[1e-1, 2.5E0] (B)")
    assert cl[0] == ("G", 0.0, 0.0, "A") and cl[1] == ("F", 0.1, 2.5, "B")
    for bad in ("G_[0,1] ((A) | (B))", "G_[0,1] (F_[0,1] (A))", "", "((A)", "G_[1,0] (A)", "A & & B", "G_[a,b] (A)"):
        with pytest.raises(ParseError):
            parse_cspec(bad)


def test_batch_packing_round_trip():
    p = parse_program(HEAT)
    ctrl = (np.linspace(0, 9, 10), np.arange(10.0))
    b = ProblemBatch.from_problems([p, p], [ctrl, None])
    q, c = b.get_problem(0)
    for a in q.predicates + p.predicates:
        a.label = ""
    assert q == p and np.array_equal(c[0], ctrl[0]) and np.array_equal(c[1], ctrl[1])
    assert b.get_problem(1)[1] is None
    assert b.algorithmic_bytes() == (105 + 16 * 10 + 50 * 2 + 9) + (105 + 0 + 50 * 2 + 9)


def test_fuzzed_programs_parse_or_raise_parse_error_only():
    """LLM output is untrusted text: any mutation of a valid program either parses to a Problem or raises ParseError —
    no other exception type, no execution, no hang (hypothesis; character- and line-level edits, splices of two programs)."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    wave = render_program(synth.make_batch(1, "wave", seed=3).get_problem(0)[0])
    alphabet = "0123456789.eE+-*/()[]{}<>=:,\"' \nxlambdaNLTgdtorifelsenp_GF&|AB#\\\x00"

    @st.composite
    def mutated(draw):
        text = draw(st.sampled_from([HEAT, wave]))
        for _ in range(draw(st.integers(1, 6))):
            kind = draw(st.integers(0, 4))
            i = draw(st.integers(0, max(len(text) - 1, 0)))
            if kind == 0:
                text = text[:i] + text[i + draw(st.integers(1, 12)):]
            elif kind == 1:
                text = text[:i] + draw(st.text(alphabet, max_size=12)) + text[i:]
            elif kind == 2:
                lines = text.split("\n")
                j = draw(st.integers(0, len(lines) - 1))
                del lines[j]
                text = "\n".join(lines)
            elif kind == 3:
                other = wave if text.startswith("\nfrom femformal.core.fem import heat") else HEAT
                j = draw(st.integers(0, len(other) - 1))
                text = text[:i] + other[j:]
            else:
                lines = text.split("\n")
                j, k = draw(st.integers(0, len(lines) - 1)), draw(st.integers(0, len(lines) - 1))
                lines[j], lines[k] = lines[k], lines[j]
                text = "\n".join(lines)
        return text

    @settings(max_examples=400, deadline=2000, suppress_health_check=[HealthCheck.too_slow])
    @given(mutated())
    def run(text):
        try:
            p = parse_program(text)
        except ParseError:
            return
        assert isinstance(p, Problem) and p.n_elem >= 1 and p.dt > 0 and p.T > 0 and len(p.predicates) >= 1
        # whatever the template matcher accepts, the general interpreter accepts with the same result
        assert parse_program(text, fast=False) == p

    run()


def test_parse_many_in_worker_processes_equals_serial():
    """frontend.parse_many: long lists are parsed by forked workers; same problems, same failures, same order."""
    from pde_controller_b200.frontend import parse_many
    b = synth.make_batch(600, "mixed", seed=2)
    texts = [render_program(b.get_problem(s)[0]) for s in range(b.n)]
    texts[17] = "garbage"
    texts[301] = texts[301].replace("dt = ", "dt = 0 * ")
    serial = parse_many(texts, workers=1)
    par = parse_many(texts, workers=4, min_parallel=100)
    assert len(par) == len(texts)
    assert [x[0] for x in par] == [x[0] for x in serial]
    assert par[17][0] is None and par[301][0] is None and par[17][1]
    assert sum(p is None for p, _ in par) == 2


def test_template_fast_path_equals_general_interpreter():
    """The canonical-template matcher (one regular expression, ~10x faster) and the ast interpreter agree wherever the
    matcher accepts; it accepts every left-nested synthetic formula (joins `&` and `|`) and declines other spellings."""
    taken = 0
    for wl, seed in (("heat", 2), ("wave", 3), ("mixed", 4)):
        b = synth.make_batch(400, wl, seed=seed)
        for s in range(b.n):
            prob = b.get_problem(s)[0]
            text = render_program(prob)
            general = parse_program(text, fast=False)
            fast = parse_program(text, fast=True)          # every rendered program is the template, `|` joins included
            assert fast == general
            taken += 1
    assert taken == 1200
    # a formula that is not left-nested is declined (the interpreter's cspec parser reads it)
    nested_right = HEAT.replace('cspec = "((F_[0.5, 2.25] (A)) & (G_[4.125, 7.5] (B)))"',
                                'cspec = "((F_[0.5, 2.25] (A)) | ((G_[4.125, 7.5] (B)) & (F_[0.5, 2.25] (A))))"')
    with pytest.raises(ParseError):
        parse_program(nested_right, fast=True)
    assert [t[0] for t in parse_program(nested_right).tree] == ["leaf", "leaf", "leaf", "and", "or"]
    # spellings outside the template fall through to the interpreter instead of failing
    odd = HEAT.replace("N = 30", "N = 15 * 2").replace('"A" : apc0', '"A":apc0')
    with pytest.raises(ParseError):
        parse_program(odd, fast=True)
    assert parse_program(odd) == parse_program(odd, fast=False)
    # values the general route rejects are rejected either way (the matcher declines, the interpreter words the error)
    for bad in (HEAT.replace("dt = .05", "dt = 0"), HEAT.replace("[11, 40]", "[40, 11]"), HEAT.replace("N = 30", "N = 1")):
        with pytest.raises(ParseError):
            parse_program(bad)
