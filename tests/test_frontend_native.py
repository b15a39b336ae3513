"""Native template matcher (pdeb_parse_programs, csrc/pdeb_frontend.cpp) against the Python front-end.

Contract under test: whatever the native matcher ACCEPTS, the Python template matcher accepts with identical values;
what it declines is left to the Python front-end (never an error).  Host-only: no GPU needed."""
import glob
import json
import os

import numpy as np
import pytest

from pde_controller_b200 import synth
from pde_controller_b200.frontend import ParseError, parse_native, parse_program, render_program
from test_frontend import HEAT as _HEAT

REF = "/root/reference"
# test_frontend.HEAT spells one intercept as a sum ("+ -2.5e+2 + 560"): valid for the interpreter, not the template
HEAT = _HEAT.replace("-2.5e+2 + 560", "310.0")
assert parse_program(HEAT, fast=True) == parse_program(_HEAT)


def native_problem(batch, i):
    p, _ = batch.get_problem(i)
    return p


def same_problem(a, b):
    """Problem equality up to the clause labels (the batch form does not carry them; get_problem names them A, B, ...)."""
    return a == b


def test_rendered_programs_are_accepted_and_read_identically():
    texts = []
    for wl, seed in (("heat", 2), ("wave", 3), ("mixed", 4)):
        b = synth.make_batch(700, wl, seed=seed)
        texts += [render_program(b.get_problem(s)[0]) for s in range(b.n)]
    batch, acc = parse_native(texts, nthreads=4)
    assert acc.all()
    for i in range(0, len(texts), 7):
        assert same_problem(native_problem(batch, i), parse_program(texts[i], fast=True)), i
    # threads do not change the result
    b1, a1 = parse_native(texts, nthreads=1)
    assert np.array_equal(b1.slab, batch.slab) and np.array_equal(a1, acc)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_every_reference_record_is_the_template():
    texts = []
    for f in sorted(glob.glob(os.path.join(REF, "utils/few_shot_prompts/examples/*/examples.jsonl"))):
        for line in open(f):
            rec = json.loads(line)
            texts += [rec[k] for k in sorted(rec) if "python" in k]
    assert len(texts) == 14
    batch, acc = parse_native(texts)
    assert acc.all()
    for i, t in enumerate(texts):
        assert same_problem(native_problem(batch, i), parse_program(t))


@pytest.mark.parametrize("mutate", [
    lambda t: t.replace("N = 30", "N = 030"),                         # leading-zero integer: SyntaxError for the interpreter
    lambda t: t.replace("L = 140", "L = 0140"),
    lambda t: t.replace("[11, 40]", "[011, 40]"),
    lambda t: t.replace("T = 9", "T = 9\x00"),
    lambda t: t.replace("\n", "\r"),                                    # bare CR line ends: not cleaned, declined
    lambda t: t.replace(" or x > 77", "or x > 77", 1),                  # a number fused with a keyword: the interpreter's call
    lambda t: t.replace(" else ", " else", 1),
    lambda t: t.replace("lambda x:", "lambdax:", 1),
    lambda t: t.replace('cspec = "((F_[0.5, 2.25] (A))', 'cspec = "((F_[0.5,2.25] (A))'),   # inside the strings: byte for byte
    lambda t: t.replace('cspec = "', 'cspec = "#'),
    lambda t: t.replace("L = 140", "L = - 140"),
    lambda t: "\n".join("    " + l for l in t.strip().split("\n")),   # indented: dedent is the Python route's job
    lambda t: t.replace("N = 30", "N = 15 * 2"),
    lambda t: t.replace('"A" : apc0', '"A":apc0'),
    lambda t: t.replace("x < 21 or x > 77 else 1420000", "x < 22 or x > 77 else 1420000"),   # E thresholds differ from rho's
    lambda t: t.replace("(A)) &", "(B)) &"),
    lambda t: t.replace("dt = .05", "dt = 0"),
    lambda t: t.replace("[11, 40]", "[40, 11]"),
    lambda t: t.replace("N = 30", "N = 1"),
    lambda t: t.replace("apc1 = ", "apc2 = "),
    lambda t: t + "\nprint(1)",
    lambda t: t.replace("fosys = heatlinfem.heatlinfem_mix", "sosys = mechlinfem.mechlinfem"),   # import says heat
    lambda t: "",
    lambda t: "x" * 20000,
])
def test_near_misses_are_declined_not_guessed(mutate):
    text = mutate(HEAT)
    batch, acc = parse_native([HEAT, text, HEAT])
    assert list(acc) == [True, False, True]
    assert same_problem(native_problem(batch, 2), parse_program(HEAT))


def test_spellings_the_template_allows():
    variants = [HEAT, HEAT.strip(), "\n\n" + HEAT + "\n\n  ", HEAT.replace("\n\n", "\n\n\n"),
                HEAT.replace("from femformal.core.fem import heatlinfem as heatlinfem\n", ""),
                HEAT.replace('cspec = "((F_[0.5, 2.25] (A)) & (G_[4.125, 7.5] (B)))"', 'cspec = "((F_[0.5, 2.25] (A)) | (G_[4.125, 7.5] (B)))"'),
                HEAT.replace("L = 140", "L = 1.4e2").replace("g = [305, None]", "g = [+305., None]").replace("dt = .05", "dt = 5E-2")]
    batch, acc = parse_native(variants)
    assert acc.all()
    for i, t in enumerate(variants):
        assert same_problem(native_problem(batch, i), parse_program(t, fast=True)), i
    assert native_problem(batch, 5).tree[-1][0] == "or"


def test_blanks_line_ends_and_comments_do_not_matter():
    """What Python's tokenizer does not care about, the native matcher does not care about either: runs of blanks and tabs
    between tokens (none at all where nothing would fuse), CR LF, blanks at line ends, empty lines, comment lines, a comment
    behind a statement.  Same numbers as the general interpreter reads from the very text."""
    import random
    import re
    rng = random.Random(7)
    wave = render_program(synth.make_batch(1, "wave", seed=3).get_problem(0)[0])
    heat3 = render_program(synth.make_batch(40, "heat", seed=5).get_problem(17)[0])

    def respace(text):
        out = []
        for line in text.strip().split("\n"):
            if line.startswith(("cspec", "cregions")):
                head, eq, rest = line.partition(" = ")
                out.append(head + rng.choice(["=", " =", "= ", "  =\t"]) + rest)      # the strings themselves stay as they are
                continue
            # blanks next to punctuation may go or grow; blanks between two names / numbers may only grow
            line = re.sub(r"(?<=[A-Za-z0-9_.]) (?=[A-Za-z0-9_.])", lambda m: rng.choice([" ", "  ", "\t", " \t "]), line)
            line = re.sub(r"(?<![A-Za-z0-9_. ]) | (?![A-Za-z0-9_. ])", lambda m: rng.choice(["", " ", "  ", "\t"]), line)
            # a comment behind a statement is only recognised where no quote precedes it (a '#' in a string is not one)
            tails = ["", " ", "\t "] + (["  # note", " # a < b \"quoted\""] if '"' not in line else [])
            out.append(line + rng.choice(tails))
            if rng.random() < 0.3:
                out.append(rng.choice(["", "# comment", "#", "   ", "#x = 1"]))
        return rng.choice(["\n", "\r\n"]).join(out) + rng.choice(["", "\n", "\r\n\r\n"])

    variants = [respace(src) for src in (HEAT, wave, heat3) for _ in range(60)]
    batch, acc = parse_native(variants)
    assert acc.all(), [variants[i] for i in np.nonzero(~acc)[0][:1]]
    for i, t in enumerate(variants):
        assert same_problem(native_problem(batch, i), parse_program(t)), t


def test_more_clauses_than_the_batch_has_planes_are_declined():
    p = synth.make_batch(1, "heat", seed=1).get_problem(0)[0]
    while len(p.predicates) < 3:
        p.predicates.append(p.predicates[0])
    p.tree = p.default_tree()
    t = render_program(p)
    assert parse_native([t], pmax=8)[1].all()
    assert not parse_native([t], pmax=2)[1].any()


def test_fuzz_native_never_accepts_what_python_declines_or_reads_differently():
    from hypothesis import HealthCheck, given, settings, strategies as st
    wave = render_program(synth.make_batch(1, "wave", seed=3).get_problem(0)[0])
    heat3 = render_program(synth.make_batch(40, "heat", seed=5).get_problem(17)[0])
    alphabet = "0123456789.eE+-*/()[]{}<>=:,\"' \nxlambdaNLTgdtorifelsenp_GF&|AB#\\\x00\t\r  ##"

    @st.composite
    def mutated(draw):
        text = draw(st.sampled_from([HEAT, wave, heat3]))
        for _ in range(draw(st.integers(0, 4))):
            kind = draw(st.integers(0, 3))
            i = draw(st.integers(0, max(len(text) - 1, 0)))
            if kind == 0:
                text = text[:i] + text[i + draw(st.integers(1, 6)):]
            elif kind == 1:
                text = text[:i] + draw(st.text(alphabet, max_size=6)) + text[i:]
            elif kind == 2:
                # rewrite one number: the commonest way two readers of the same text disagree
                num = draw(st.sampled_from(["0", "00", "007", "1.", ".5", "1e3", "1E-2", "-2.5", "+3", "1e", "1e+", "0x10", "1_0",
                                            "9" * 30, "1e999", "-0.0", "0.50", "05", "5e05", "٣"]))
                j = text.find(" ", i)
                text = text[:i] + num + (text[j:] if j >= 0 else "")
            else:
                lines = text.split("\n")
                a, b = draw(st.integers(0, len(lines) - 1)), draw(st.integers(0, len(lines) - 1))
                lines[a], lines[b] = lines[b], lines[a]
                text = "\n".join(lines)
        return text

    @settings(max_examples=1500, deadline=2000, suppress_health_check=[HealthCheck.too_slow])
    @given(mutated())
    def run(text):
        batch, acc = parse_native([text])
        if not acc[0]:
            return
        want = parse_program(text)                     # must not raise: the Python front-end accepts it too, with the same values
        assert same_problem(native_problem(batch, 0), want)

    run()


def test_blob_route_and_per_text_route_agree():
    """parse_native hands the whole list over as ONE byte buffer when every text is ASCII and falls back to one buffer per
    text otherwise (a non-ASCII character, an entry that is not a str): same numbers, same accepted mask."""
    src = synth.make_batch(64, "heat", seed=11)
    texts = [render_program(src.get_problem(i)[0]) for i in range(64)]
    a, acc_a = parse_native(texts)
    assert acc_a.all()
    mixed = list(texts)
    mixed[5] = mixed[5] + "# café\n"            # not ASCII: the per-text route for the whole list
    mixed[9] = None                                  # not a str: declined, never an exception
    b, acc_b = parse_native(mixed)
    assert not acc_b[9] and acc_b[[i for i in range(64) if i not in (5, 9)]].all()
    for name in ("L", "T", "dt", "E_in", "rho_out", "pred_m", "pred_tlo", "pred_rel", "tree_op"):
        x, y = getattr(a, name), getattr(b, name)
        keep = np.array([i for i in range(64) if acc_b[i]])
        assert np.array_equal(x[..., keep], y[..., keep]), name
    c, acc_c = parse_native(["", "x = 1", texts[0]])
    assert acc_c.tolist() == [False, False, True]


def test_eval_robustness_batch_assembles_controls_without_a_device(monkeypatch):
    """Host side of eval_robustness_batch: given controls of every accepted shape land in the batch's knot planes, a
    missing one follows the mode, mismatched lengths raise; the device call is replaced by a recorder."""
    from pde_controller_b200 import abi, eval_script as es, native

    src = synth.make_batch(6, "heat", seed=3)
    texts = [render_program(src.get_problem(i)[0]) for i in range(6)]
    seen = {}

    def fake(batch, **kw):
        seen["batch"] = batch
        return {"robustness": np.arange(batch.n, dtype=np.float64), "status": np.zeros(batch.n, np.int32), "clause": None}

    monkeypatch.setattr(native, "evaluate_batch", fake)
    ctrl = [(np.array([0.0, 1.0, 2.0]), np.array([5.0, 6.0, 7.0])),          # arrays
            ([0.0, 4.0], [1.0, 2.0]),                                          # lists
            None,                                                              # mode "zero": no knots
            (np.array([[0.0, 1.0]]), np.array([[3.0], [4.0]])),                # odd shapes are flattened
            ((0.0,), (9.0,)),
            (np.zeros(0), np.zeros(0))]
    monkeypatch.setattr(es, "_CONTROL_MODE", "zero")      # a missing control means zero input here (the default mode searches on the GPU)
    rob, rt, st = es.eval_robustness_batch(texts + ["not a program"], controls=ctrl + [None], return_status=True)
    b = seen["batch"]
    assert b.n == 7 and b.n_knots.tolist() == [3, 2, 0, 2, 1, 0, 0]
    assert b.knot_t[:3, 0].tolist() == [0.0, 1.0, 2.0] and b.knot_y[:3, 0].tolist() == [5.0, 6.0, 7.0]
    assert b.knot_t[:2, 1].tolist() == [0.0, 4.0] and b.knot_y[:2, 3].tolist() == [3.0, 4.0] and b.knot_y[0, 4] == 9.0
    assert rob[:6] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0] and rob[6] == abi.FAILED_ROBUSTNESS and st[6] == abi.ST_PARSE
    with pytest.raises(ValueError):
        es.eval_robustness_batch(texts[:2], controls=[([0.0, 1.0], [1.0]), ([0.0], [1.0, 2.0])])
    with pytest.raises(ValueError):
        es.eval_robustness_batch(texts[:2], controls=[None])


def test_control_arrays_form_equals_the_list_form(monkeypatch):
    """eval_robustness_batch(control_arrays=(T, Y, counts)): the bulk form of the control list (no Python object per program)
    assembles the same batch; a negative count means "no control" and follows the mode."""
    from pde_controller_b200 import eval_script as es, native

    src = synth.make_batch(40, "heat", seed=4)
    probs, ctrls = zip(*(src.get_problem(i) for i in range(40)))
    texts = [render_program(p) for p in probs]
    ctrls = list(ctrls)
    for i in (3, 17):
        ctrls[i] = None
    seen = []

    def fake(batch, **kw):
        seen.append(batch)
        return {"robustness": np.arange(batch.n, dtype=np.float64), "status": np.zeros(batch.n, np.int32), "clause": None}

    monkeypatch.setattr(native, "evaluate_batch", fake)
    monkeypatch.setattr(es, "_CONTROL_MODE", "zero")
    kmax = max(len(c[0]) for c in ctrls if c is not None)
    T, Y, K = np.zeros((40, kmax + 2)), np.zeros((40, kmax + 2)), np.full(40, -1)
    for i, c in enumerate(ctrls):
        if c is not None:
            K[i] = len(c[0]); T[i, : K[i]] = c[0]; Y[i, : K[i]] = c[1]
    a = es.eval_robustness_batch(texts, controls=ctrls, return_controls=True)
    b = es.eval_robustness_batch(texts, control_arrays=(T, Y, K), return_controls=True)
    ba, bb = seen
    assert a[0] == b[0] and ba.kmax == bb.kmax
    for name in ("n_knots", "knot_t", "knot_y", "T", "pred_m", "tree_op"):
        assert np.array_equal(getattr(ba, name), getattr(bb, name)), name
    assert b[2][3] is None and np.array_equal(b[2][5][1], ctrls[5][1])
    with pytest.raises(ValueError):
        es.eval_robustness_batch(texts, controls=ctrls, control_arrays=(T, Y, K))
    with pytest.raises(ValueError):
        es.eval_robustness_batch(texts, control_arrays=(T, Y[:, :3], K))
    with pytest.raises(ValueError):
        es.eval_robustness_batch(texts, control_arrays=(T, Y, K + 100))


def test_programs_as_bytes():
    """parse_native takes the programs as bytes too (no encode), alone or mixed with str in the per-text route."""
    src = synth.make_batch(16, "wave", seed=6)
    texts = [render_program(src.get_problem(i)[0]) for i in range(16)]
    a, acc_a = parse_native(texts)
    b, acc_b = parse_native([t.encode() for t in texts])
    c, acc_c = parse_native([t.encode() if i % 2 else t for i, t in enumerate(texts)] + ["é"])
    assert acc_a.all() and acc_b.all() and acc_c[:16].all() and not acc_c[16]
    for name in ("L", "T", "E_in", "pred_b", "pred_thi", "tree_lhs"):
        assert np.array_equal(getattr(a, name), getattr(b, name)) and np.array_equal(getattr(a, name), getattr(c, name)[..., :16]), name


def test_blank_level_mutations_never_make_the_two_front_ends_disagree():
    """Blanks, tabs, line ends, '#', form feeds, backslashes and semicolons dropped into (or blanks taken out of) canonical
    programs at random places: whatever the native matcher still accepts, the Python front-end reads identically."""
    import random
    rng = random.Random(11)
    srcs = [HEAT, render_program(synth.make_batch(1, "wave", seed=3).get_problem(0)[0]),
            render_program(synth.make_batch(40, "heat", seed=5).get_problem(17)[0])]
    ins = [" ", "  ", "\t", "#", " #", "\r", "\r\n", "\n", " # x", "\n#\n", "\f", "\v", "\\", ";"]
    texts = []
    for _ in range(8000):
        t = rng.choice(srcs)
        for _ in range(rng.randint(1, 4)):
            i = rng.randrange(len(t))
            t = (t[:i] + t[i + 1:]) if (rng.random() < 0.35 and t[i] in " \n") else (t[:i] + rng.choice(ins) + t[i:])
        texts.append(t)
    batch, acc = parse_native(texts)
    assert 50 < acc.sum() < len(texts)
    for i in np.nonzero(acc)[0]:
        assert same_problem(native_problem(batch, int(i)), parse_program(texts[int(i)])), texts[int(i)]
