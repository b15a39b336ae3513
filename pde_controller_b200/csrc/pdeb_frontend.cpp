/*
 * pdeb_frontend.cpp — native front-end for the CANONICAL program template (host only, no device involved).
 *
 * The reference's call surface is program TEXT: eval_robustness_helper(program, nl)
 * (test/PDEcontrol/evaluation/eval/eval_script.py:43-79), and every program of its records
 * (utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3, one_d_wave/examples.jsonl:1-3, ...) is the same dozen
 * lines with different numbers.  pdeb_parse_programs matches exactly that template and writes the
 * numbers straight into the arrays of a PdebBatch, several host threads at a time: ~1 us per program instead of the
 * ~85 us of the Python template matcher (pde_controller_b200/frontend.py::_parse_fast, whose regular expression this file
 * restates) and the ~1 ms of the general interpreter.
 *
 * Contract: a text this matcher ACCEPTS is accepted by the Python front-end with identical values (tests compare them on
 * rendered programs, on the reference's records and under fuzz); a text it declines is not an error: accepted[i] stays 0
 * and the caller hands that program to the Python front-end, which decides.  Nothing is executed.
 * What may differ from the canonical text is what Python's tokenizer does not care about: runs of blanks and tabs between
 * tokens (Cursor::lit), and -- in a second attempt on a cleaned copy -- CR LF, blanks at line ends, empty and comment lines,
 * comments behind statements (normalise).  The cspec / cregions STRINGS are matched byte for byte.
 */
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pdeb200.h"

namespace {

struct Tok { const char* p; int len; };

struct Cursor {
    const char* p;
    const char* end;
    /* strict: byte for byte (inside the cspec / cregions strings, whose own grammar is the Python front-end's) */
    bool lits(const char* s) {
        const size_t n = strlen(s);
        if ((size_t)(end - p) < n || memcmp(p, s, n) != 0) return false;
        p += n;
        return true;
    }
    static bool ident(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '.'; }
    /* statement level: a blank of the pattern stands for ANY run of blanks and tabs, as Python's tokenizer sees it -- an
     * empty run too, unless that would fuse two names / numbers / keywords ("lambda x", "137 or", "else 5": a blank with a
     * name character on both sides, where the text of a number or name in front of / behind the pattern counts as one) */
    bool lit(const char* s) {
        const char* q = p;
        for (const char* c = s; *c; ++c) {
            if (*c != ' ') {
                if (q >= end || *q != *c) return false;
                ++q;
                continue;
            }
            const char prev = (c == s) ? 'x' : c[-1], next = c[1] ? c[1] : 'x';
            const char* r = q;
            while (r < end && (*r == ' ' || *r == '\t')) ++r;
            if (r == q && ident(prev) && ident(next)) return false;
            q = r;
        }
        p = q;
        return true;
    }
    bool newlines() {                       /* \n+ */
        if (p >= end || *p != '\n') return false;
        while (p < end && *p == '\n') ++p;
        return true;
    }
    static bool digit(char c) { return c >= '0' && c <= '9'; }
    /* [-+]?(?:\d+(?:\.\d*)?|\.\d+)(?:[eE][-+]?\d+)? */
    bool num(Tok* t) {
        const char* s = p;
        const char* q = p;
        if (q < end && (*q == '-' || *q == '+')) ++q;
        const char* d0 = q;
        while (q < end && digit(*q)) ++q;
        if (q > d0) {
            if (q < end && *q == '.') { ++q; while (q < end && digit(*q)) ++q; }
        } else {
            if (q < end && *q == '.') {
                ++q;
                const char* f0 = q;
                while (q < end && digit(*q)) ++q;
                if (q == f0) return false;
            } else return false;
        }
        if (q < end && (*q == 'e' || *q == 'E')) {
            const char* e = q + 1;
            if (e < end && (*e == '-' || *e == '+')) ++e;
            const char* x0 = e;
            while (e < end && digit(*e)) ++e;
            if (e > x0) q = e;              /* otherwise the exponent is not part of the number */
        }
        t->p = s; t->len = (int)(q - s);
        p = q;
        return true;
    }
    bool uint_tok(Tok* t) {                 /* \d+ */
        const char* s = p;
        while (p < end && digit(*p)) ++p;
        if (p == s) return false;
        t->p = s; t->len = (int)(p - s);
        return true;
    }
};

/* an integer literal with a leading zero is a SyntaxError for the interpreter of the dialect ("02298576"): decline */
bool leading_zero_int(const Tok& t) {
    const char* s = t.p;
    int n = t.len;
    if (n > 0 && (*s == '-' || *s == '+')) { ++s; --n; }
    if (n < 2 || s[0] != '0') return false;
    bool nonzero = false;
    for (int i = 0; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;       /* '.', 'e': a float, fine */
        if (s[i] != '0') nonzero = true;
    }
    return nonzero;
}

bool to_double(const Tok& t, double* out) {
    if (t.len <= 0 || t.len > 400 || leading_zero_int(t)) return false;
    /* std::from_chars: correctly rounded like Python's float(), locale-free, ~3x faster than strtod; it takes no leading
     * '+' (the token grammar allows one) and reports out-of-range values instead of returning +-inf / 0: those rare
     * tokens go through strtod, which behaves as float() does */
    const char* b = t.p;
    const char* e = t.p + t.len;
    if (*b == '+') ++b;
    const std::from_chars_result r = std::from_chars(b, e, *out);
    if (r.ec == std::errc() && r.ptr == e) return true;
    char buf[408];
    memcpy(buf, t.p, (size_t)t.len);
    buf[t.len] = 0;
    char* end = nullptr;
    *out = strtod(buf, &end);
    return end == buf + t.len;
}

bool same(const Tok& a, const Tok& b) { return a.len == b.len && memcmp(a.p, b.p, (size_t)a.len) == 0; }

struct Parsed {
    int kind, N, P;
    double L, dt, T, g0, x_a, x_b, rho_in, rho_out, E_in, E_out;
    double xlo[PDEB_MAX_PRED], xhi[PDEB_MAX_PRED], m[PDEB_MAX_PRED], b[PDEB_MAX_PRED], tlo[PDEB_MAX_PRED], thi[PDEB_MAX_PRED];
    uint8_t rel[PDEB_MAX_PRED], top[PDEB_MAX_PRED], op[PDEB_MAX_PRED];     /* op[j]: join in front of clause j (j >= 1) */
};

/* num(?:\*num)? -> product */
bool value(Cursor& c, double* out) {
    Tok a, b;
    double x, y;
    if (!c.num(&a) || !to_double(a, &x)) return false;
    if (c.p < c.end && *c.p == '*') {
        Cursor save = c;
        ++c.p;
        if (c.num(&b)) { if (!to_double(b, &y)) return false; *out = x * y; return true; }
        c = save;                           /* "*" not followed by a number: the optional group does not match */
    }
    *out = x;
    return true;
}

/* one clause of the cspec:  (G_[t0, t1] (A))  with the expected label */
bool clause(Cursor& c, int j, Parsed* q) {
    Tok a, b;
    if (!c.lits("(")) return false;
    if (c.p >= c.end || (*c.p != 'G' && *c.p != 'F')) return false;
    q->top[j] = (*c.p == 'G') ? PDEB_TOP_G : PDEB_TOP_F;
    ++c.p;
    if (!c.lits("_[") || !c.num(&a) || !c.lits(", ") || !c.num(&b) || !c.lits("] (")) return false;
    if (c.p >= c.end || *c.p != (char)('A' + j)) return false;
    ++c.p;
    if (!c.lits("))")) return false;
    return to_double(a, &q->tlo[j]) && to_double(b, &q->thi[j]);
}

bool spec_nested(const char* s, const char* e, Parsed* q) {
    Cursor c{s, e};
    for (int k = 0; k < q->P - 1; ++k) if (!c.lits("(")) return false;
    if (!clause(c, 0, q)) return false;
    for (int j = 1; j < q->P; ++j) {
        if (!c.lits(" ")) return false;
        if (c.p >= c.end || (*c.p != '&' && *c.p != '|')) return false;
        q->op[j] = (*c.p == '&') ? PDEB_TREE_AND : PDEB_TREE_OR;
        ++c.p;
        if (!c.lits(" ") || !clause(c, j, q) || !c.lits(")")) return false;
    }
    return c.p == e;
}

bool spec_flat(const char* s, const char* e, Parsed* q) {      /* ((A) & (B) & (C)), ((A)) : conjunctions only */
    Cursor c{s, e};
    if (!c.lits("(") || !clause(c, 0, q)) return false;
    for (int j = 1; j < q->P; ++j) {
        if (!c.lits(" & ") || !clause(c, j, q)) return false;
        q->op[j] = PDEB_TREE_AND;
    }
    return c.lits(")") && c.p == e;
}

bool parse_raw(const char* text, int64_t len, Parsed* q) {
    const char* s = text;
    const char* e = text + len;
    if (memchr(text, 0, (size_t)len)) return false;                         /* NUL bytes: the Python route strips them */
    while (s < e && (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\f' || *s == '\v')) ++s;
    while (e > s && (e[-1] == ' ' || e[-1] == '\t' || e[-1] == '\n' || e[-1] == '\r' || e[-1] == '\f' || e[-1] == '\v')) --e;
    if (s > text && s[-1] != '\n') return false;                            /* an indented first line: dedent is the Python route's job */
    if (e - s > 8192 || e - s < 64) return false;
    for (const char* p = s; p + 1 < e; ++p)
        if (*p == '\n' && (p[1] == ' ' || p[1] == '\t')) return false;      /* indented lines: decline */
    Cursor c{s, e};
    int imod = -1;
    if (c.lit("from femformal.core.fem import ")) {
        if (c.lit("heatlinfem as heatlinfem")) imod = 0;
        else if (c.lit("mechlinfem as mechlinfem")) imod = 1;
        else return false;
        if (!c.newlines()) return false;
    }
    Tok tN, tL, txa, txb, txa2, txb2, tg0, tdt, tT;
    if (!c.lit("N = ") || !c.uint_tok(&tN) || !c.newlines()) return false;
    if (!c.lit("L = ") || !c.num(&tL) || !c.newlines()) return false;
    if (!c.lit("rho = lambda x: ") || !value(c, &q->rho_out) || !c.lit(" if x < ") || !c.num(&txa) || !c.lit(" or x > ") ||
        !c.num(&txb) || !c.lit(" else ") || !value(c, &q->rho_in) || !c.newlines()) return false;
    if (!c.lit("E = lambda x: ") || !value(c, &q->E_out) || !c.lit(" if x < ") || !c.num(&txa2) || !c.lit(" or x > ") ||
        !c.num(&txb2) || !c.lit(" else ") || !value(c, &q->E_in) || !c.newlines()) return false;
    if (!same(txa, txa2) || !same(txb, txb2)) return false;                  /* thresholds of rho and E textually identical */
    if (!c.lit("xpart = np.linspace(0, L, N + 1)") || !c.newlines()) return false;
    if (!c.lit("g = [") || !c.num(&tg0) || !c.lit(", None]") || !c.newlines()) return false;
    if (!c.lit("f_nodal = np.zeros(N + 1)") || !c.newlines()) return false;
    if (!c.lit("dt = ") || !c.num(&tdt) || !c.newlines()) return false;
    if (!c.lit("T = ") || !c.num(&tT) || !c.newlines()) return false;
    int heat;
    if (c.lit("fosys = heatlinfem.heatlinfem_mix")) heat = 1;
    else if (c.lit("sosys = mechlinfem.mechlinfem")) heat = 0;
    else return false;
    if (!c.lit("(xpart, rho, E, g, f_nodal, dt)") || !c.newlines()) return false;
    if (imod >= 0 && (imod == 0) != (heat == 1)) return false;
    /* numbers of the head */
    if (tN.len > 9 || leading_zero_int(tN)) return false;
    {
        char buf[16];
        memcpy(buf, tN.p, (size_t)tN.len);
        buf[tN.len] = 0;
        q->N = atoi(buf);
    }
    if (!to_double(tL, &q->L) || !to_double(txa, &q->x_a) || !to_double(txb, &q->x_b) || !to_double(tg0, &q->g0) ||
        !to_double(tdt, &q->dt) || !to_double(tT, &q->T)) return false;
    if (q->N < 2 || q->N > PDEB_MAX_ELEM) return false;
    if (!(q->L > 0) || !std::isfinite(q->L) || !(q->dt > 0) || !(q->T > 0) || !std::isfinite(q->dt) || !std::isfinite(q->T)) return false;
    const double mat[4] = {q->rho_in, q->rho_out, q->E_in, q->E_out};
    for (double v : mat) if (!(v > 0) || !std::isfinite(v)) return false;
    if (!std::isfinite(q->g0)) return false;
    q->kind = heat ? PDEB_KIND_HEAT : PDEB_KIND_WAVE;
    /* apc lines */
    int P = 0;
    while (P < PDEB_MAX_PRED) {
        Cursor save = c;
        Tok a, b, m1, b1, m2;
        if (!c.lit("apc")) { c = save; break; }
        if (c.p >= c.end || *c.p != (char)('0' + P)) return false;              /* indices 0, 1, 2 ... in order */
        ++c.p;
        if (!c.lit(" = logic.APCont([") || !c.num(&a) || !c.lit(", ") || !c.num(&b) || !c.lit("], \"")) return false;
        if (c.p >= c.end) return false;
        if (*c.p == '>') q->rel[P] = PDEB_REL_GT; else if (*c.p == '<') q->rel[P] = PDEB_REL_LT; else if (*c.p == '=') q->rel[P] = PDEB_REL_EQ; else return false;
        ++c.p;
        if (!c.lit("\", lambda x: ") || !c.num(&m1) || !c.lit(" * x + ") || !c.num(&b1) || !c.lit(", lambda x: ") || !c.num(&m2) ||
            !c.lit(")") || !c.newlines()) return false;
        if (leading_zero_int(m2)) return false;
        if (!to_double(a, &q->xlo[P]) || !to_double(b, &q->xhi[P]) || !to_double(m1, &q->m[P]) || !to_double(b1, &q->b[P])) return false;
        ++P;
    }
    if (P < 1) return false;
    q->P = P;
    /* cregions = {"A" : apc0, "B" : apc1}   (content compared after stripping) */
    if (!c.lit("cregions = {")) return false;
    const char* r0 = c.p;
    while (c.p < c.end && *c.p != '}' && *c.p != '\n') ++c.p;
    if (c.p >= c.end || *c.p != '}') return false;
    const char* r1 = c.p;
    ++c.p;
    if (!c.newlines()) return false;
    while (r0 < r1 && (*r0 == ' ' || *r0 == '\t')) ++r0;
    while (r1 > r0 && (r1[-1] == ' ' || r1[-1] == '\t')) --r1;
    {
        Cursor rc{r0, r1};
        for (int j = 0; j < P; ++j) {
            char item[16] = "\"A\" : apc0";
            item[1] = (char)('A' + j);
            item[9] = (char)('0' + j);
            if (j > 0 && !rc.lits(", ")) return false;
            if (!rc.lits(item)) return false;
        }
        if (rc.p != r1) return false;
    }
    /* cspec = "..."   then nothing but the end */
    if (!c.lit("cspec = \"")) return false;
    const char* s0 = c.p;
    while (c.p < c.end && *c.p != '"' && *c.p != '\n') ++c.p;
    if (c.p >= c.end || *c.p != '"') return false;
    const char* s1 = c.p;
    ++c.p;
    if (c.p != c.end) return false;
    if (!spec_nested(s0, s1, q) && !spec_flat(s0, s1, q)) return false;
    for (int j = 0; j < P; ++j) {
        if (q->xhi[j] < q->xlo[j] || !(q->thi[j] >= q->tlo[j]) || q->tlo[j] < 0) return false;
        if (!std::isfinite(q->xlo[j]) || !std::isfinite(q->xhi[j]) || !std::isfinite(q->m[j]) || !std::isfinite(q->b[j])) return false;
    }
    return true;
}

/*
 * What a program may differ in from the canonical text without meaning anything else to Python: CR LF line ends, blanks
 * at the end of a line, empty lines, comment lines, a comment behind a statement (only where no quote precedes the '#':
 * a '#' inside a string is not a comment).  The text is first tried as it is; only if that fails is a cleaned copy tried.
 */
bool normalise(const char* text, int64_t len, std::string* out) {
    out->clear();
    out->reserve((size_t)len);
    bool changed = false;
    const char* p = text;
    const char* e = text + len;
    /* any other control character decides things this pass does not model (a lone CR ends a line for Python's tokenizer,
     * also inside a comment; form feeds reset indentation): such a text is left to the Python front-end */
    for (const char* c = text; c < e; ++c) {
        const unsigned char u = (unsigned char)*c;
        if (u < 0x20 && u != '\n' && u != '\t' && !(u == '\r' && c + 1 < e && c[1] == '\n')) return false;
        if (u == 0x7f) return false;
    }
    while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        const char* q = le;
        if (q > p && q[-1] == '\r') { --q; changed = true; }
        const char* hash = nullptr;
        for (const char* c = p; c < q; ++c) {
            if (*c == '"' || *c == '\'') break;
            if (*c == '#') { if (c == p || c[-1] == ' ' || c[-1] == '\t') hash = c; break; }
        }
        if (hash) { q = hash; changed = true; }
        while (q > p && (q[-1] == ' ' || q[-1] == '\t')) { --q; changed = true; }
        if (q > p) { out->append(p, (size_t)(q - p)); out->push_back('\n'); }
        else if (hash) changed = true;
        p = nl ? nl + 1 : e;
    }
    return changed;
}

bool parse_one(const char* text, int64_t len, Parsed* q, std::string* scratch) {
    if (parse_raw(text, len, q)) return true;
    if (len > 16384 || !normalise(text, len, scratch)) return false;
    return parse_raw(scratch->data(), (int64_t)scratch->size(), q);
}

void store(const Parsed& q, const PdebBatch* out, int64_t s) {
    const int64_t st = out->stride;
    ((uint8_t*)out->kind)[s] = (uint8_t)q.kind;
    ((int32_t*)out->n_elem)[s] = q.N;
    ((double*)out->L)[s] = q.L;            ((double*)out->dt)[s] = q.dt;
    ((double*)out->T)[s] = q.T;            ((double*)out->g0)[s] = q.g0;
    ((double*)out->x_a)[s] = q.x_a;        ((double*)out->x_b)[s] = q.x_b;
    ((double*)out->rho_in)[s] = q.rho_in;  ((double*)out->rho_out)[s] = q.rho_out;
    ((double*)out->E_in)[s] = q.E_in;      ((double*)out->E_out)[s] = q.E_out;
    ((int32_t*)out->n_pred)[s] = q.P;
    for (int j = 0; j < q.P; ++j) {
        const int64_t a = (int64_t)j * st + s;
        ((double*)out->pred_xlo)[a] = q.xlo[j];  ((double*)out->pred_xhi)[a] = q.xhi[j];
        ((double*)out->pred_m)[a] = q.m[j];      ((double*)out->pred_b)[a] = q.b[j];
        ((double*)out->pred_tlo)[a] = q.tlo[j];  ((double*)out->pred_thi)[a] = q.thi[j];
        ((uint8_t*)out->pred_rel)[a] = q.rel[j]; ((uint8_t*)out->pred_top)[a] = q.top[j];
    }
    /* formula tree, children before parents: leaf0, then for every further clause its leaf and the join with what came before */
    ((int32_t*)out->n_tree)[s] = 2 * q.P - 1;
    ((uint8_t*)out->tree_op)[s] = PDEB_TREE_LEAF; ((int8_t*)out->tree_lhs)[s] = 0; ((int8_t*)out->tree_rhs)[s] = 0;
    for (int j = 1; j < q.P; ++j) {
        const int64_t leaf = (int64_t)(2 * j - 1) * st + s, join = (int64_t)(2 * j) * st + s;
        ((uint8_t*)out->tree_op)[leaf] = PDEB_TREE_LEAF; ((int8_t*)out->tree_lhs)[leaf] = (int8_t)j; ((int8_t*)out->tree_rhs)[leaf] = 0;
        ((uint8_t*)out->tree_op)[join] = q.op[j];
        ((int8_t*)out->tree_lhs)[join] = (int8_t)(j == 1 ? 0 : 2 * j - 2);
        ((int8_t*)out->tree_rhs)[join] = (int8_t)(2 * j - 1);
    }
}

}  // namespace

extern "C" int pdeb_parse_programs(const char* const* texts, const int64_t* lengths, int64_t n, const PdebBatch* out,
                                   int64_t first, uint8_t* accepted, int nthreads) {
    if (!texts || !lengths || !out || !accepted || n < 0 || first < 0) return PDEB_E_ARG;
    if (first + n > out->n || out->stride < out->n || out->pmax < 1 || out->tmax < 1) return PDEB_E_ARG;
    const void* req[] = {out->kind, out->n_elem, out->L, out->dt, out->T, out->g0, out->x_a, out->x_b, out->rho_in, out->rho_out,
                         out->E_in, out->E_out, out->n_pred, out->pred_xlo, out->pred_xhi, out->pred_m, out->pred_b, out->pred_tlo,
                         out->pred_thi, out->pred_rel, out->pred_top, out->n_tree, out->tree_op, out->tree_lhs, out->tree_rhs};
    for (const void* p : req) if (!p) return PDEB_E_ARG;
    auto work = [&](int64_t lo, int64_t hi) {
        std::string scratch;
        for (int64_t i = lo; i < hi; ++i) {
            Parsed q;
            accepted[i] = 0;
            if (!texts[i] || lengths[i] <= 0) continue;
            if (!parse_one(texts[i], lengths[i], &q, &scratch)) continue;
            if (q.P > out->pmax || 2 * q.P - 1 > out->tmax) continue;           /* does not fit this batch's planes: Python route */
            store(q, out, first + i);
            accepted[i] = 1;
        }
    };
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if (n < 2048 || nt == 1) { work(0, n); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& t : th) t.join();
    return 0;
}
