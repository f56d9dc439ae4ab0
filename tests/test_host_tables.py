"""CPU-only checks of the product's HOST side (pattern compiler, literal extraction, byte-level tables).

The tables the kernels would walk are walked here by tests/table_model.py (a Python model of the
kernels) and compared with (1) the reference's own known-answer vectors and (2) the oracle on
generated patterns/texts.  No GPU, no compute call into the library.
"""
import json
import os
import random

import pytest

import forgex_b200 as fx
from forgex_b200 import _lib
from tests import oracle_lib as O
from tests.table_model import BufferPrefix, Model, NfaModel, SpanLinear, SparseIn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OPS = {"match": "match", "in": "in", "regex": "regex"}

# Reference vectors whose pattern cannot be built EAGERLY within the state cap: Forgex only
# visits a handful of their DFA states lazily.  The product reports FX_ERR_DFA_STATE_CAP for them.
EAGER_CAP_PATTERNS = {b".*a(a|b){500}c{20}", b"[ab]*a[ab]{20}"}


def load(name):
    with open(os.path.join(GOLD, "reference_%s.json" % name)) as fh:
        return json.load(fh)["vectors"]


_cache = {}


def model_for(pattern, op):
    key = (pattern, op)
    if key not in _cache:
        p = fx.Pattern(pattern, op)
        if p.status == 0 and p.info()["nfa_engine"]:
            _cache[key] = (p, NfaModel(p))           # past the eager state cap: the NFA engine answers
        else:
            _cache[key] = (p, Model(p, use_direct=(len(_cache) % 2 == 0)) if p.status == 0 else None)
    return _cache[key]


def linear_answer(pattern, text):
    p, _ = model_for(pattern, "regex")
    if p.status != 0 or p.span_tables() is None:
        return None
    f, t = SpanLinear(p).regex(text)
    return text[f - 1:t] if f > 0 and t > 0 else b""


def product_answer(kind, pattern, text):
    p, m = model_for(pattern, kind)
    if p.status != 0:
        if 1 <= p.status <= 24:       # invalid pattern: operators give .false., regex gives ''
            return False if kind != "regex" else b""
        return ("status", p.status)
    if kind == "regex":
        f, t = m.regex(text)
        return text[f - 1:t] if f > 0 and t > 0 else b""
    return m.boolean(text)


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "include", "forgex_b200.h")).read()
    for name in _lib.SYMBOLS:
        assert hasattr(lib, name), name
        assert name + "(" in hdr, "%s is bound but not declared in include/forgex_b200.h" % name
    import re
    declared = set(re.findall(r"\b(fx_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


_sparse_cache = {}


def sparse_answer(pattern, text):
    """the `.in.` answer of the sparse-start kernel's model, or None when the pattern does not take that path"""
    p, _ = model_for(pattern, "in")
    if p.status != 0 or not p.info()["sparse"]:
        return None
    if pattern not in _sparse_cache:
        _sparse_cache[pattern] = SparseIn(p)
    return _sparse_cache[pattern].boolean(text)


def test_reference_api_vectors_through_product_tables(monkeypatch):
    monkeypatch.setenv("FX_SPARSE_MAX_FIRST", "128")
    bad, capped, nsparse = [], 0, 0
    for v in load("api"):
        pat, text = bytes.fromhex(v["pattern"]), bytes.fromhex(v["text"])
        got = product_answer(v["kind"], pat, text)
        assert not isinstance(got, tuple), (v["src"], got)      # every reference vector gets an answer
        if model_for(pat, v["kind"])[0].info()["nfa_engine"]:
            assert pat in EAGER_CAP_PATTERNS, pat
            capped += 1
        exp = bytes.fromhex(v["expect"]) if v["kind"] == "regex" else v["expect"]
        if got != exp:
            bad.append("%s %s %r -> %r, expected %r" % (v["src"], v["kind"], pat, got, exp))
        if v["kind"] == "regex":
            lin = linear_answer(pat, text)
            if lin is not None and lin != exp:
                bad.append("%s regex(linear) %r -> %r, expected %r" % (v["src"], pat, lin, exp))
        if v["kind"] == "in":
            sp = sparse_answer(pat, text)
            nsparse += sp is not None
            if sp is not None and sp != exp:
                bad.append("%s in(sparse) %r -> %r, expected %r" % (v["src"], pat, sp, exp))
    assert not bad, "%d vectors fail:\n%s" % (len(bad), "\n".join(bad[:40]))
    assert capped == 4
    assert nsparse > 50


def f_eq(a, b):
    n = max(len(a), len(b))
    return a.ljust(n, b" ") == b.ljust(n, b" ")


def test_reference_literal_vectors():
    bad = []
    for v in load("ast"):
        pat = bytes.fromhex(v["pattern"])
        # the reference test builds the tree from the pattern as written (no trimming): REGEX op with
        # a pattern that has no trailing blanks is the same thing
        assert pat == pat.rstrip(b" ")
        p = fx.Pattern(pat, "regex")
        if not (p.status == 0 or p.status > 24):
            bad.append("%s invalid %r" % (v["src"], pat))
            continue
        lit = p.literals()
        got = lit[1] if v["kind"] == "prefix" else lit[2]
        exp = bytes.fromhex(v["expect"])
        if not (len(exp.decode("utf-8", "replace")) == len(got.decode("utf-8", "replace")) and f_eq(exp, got)):
            bad.append("%s %s %r -> %r, expected %r" % (v["src"], v["kind"], pat, got, exp))
    assert not bad, "\n".join(bad[:40])


def test_reference_status_and_validity_vectors():
    bad = []
    for v in load("error"):
        pat = bytes.fromhex(v["pattern"])
        r = fx.regex(pat, b"") if False else None   # regex() needs a GPU; the status comes from the compile step
        st = fx.Pattern(pat, "regex").status
        st = st if st <= 24 else 0
        if st != v["expect"] or fx.status_message(st) != O.error_message(v["expect"]):
            bad.append("%s %r -> %d, expected %d" % (v["src"], pat, st, v["expect"]))
    for v in load("validate"):
        pat = bytes.fromhex(v["pattern"])
        if fx.is_valid_regex(pat) != v["expect"]:
            bad.append("%s %r validity" % (v["src"], pat))
    assert not bad, "\n".join(bad[:40])


def test_nfa_engine_tables_against_the_oracle(monkeypatch):
    """the NFA engine's tables (patterns past the eager state cap), walked by a Python model of k_nfa_bool /
    k_nfa_regex: forced on ordinary patterns with FX_STATE_CAP=3 and compared with the oracle"""
    monkeypatch.setenv("FX_STATE_CAP", "3")
    rng = random.Random(808)
    texts = [gen_text(rng) for _ in range(60)] + [b"", b" ", b"foobar", b"abc\nabc", b"\xe3\x81\x82a\xff", b"aaaa", b"aaab"]
    used = 0
    for pat in [b"foo(bar|baz)", rb"\d{3}-\d{4}", rb"^ERROR.*timeout=\d+$", b"(a|b)*a(a|b){3}", b"a*", b"^$", b"aa[bc]", b"ab+c", "[ぁ-ん]+a".encode(),
                b"ab", b"a{1,7}", b"[/]"] + \
               [gen_pattern(rng).encode() for _ in range(50)]:
        for op in ("in", "match", "regex"):
            p = fx.Pattern(pat, op)
            if p.status != 0 or not p.info()["nfa_engine"]:
                continue
            m = NfaModel(p)
            for t in texts:
                if op == "regex":
                    assert m.regex(t) == O.regex(pat, t)[2:4], (pat, t)
                else:
                    exp = O.op_match(pat, t) if op == "match" else O.op_in(pat, t)
                    assert m.boolean(t) == bool(exp), (pat, op, t)
            used += 1
    assert used > 80, used


def test_batch_validity_and_status_vectors():
    """fx_is_valid_regex_batch: all 125 status + 207 validity vectors in ONE call each (is_valid_regex is elemental in
    the reference, forgex.F90:58-71); element i must equal the single-pattern answer"""
    err, val = load("error"), load("validate")
    pats = [bytes.fromhex(v["pattern"]) for v in err + val] + [b"", b"   ", b"a b  "]
    valid, status = fx.is_valid_regex_batch(pats)
    assert len(valid) == len(pats)
    for i, v in enumerate(err):
        assert status[i] == v["expect"] and valid[i] == (v["expect"] == 0), (v["src"], pats[i])
    for i, v in enumerate(val):
        assert bool(valid[len(err) + i]) == v["expect"], (v["src"], pats[len(err) + i])
    for i, pat in enumerate(pats):
        assert bool(valid[i]) == fx.is_valid_regex(pat)
    assert fx.is_valid_regex_batch([])[0].size == 0


def test_fortran_side_route_from_an_anchored_dfa():
    """fx_compile_from_dfa: what a Fortran host would hand over after exploring automaton%construct
    breadth-first -- the anchored code-point DFA + the three literals -- must give the same answers as compiling the
    pattern text: `.in.`, `.match.`, spans (anchored emulation and the linear span path), against the oracle"""
    rng = random.Random(31337)
    pats = [b"foo(bar|baz)", rb"\d{3}-\d{4}", "[α-ωぁ-ん]+\\s\\w{2,8}".encode(), rb"^ERROR.*timeout=\d+$", b"(a|b)*a(a|b){3}", b"a*", b"x|y+",
            rb"\s\S+", b"[^a-c]x?", "あ+い".encode(), rb"\n$", b"(|^)a", b"ab+c", b"aab*"]
    pats += [gen_pattern(rng).encode() for _ in range(60)]
    texts = [gen_text(rng) for _ in range(120)] + [b"", b" ", b"foobar", b"123-4567", b"ERROR x timeout=1\n", "αβ ab_1".encode()]
    checked = 0
    for pat in pats:
        ref = fx.Pattern(pat, "regex")
        if ref.status != 0:
            continue
        d = ref.cp_automaton()
        lits = ref.literals()
        for op in ("regex", "in", "match"):
            if op == "match" and (pat.startswith(b"^") or pat.rstrip(b" ").endswith(b"$") or pat != pat.strip(b" ")):
                continue                  # operator__match preprocesses its pattern (forgex.F90:182-190): another DFA
            direct = fx.Pattern(pat, op)
            if direct.status != 0:
                continue
            q = fx.Pattern.from_dfa(op, d["cuts"], d["delta"], d["accept"], d["q0"], lits if op != "match" else direct.literals(), pattern=pat)
            assert q.status == 0, (pat, op, q.status)
            m = Model(q)
            c = O.Compiled(pat, 1 if op == "match" else 0)
            lin = SpanLinear(q) if op == "regex" and q.span_tables() is not None else None
            for t in texts:
                if op == "regex":
                    f, to = m.regex(t)
                    of, ot = O.regex(pat, t)[2:4]
                    assert (f, to) == (of, ot), (pat, t, (f, to), (of, ot))
                    if lin is not None:
                        assert lin.regex(t) == (of, ot), (pat, t, "linear")
                else:
                    got = m.boolean(t)
                    exp = O.op_match(pat, t) if op == "match" else O.op_in(pat, t)
                    assert got == bool(exp), (pat, op, t, got, exp)
                checked += 1
    assert checked > 10000, checked


# ---- generated patterns and texts: product tables vs oracle -------------------------------------
ATOMS = ["a", "b", "c", "ab", "ba", ".", "\\d", "\\w", "\\s", "\\S", "\\D", "[ab]", "[^a]", "[a-c]", "[^a-cx]", "\\n",
         "^", "$", "x", " ", "é", "あ", "[ぁ-ん]", "[α-ω]", "\\x41", "\\x{3042}", "[\\x00-\\x20]", "\\t", "-", "}",
         "\\.", "[\\n]", "[\\d-]", "(|^)", "\\x00", "[^\\n]"]
SUFFIXES = ["", "", "", "*", "+", "?", "{2}", "{0,2}", "{1,}", "{,3}", "{0}", "{2,3}"]


def gen_pattern(rng, depth=0):
    n = rng.randint(1, 4)
    parts = []
    for _ in range(n):
        r = rng.random()
        if depth < 2 and r < 0.25:
            inner = gen_pattern(rng, depth + 1)
            if rng.random() < 0.5:
                inner += "|" + gen_pattern(rng, depth + 1)
            atom = "(" + inner + ")"
        else:
            atom = rng.choice(ATOMS)
        parts.append(atom + rng.choice(SUFFIXES))
    return "".join(parts)


TEXT_PIECES = [b"a", b"b", b"c", b"ab", b"x", b" ", b"\n", b"\r\n", b"\t", b"0", b"7", b"_", "é".encode(), "あ".encode(),
               "ん".encode(), "α".encode(), "ω".encode(), "　".encode(), b"\x00", b"\x80", b"\xbf", b"\xc3", b"\xe3\x81",
               b"\xf0\x9f\x98", b"\xff", b"\xc0\x80", b"\xc1\xa1", b"\xe0\x81\xa1", b"\xf0\x80\x81\xa1", b"\xef\xbf\xbf",
               b"\xf4\x90\x80\x81", b"\xf7\xbf\xbf\xbf", b"A", b"-", b".", b"}", b"\x1f", b"\x0b"]


def gen_text(rng):
    return b"".join(rng.choice(TEXT_PIECES) for _ in range(rng.randint(0, 8)))


@pytest.mark.parametrize("seed", range(12))
def test_generated_patterns_match_oracle(seed, monkeypatch):
    monkeypatch.setenv("FX_SPARSE_MAX_FIRST", "128")    # let every eligible pattern through the sparse-start model
    rng = random.Random(1000 + seed)
    bad = []
    checked = 0
    nsparse = 0
    nbufpre = 0
    for _ in range(120):
        pat = gen_pattern(rng).encode()
        texts = [gen_text(rng) for _ in range(12)] + [b"", b" "]
        for kind in ("match", "in", "regex"):
            p = fx.Pattern(pat, kind)
            valid = O.is_valid(pat if kind != "match" else b"x")[0]  # placeholder, real check below
            if 1 <= p.status <= 24:
                # both sides must reject: the oracle answers False / '' for an invalid pattern
                for t in texts[:2]:
                    o = O.op_match(pat, t) if kind == "match" else O.op_in(pat, t) if kind == "in" else O.regex(pat, t)[4]
                    if kind == "regex":
                        if o != p.status:
                            bad.append("status %r: product %d oracle %d" % (pat, p.status, o))
                    elif o != 0:
                        bad.append("invalid-on-product only: %s %r" % (kind, pat))
                continue
            if p.status != 0:
                continue  # cap
            m = NfaModel(p) if p.info()["nfa_engine"] else Model(p, use_direct=rng.random() < 0.5)
            span = SpanLinear(p) if kind == "regex" and p.span_tables() is not None else None
            bufpre = BufferPrefix(p) if kind == "regex" and p.info()["prefix_scan"] else None
            nbufpre += bufpre is not None
            sparse = SparseIn(p) if kind == "in" and p.info()["sparse"] else None
            nsparse += sparse is not None
            for t in texts:
                checked += 1
                if kind == "regex":
                    res, ln, f, to, st = O.regex(pat, t)
                    got = m.regex(t)
                    if st != 0 or got != (f, to):
                        bad.append("regex %r on %r: product %r oracle %r (status %d)" % (pat, t, got, (f, to), st))
                    if span is not None:
                        got2 = span.regex(t)
                        if got2 != (f, to):
                            bad.append("regex(linear) %r on %r: product %r oracle %r" % (pat, t, got2, (f, to)))
                    if bufpre is not None:
                        got3 = bufpre.regex(t)
                        if got3 != (f, to):
                            bad.append("regex(buffer, prefix) %r on %r: product %r oracle %r" % (pat, t, got3, (f, to)))
                else:
                    o = O.op_match(pat, t) if kind == "match" else O.op_in(pat, t)
                    got = m.boolean(t)
                    if o < 0 or bool(o) != got:
                        bad.append("%s %r on %r: product %r oracle %r" % (kind, pat, t, got, o))
                    if sparse is not None and sparse.boolean(t) != bool(o):
                        bad.append("in(sparse) %r on %r: product %r oracle %r" % (pat, t, sparse.boolean(t), o))
    assert not bad, "%d mismatches (of %d):\n%s" % (len(bad), checked, "\n".join(bad[:30]))
    assert checked > 1000
    assert nsparse > 5


def test_the_host_decides_correctly_which_patterns_need_the_literal_gates():
    """fx_pattern_info.gated is the launcher's decision to send every string of a boolean batch through the wrapper's
    literal gates (eval_bool_slow) instead of walking the automaton directly.  For every pattern the host leaves
    un-gated, the un-gated evaluation (degenerate texts + automaton walk: Model(gates=False)) must equal the oracle --
    on the fuzz texts and on the texts the gates exist for: each literal, its prefixes, blanks.  (The launcher once
    took `' +ab'` -- prefix literal ' ', blank but not empty -- for un-gated: `' ' .match. ' +ab'` is true by the length
    rule of do_matching_exactly, api_internal_m.F90:199-233; found by the extended GPU fuzz, not by this suite.)"""
    rng = random.Random(4242)
    fixed = [rb" +(\w{2,3}\.?){1,}", b" +ab", b"ab +", b"  +", rb" \w+ ", b"a* ", b" a*", b" ", b"  ", b" ?", b"( |ab)", b"ab", b"a b", b"ab.*cd", b"x?ab"]
    pats = fixed + [gen_pattern(rng).encode() for _ in range(400)]
    ungated = gated = 0
    bad = []
    for pat in pats:
        for kind in ("match", "in"):
            p = fx.Pattern(pat, kind)
            if p.status != 0 or p.info()["nfa_engine"]:
                continue
            all_, pre, suf = p.literals()
            if p.info()["gated"]:
                gated += 1
                continue
            ungated += 1
            m = Model(p, gates=False)
            texts = [gen_text(rng) for _ in range(8)] + [b"", b" ", b"  ", b"a", all_, pre, suf, pre[:1], suf[-1:], pre + suf, pre + b"x" + suf]
            for t in texts:
                o = O.op_match(pat, t) if kind == "match" else O.op_in(pat, t)
                if bool(o) != bool(m.boolean(t)):
                    bad.append((pat, kind, t, o))
    assert not bad, bad[:10]
    assert ungated > 100 and gated > 20, (ungated, gated)


PREFIX_HEADS = [b"foo", b"ab", b"ERROR", b"x-", b"key=", "\u3042\u3044".encode(), b"a\\.b", b"q", b"zz", b"abab"]
PREFIX_TAILS = [b"(bar|baz)", b".*end", b"[0-9]+", b"b*", b"(x|y)?z", b"\\s\\w+", b"+", b"{2,3}c", b"(|^)k", b".", b"[^ ]*$"]
PREFIX_PIECES = [b"foo", b"fo", b"foobar", b"foobaz", b"ab", b"abab", b"a", b"b", b"ERROR", b"ERR", b" end", b"end", b"x-", b"x",
                 b"key=", b"key", b"7", b"42", b" ", b"\n", b"\r\n", b"z", b"zz", b"q", b"k", "\u3042\u3044".encode(), "\u3042".encode(),
                 b"\xc1\xa6oo", b"\xc1\xa1b", b"\xe3\x81", b"\xff", b"a.b", b"a-b", b"c", b"y", b"_w"]


@pytest.mark.parametrize("seed", range(4))
def test_prefix_literal_patterns_on_the_buffer_path(seed):
    """patterns that begin with a literal: the long-buffer formulation (prefix occurrences as candidates, every
    boundary when the literal occurs nowhere) against the oracle"""
    rng = random.Random(7000 + seed)
    bad, eligible, rejected = [], 0, 0
    for _ in range(60):
        pat = rng.choice(PREFIX_HEADS) + rng.choice(PREFIX_TAILS)
        p = fx.Pattern(pat, "regex")
        if p.status != 0:
            continue
        inf = p.info()
        if not inf["prefix_scan"]:
            rejected += 1
            continue
        eligible += 1
        m = BufferPrefix(p)
        for _ in range(25):
            text = b"".join(rng.choice(PREFIX_PIECES) for _ in range(rng.randint(0, 9)))
            res, ln, f, to, st = O.regex(pat, text)
            got = m.regex(text)
            if st != 0 or got != (f, to):
                bad.append("%r on %r: model %r oracle %r" % (pat, text, got, (f, to)))
    assert not bad, "%d mismatches:\n%s" % (len(bad), "\n".join(bad[:30]))
    assert eligible > 15 and rejected > 0


def test_sparse_start_and_prefix_scan_eligibility():
    """which patterns the host sends to the sparse-start sweep (K2c / K4) and to the prefix-candidate buffer scan"""
    def inf(pat, op):
        return fx.Pattern(pat, op).info()
    i = inf(b"foo(bar|baz)", "in")
    assert (i["sparse"], i["sparse_lo"], i["sparse_hi"], i["sparse_high"], i["sparse_second"], i["prefix_mode"]) == \
        (1, [0x66], [0x66], 1, 0x6F, 1)           # 'f', then only 'o' (or a lead byte: overlong forms) can follow
    i = inf(rb"^ERROR.*timeout=\d+$", "in")
    assert (i["sparse"], i["sparse_lo"], i["sparse_hi"], i["sparse_second"]) == (1, [0, 10, 13], [0, 10, 13], -1)
    assert inf(rb"\w+@\w+", "in")["sparse"] == 0       # 63 possible first bytes: not sparse
    assert inf(b"[^a]b", "in")["sparse"] == 0          # a stray continuation byte (U+FFFF) can start a match
    assert inf(b"^", "in")["sparse"] == 0              # the leading NUL alone already accepts
    assert inf(b"(a|b)*a(a|b){12}", "in")["sparse"] == 0   # anchored automaton above the optional-build cap
    assert inf(b"foobar", "in")["sparse"] == 0         # literal-only: the automaton is never consulted
    for pat, ok in [(rb"ERROR.*timeout=\d+", 1), (b"foo(bar|baz)", 1), (b"key=[0-9]*", 1), ("\u3042\u3044+".encode(), 1),
                    (b"ab+c", 0), (b"aab*", 0), (b"abab+", 0), (b"fo+bar", 0), (rb"^ERROR.*timeout=\d+$", 0), (b"[a-z]+", 0)]:
        assert inf(pat, "regex")["prefix_scan"] == ok, pat


def test_value_forms_for_pure_callers():
    """fx_in_value / fx_match_value / fx_regex_sub (the shapes a `pure` Fortran operator can bind): the
    boolean comes back as the function value, a failure as -status; an invalid pattern is `false` / status in the
    regex form, exactly as fx_in / fx_match / fx_regex answer it.  Without a device the matching itself answers
    FX_ERR_NO_DEVICE (the product has no CPU fallback)."""
    import ctypes as C
    import torch
    from forgex_b200 import _lib as L
    lib = L.lib()
    assert lib.fx_in_value(b"(", 1, b"abc", 3) == 0 and lib.fx_match_value(b"a{2,1}", 6, b"aa", 2) == 0    # invalid pattern: false
    f, t, ln, st, rc = C.c_int64(7), C.c_int64(7), C.c_int64(7), C.c_int(7), C.c_int(7)
    lib.fx_regex_sub(b"(", 1, b"abc", 3, C.byref(f), C.byref(t), C.byref(ln), C.byref(st), C.byref(rc))
    assert rc.value == 0 and st.value == O.regex(b"(", b"abc")[4] != 0 and ln.value == 0
    have = torch.cuda.is_available()
    assert lib.fx_in_value(b"ab", 2, b"xaby", 4) == (1 if have else -L.FX_ERR_NO_DEVICE)
    assert lib.fx_match_value(b"ab", 2, b"xaby", 4) == (0 if have else -L.FX_ERR_NO_DEVICE)
    lib.fx_regex_sub(b"ab", 2, b"xaby", 4, C.byref(f), C.byref(t), C.byref(ln), C.byref(st), C.byref(rc))
    if have:
        assert (rc.value, st.value, f.value, t.value, ln.value) == (0, 0, 2, 3, 2)
    else:
        assert rc.value == L.FX_ERR_NO_DEVICE
