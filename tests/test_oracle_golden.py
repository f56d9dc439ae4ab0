"""Pin the CPU oracle against every known-answer vector the reference's own tests hold.

Vectors: tests/golden/reference_*.json, transcribed from /root/reference/test/** by
tools/transcribe_vectors.py.  Judged the way src/test_m.F90 judges them.
"""
import json
import os

import pytest

from tests import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, "reference_%s.json" % name)) as fh:
        return json.load(fh)["vectors"]


def f_eq(a: bytes, b: bytes):  # Fortran blank-padded ==
    n = max(len(a), len(b))
    return a.ljust(n, b" ") == b.ljust(n, b" ")


def len_utf8(s: bytes):  # reference len_utf8 on well-formed text
    return len(s.decode("utf-8", errors="surrogateescape"))


def check(v):
    p = bytes.fromhex(v["pattern"])
    k = v["kind"]
    if k == "match":   # test_m.F90:70-79
        return O.op_match(p, bytes.fromhex(v["text"])) == int(v["expect"])
    if k == "in":      # test_m.F90:58-66
        return O.op_in(p, bytes.fromhex(v["text"])) == int(v["expect"])
    if k == "regex":   # test_m.F90:84-100 (+ is_eqv_str :447-462): byte-for-byte, equal length
        res = O.regex(p, bytes.fromhex(v["text"]))[0]
        return res == bytes.fromhex(v["expect"])
    if k in ("prefix", "suffix"):  # test_m.F90:103-146
        lit = O.literals(p)
        if lit is None:
            return False
        got = lit[1] if k == "prefix" else lit[2]
        exp = bytes.fromhex(v["expect"])
        return len_utf8(exp) == len_utf8(got) and f_eq(exp, got)
    if k == "error":   # test_m.F90:167-186: regex(pattern, "", status=, err_msg=)
        _, _, _, _, st = O.regex(p, b"")
        return st == v["expect"] and O.error_message(st).rstrip() == O.error_message(v["expect"]).rstrip()
    if k == "validate":  # test_m.F90:46-54
        return bool(O.is_valid(p)[0]) == v["expect"]
    raise KeyError(k)


@pytest.mark.parametrize("name,expected_count", [("api", 997), ("ast", 167), ("error", 125), ("validate", 207)])
def test_reference_vectors(name, expected_count):
    vecs = load(name)
    assert len(vecs) == expected_count
    bad = [v["src"] + " " + v["kind"] + " " + repr(bytes.fromhex(v["pattern"])) for v in vecs if not check(v)]
    assert not bad, "%d/%d reference vectors fail:\n%s" % (len(bad), len(vecs), "\n".join(bad[:40]))
