#!/usr/bin/env python3
"""Transcribe the reference's known-answer Fortran test programs into JSON golden vectors.

Runs in the build container only (reads /root/reference/test/**.f90, read-only); the GPU
box never needs it -- the output under tests/golden/ is committed.

The reference tests are stand-alone Fortran programs made of `call runner_*(...)` lines
(/root/reference/src/test_m.F90:210-420) plus a little string plumbing (character
variables, `//`, char(), repeat(), trim(), nchar(mask), char_utf8(), one do-loop, one goto).
This is a tiny interpreter for exactly that subset.  Every byte string is stored as hex
because the vectors contain invalid UTF-8 on purpose (test_case_010.f90).

Output records: {"kind": match|in|regex|prefix|suffix|error|validate, "pattern": hex,
"text": hex, "expect": bool|hex|int, "src": "file:line"}.
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# /root/reference/src/essential/utf8_m.f90:32-37
MASKS = {"fullbit": -1, "ascii_mask": 127, "lead_2_mask": -33, "lead_3_mask": -17,
         "lead_4_mask": -9, "continuation_mask": -65}

# /root/reference/src/essential/error_m.F90:12-38 (enum, bind(c): consecutive from 0)
ERR_NAMES = ["SYNTAX_VALID", "SYNTAX_ERR", "SYNTAX_ERR_PARENTHESIS_MISSING",
             "SYNTAX_ERR_PARENTHESIS_UNEXPECTED", "SYNTAX_ERR_BRACKET_MISSING",
             "SYNTAX_ERR_BRACKET_UNEXPECTED", "SYNTAX_ERR_CURLYBRACE_MISSING",
             "SYNTAX_ERR_CURLYBRACE_UNEXPECTED", "SYNTAX_ERR_INVALID_TIMES",
             "SYNTAX_ERR_ESCAPED_SYMBOL_MISSING", "SYNTAX_ERR_ESCAPED_SYMBOL_INVALID",
             "SYNTAX_ERR_EMPTY_CHARACTER_CLASS", "SYNTAX_ERR_RANGE_WITH_ESCAPE_SEQUENCES",
             "SYNTAX_ERR_MISPLACED_SUBTRACTION_OPERATOR", "SYNTAX_ERR_INVALID_CHARACTER_RANGE",
             "SYNTAX_ERR_CHAR_CLASS_SUBTRANCTION_NOT_IMPLEMENTED", "SYNTAX_ERR_STAR_INCOMPLETE",
             "SYNTAX_ERR_PLUS_INCOMPLETE", "SYNTAX_ERR_QUESTION_INCOMPLETE",
             "SYNTAX_ERR_INVALID_HEXADECIMAL", "SYNTAX_ERR_HEX_DIGITS_NOT_ENOUGH",
             "SYNTAX_ERR_UNICODE_EXCEED", "SYNTAX_ERR_UNICODE_PROPERTY_NOT_IMPLEMENTED",
             "SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN", "ALLOCATION_ERR"]
ERR = {n: i for i, n in enumerate(ERR_NAMES)}


def char_utf8(code):
    # standard UTF-8 encoder; equals the reference's char_utf8 for every code the tests use
    return chr(code).encode("utf-8")


class Tok:
    def __init__(self, s):
        self.s = s  # bytes
        self.i = 0

    def ws(self):
        while self.i < len(self.s) and self.s[self.i:self.i + 1] in b" \t":
            self.i += 1

    def peek(self):
        self.ws()
        return self.s[self.i:self.i + 1]


class Interp:
    def __init__(self, path):
        self.path = path
        self.vars = {}        # name -> bytes | int
        self.fixed = {}       # name -> fixed length for character(N) variables
        self.out = []

    # ---- expression parser over bytes -------------------------------------------------
    def expr(self, t):
        v = self.term(t)
        while True:
            t.ws()
            if t.s[t.i:t.i + 2] == b"//":
                t.i += 2
                r = self.term(t)
                v = v + r
            elif t.peek() in (b"+", b"-") and isinstance(v, int):
                op = t.peek()
                t.i += 1
                r = self.term(t)
                v = v + r if op == b"+" else v - r
            else:
                return v

    def term(self, t):
        c = t.peek()
        if c in (b"'", b'"'):
            q = c
            t.i += 1
            buf = bytearray()
            while True:
                ch = t.s[t.i:t.i + 1]
                if ch == b"":
                    raise ValueError("unterminated string in %s" % self.path)
                if ch == q:
                    if t.s[t.i + 1:t.i + 2] == q:   # doubled quote
                        buf += q
                        t.i += 2
                        continue
                    t.i += 1
                    return bytes(buf)
                buf += ch
                t.i += 1
        if c == b".":
            for lit, val in ((b".true.", True), (b".false.", False)):
                if t.s[t.i:t.i + len(lit)].lower() == lit:
                    t.i += len(lit)
                    return val
        m = re.match(rb"[+-]?\d+", t.s[t.i:])
        if m:
            t.i += m.end()
            return int(m.group())
        m = re.match(rb"[A-Za-z_][A-Za-z0-9_]*", t.s[t.i:])
        if not m:
            raise ValueError("cannot parse %r in %s" % (t.s[t.i:], self.path))
        name = m.group().decode()
        t.i += m.end()
        if t.peek() == b"(":
            t.i += 1
            args = []
            if t.peek() != b")":
                while True:
                    args.append(self.expr(t))
                    if t.peek() == b",":
                        t.i += 1
                        continue
                    break
            assert t.peek() == b")", (t.s, self.path)
            t.i += 1
            return self.call(name.lower(), args)
        if name in ERR:
            return ERR[name]
        if name in MASKS:
            return MASKS[name]
        if name.lower() in self.vars:
            return self.vars[name.lower()]
        raise KeyError("unknown name %s in %s" % (name, self.path))

    def call(self, f, a):
        if f in ("char", "achar"):
            return bytes([a[0]])
        if f == "nchar":                      # /root/reference/src/test_m.F90:412-422
            return bytes([a[0] + 256 if a[0] < 0 else a[0]])
        if f == "ichar":
            return a[0][0]
        if f == "repeat":
            return a[0] * a[1]
        if f == "trim":
            return a[0].rstrip(b" ")
        if f == "char_utf8":
            return char_utf8(a[0])
        raise KeyError("unknown function %s in %s" % (f, self.path))

    # ---- statements ------------------------------------------------------------------
    def run(self):
        raw = open(self.path, "rb").read().split(b"\n")
        # join continuation lines ('&' at end, optional '&' at start of the next line)
        stmts = []
        cur, cur_line = None, 0
        for ln, line in enumerate(raw, 1):
            s = self.strip_comment(line).rstrip()
            if cur is not None:
                s2 = s.lstrip()
                if s2.startswith(b"&"):
                    s2 = s2[1:]
                else:
                    s2 = s.lstrip()
                s = cur + s2
                cur = None
            else:
                cur_line = ln
            if s.rstrip().endswith(b"&"):
                cur = s.rstrip()[:-1]
                continue
            if s.strip():
                stmts.append((cur_line, s.strip()))
        self.exec_block(stmts)
        return self.out

    @staticmethod
    def strip_comment(line):
        q = None
        for i in range(len(line)):
            ch = line[i:i + 1]
            if q:
                if ch == q:
                    q = None
            elif ch in (b"'", b'"'):
                q = ch
            elif ch == b"!":
                return line[:i]
        return line

    def exec_block(self, stmts):
        pc = 0
        while pc < len(stmts):
            ln, s = stmts[pc]
            low = s.lower()
            m = re.match(rb"goto\s+(\d+)", low)
            if m:
                label = m.group(1)
                while not re.match(rb"%s\s+continue" % label, stmts[pc][1].lower()):
                    pc += 1
                pc += 1
                continue
            m = re.match(rb"do\s+(\w+)\s*=\s*(.+)", s, re.I)
            if m and not low.startswith(b"do while"):
                var = m.group(1).decode().lower()
                t = Tok(m.group(2))
                lo = self.expr(t)
                assert t.peek() == b","
                t.i += 1
                hi = self.expr(t)
                depth, end = 1, pc + 1
                while depth:
                    l2 = stmts[end][1].lower()
                    if re.match(rb"do\s", l2):
                        depth += 1
                    if re.match(rb"end\s*do", l2):
                        depth -= 1
                    end += 1
                body = stmts[pc + 1:end - 1]
                for v in range(lo, hi + 1):
                    self.vars[var] = v
                    self.exec_block(body)
                pc = end
                continue
            m = re.match(rb"character\((\d+)\)\s*::\s*(.+)", s, re.I)
            if m:
                for name in m.group(2).split(b","):
                    self.fixed[name.strip().decode().lower()] = int(m.group(1))
                pc += 1
                continue
            m = re.match(rb"call\s+runner_(\w+)\s*\((.*)\)\s*$", s, re.I)
            if m:
                self.runner(m.group(1).decode().lower(), m.group(2), ln)
                pc += 1
                continue
            m = re.match(rb"(\w+)\s*=\s*(.+)$", s)
            if m and not re.match(rb"(logical|integer|character|if|print|write|use|implicit|program)\b", low):
                name = m.group(1).decode().lower()
                val = self.expr(Tok(m.group(2)))
                if name in self.fixed and isinstance(val, bytes):
                    n = self.fixed[name]
                    val = (val + b" " * n)[:n]
                self.vars[name] = val
            pc += 1

    def runner(self, kind, argstr, ln):
        t = Tok(argstr)
        args = []
        while True:
            args.append(self.expr(t))
            if t.peek() == b",":
                t.i += 1
                if re.match(rb"\s*res\s*$", t.s[t.i:]):
                    break
                continue
            break
        src = "%s:%d" % (os.path.relpath(self.path, REF), ln)
        hx = lambda b: b.hex()
        if kind in ("match", "in"):
            rec = {"kind": kind, "pattern": hx(args[0]), "text": hx(args[1]), "expect": bool(args[2])}
        elif kind == "regex":
            rec = {"kind": kind, "pattern": hx(args[0]), "text": hx(args[1]), "expect": hx(args[2])}
        elif kind in ("prefix", "suffix"):
            rec = {"kind": kind, "pattern": hx(args[0]), "expect": hx(args[1])}
        elif kind == "error":
            rec = {"kind": kind, "pattern": hx(args[0]), "text": hx(args[1]), "expect": int(args[2]),
                   "expect_name": ERR_NAMES[int(args[2])]}
        elif kind == "validate":
            rec = {"kind": kind, "pattern": hx(args[0]), "expect": bool(args[1])}
        else:
            raise KeyError(kind)
        rec["src"] = src
        self.out.append(rec)


def main():
    os.makedirs(OUT, exist_ok=True)
    groups = {"api": "test/test_api", "ast": "test/test_ast", "error": "test/test_error",
              "validate": "test/test_invalid_patterns"}
    total = 0
    for name, d in groups.items():
        recs = []
        for f in sorted(os.listdir(os.path.join(REF, d))):
            if f.endswith(".f90"):
                recs += Interp(os.path.join(REF, d, f)).run()
        with open(os.path.join(OUT, "reference_%s.json" % name), "w") as fh:
            json.dump({"source": "transcribed from /root/reference/%s by tools/transcribe_vectors.py" % d,
                       "vectors": recs}, fh, indent=0, separators=(",", ":"))
            fh.write("\n")
        kinds = {}
        for r in recs:
            kinds[r["kind"]] = kinds.get(r["kind"], 0) + 1
        print(name, len(recs), kinds)
        total += len(recs)
    print("total", total)


if __name__ == "__main__":
    sys.exit(main())
