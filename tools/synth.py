"""Deterministic synthetic inputs for the five BASELINE configs.

numpy only (host side); bench.py has the torch/device twins for the full-size runs.  Every generator
returns (buf uint8[...], offsets int64[n+1]) or (buf, n, stride) and is a pure function of (n, seed).
"""
import numpy as np

SEEDS = {"c1": 0xF06E0001, "c2": 0xF06E0002, "c3": 0xF06E0003, "c4": 0xF06E0004, "c5": 0xF06E0005}
PATTERNS = {
    "c1": rb"\d{3}-\d{4}",
    "c2": rb"foo(bar|baz)",
    "c3": "[α-ωぁ-ん]+\\s\\w{2,8}".encode("utf-8"),
    "c4": rb"^ERROR.*timeout=\d+$",
    "c5": rb"(a|b)*a(a|b){12}",
}
OPS = {"c1": "match", "c2": "in", "c3": "regex", "c4": "regex", "c5": "in"}


def rng_for(cfg, stream=0):
    return np.random.Generator(np.random.PCG64([SEEDS[cfg], stream]))


def gen_c1(n, seed_stream=0):
    """n x 8 bytes: half `ddd-dddd`, half the same with one position replaced by a printable byte that breaks it"""
    r = rng_for("c1", seed_stream)
    buf = r.integers(48, 58, size=(n, 8), dtype=np.uint8)
    buf[:, 3] = ord("-")
    bad = r.random(n) < 0.5
    pos = r.integers(0, 8, size=n)
    repl = r.integers(0x20, 0x7F, size=n, dtype=np.uint8)
    # make the replacement non-conforming: a non-digit at digit positions, a non-hyphen at position 3
    digit_pos = pos != 3
    repl = np.where(digit_pos & (repl >= 48) & (repl <= 57), repl + 17, repl).astype(np.uint8)   # '0'..'9' -> 'A'..'J'
    repl = np.where(~digit_pos & (repl == ord("-")), ord("_"), repl).astype(np.uint8)
    rows = np.nonzero(bad)[0]
    buf[rows, pos[rows]] = repl[rows]
    return buf.reshape(-1), n, 8


def gen_c2(n, seed_stream=0):
    """n lines, length U[64,256], bytes U[0x20,0x7E]; 1/16 with foobar|foobaz planted, 1/16 near misses"""
    r = rng_for("c2", seed_stream)
    lens = r.integers(64, 257, size=n)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    buf = r.integers(0x20, 0x7F, size=int(offsets[-1]), dtype=np.uint8)
    kind = r.integers(0, 16, size=n)
    where = (r.random(n) * (lens - 6)).astype(np.int64)
    words = [b"foobar", b"foobaz", b"foobax", b"fooba "]
    which = r.integers(0, 2, size=n)
    for sel, base in ((kind == 0, 0), (kind == 1, 2)):
        rows = np.nonzero(sel)[0]
        for w in (0, 1):
            rr = rows[which[rows] == w]
            lit = np.frombuffer(words[base + w], dtype=np.uint8)
            idx = (offsets[rr] + where[rr])[:, None] + np.arange(6)[None, :]
            buf[idx] = lit[None, :]
    return buf, offsets


_GREEK = [chr(c) for c in range(0x03B1, 0x03CA)]
_HIRA = [chr(c) for c in range(0x3041, 0x3094)]
_WORD = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789_"
_SEPS = [" ", "\t", "　", ","]
_BAD = [b"\x80", b"\xbf", b"\xc2", b"\xe3\x81", b"\xf0\x9f\x98", b"\xf8", b"\xff", b"\xc0\x80"]


def gen_c3(n, seed_stream=0):
    """n strings of 32..160 bytes built from Greek / hiragana / ASCII-word runs and separators; 1 % carry one
    injected malformed sequence.  Never emits F4 90 80 81."""
    r = rng_for("c3", seed_stream)
    out = []
    offsets = np.zeros(n + 1, dtype=np.int64)
    total = 0
    for i in range(n):
        target = int(r.integers(32, 161))
        parts = []
        size = 0
        while size < target:
            k = int(r.integers(0, 3))
            run = int(r.integers(1, 9))
            if k == 0:
                tok = "".join(_GREEK[j] for j in r.integers(0, len(_GREEK), size=run))
            elif k == 1:
                tok = "".join(_HIRA[j] for j in r.integers(0, len(_HIRA), size=run))
            else:
                tok = "".join(_WORD[j] for j in r.integers(0, len(_WORD), size=run))
            tok += _SEPS[int(r.integers(0, 4))]
            b = tok.encode("utf-8")
            if size + len(b) > 160:
                if size >= 32:
                    break
                b = b"x" * (32 - size)   # pad with ASCII to reach the minimum without splitting a character
            parts.append(b)
            size += len(b)
        s = b"".join(parts)
        if r.random() < 0.01:
            bad = _BAD[int(r.integers(0, len(_BAD)))]
            # inject at a token boundary so that well-formed characters are not cut in half
            cut = sum(len(p) for p in parts[: int(r.integers(0, len(parts) + 1))])
            s = s[:cut] + bad + s[cut:]
            s = s[:160]
        out.append(s)
        total += len(s)
        offsets[i + 1] = total
    return np.frombuffer(b"".join(out), dtype=np.uint8).copy(), offsets


def gen_c4_block(nbytes, seed_stream=0, with_match=False):
    """a block of log lines (80..200 bytes, 90 % LF / 10 % CRLF; 5 % start with ERROR) that holds NO line matching
    `^ERROR.*timeout=\\d+$` unless with_match; decoys: ERROR lines ending `timeout=12x`, `timeout=` mid-line."""
    r = rng_for("c4", seed_stream)
    lines = []
    size = 0
    while size < nbytes:
        ln = int(r.integers(80, 201))
        u = r.random()
        level = b"ERROR" if u < 0.05 else (b"WARN " if u < 0.35 else b"INFO ")
        body = bytes(r.integers(0x20, 0x7F, size=ln, dtype=np.uint8))
        body = body.replace(b"timeout=", b"timeout:")
        v = r.random()
        if level == b"ERROR" and v < 0.3:
            tail = b" timeout=12x"                      # near miss at the end
            line = level + body[: ln - 5 - len(tail)] + tail
        elif v < 0.4:
            mid = ln // 2
            line = level + body[:mid] + b"timeout=" + body[mid + 8: ln - 5] + b"!"   # `timeout=` mid-line, never at the end
        else:
            line = level + body[: ln - 6] + b"."
        term = b"\r\n" if r.random() < 0.1 else b"\n"
        lines.append(line + term)
        size += len(line) + len(term)
    blob = b"".join(lines)[:nbytes]
    return np.frombuffer(blob, dtype=np.uint8).copy()


C4_MATCH_LINE = b"ERROR worker 17 gave up waiting for the upstream after retries timeout=30000"


def gen_c4(nbytes, match_at=0.999, seed_stream=0, crlf=False):
    """buffer of nbytes with exactly one fully matching line placed near byte fraction match_at"""
    buf = gen_c4_block(nbytes, seed_stream)
    if match_at is None:
        return buf
    pos = int(nbytes * match_at)
    # align the planted line to a line start: find the previous LF
    view = buf[:pos]
    nl = np.nonzero(view == 10)[0]
    start = int(nl[-1]) + 1 if len(nl) else 0
    line = C4_MATCH_LINE + (b"\r\n" if crlf else b"\n")
    end = start + len(line)
    if end >= nbytes:
        raise ValueError("buffer too small for the planted line")
    buf[start:end] = np.frombuffer(line, dtype=np.uint8)
    # the bytes after the planted line belong to a cut line: make it a harmless INFO line start
    nxt = np.nonzero(buf[end:] == 10)[0]
    if len(nxt):
        buf[end: end + min(5, int(nxt[0]))] = np.frombuffer(b"INFO ", dtype=np.uint8)[: min(5, int(nxt[0]))]
    return buf


def gen_c5(n, seed_stream=0):
    """n x 64 bytes over {a,b} with `c` at each position with p = 1/8"""
    r = rng_for("c5", seed_stream)
    buf = r.integers(0, 2, size=(n, 64), dtype=np.uint8) + ord("a")
    buf[r.random((n, 64)) < 0.125] = ord("c")
    return buf.reshape(-1), n, 64
