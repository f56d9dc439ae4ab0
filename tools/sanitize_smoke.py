"""small run of every kernel family, meant for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`
(`quick` as first argument: only the kernels that were new in round 1).

Round 1 results on a B200 (same checksum as the native run every time): memcheck 0 errors, racecheck 0 hazards,
synccheck 0 errors.  `round2` as first argument: the kernels that were new or rewritten in round 2 (results in
profiles/r02_sanitizer.txt).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import forgex_b200 as fx  # noqa: E402
from tools import synth  # noqa: E402


def pack(strings):
    off = np.zeros(len(strings) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in strings], out=off[1:])
    return np.frombuffer(b"".join(strings), dtype=np.uint8).copy(), off


def quick():
    """the round's new kernels only: K2c (ragged + fixed), K4 sparse sweep, prefix scan, K3f"""
    buf, off = synth.gen_c2(300)
    total = int(fx.Pattern(b"foo(bar|baz)", "in").in_batch(buf, off).sum())
    total += int(fx.Pattern(b"^foo", "in").in_batch(buf, off).sum())
    raw = np.random.default_rng(1).integers(0x20, 0x7F, size=50 * 160 + 7, dtype=np.uint8)
    total += int(fx.Pattern(b"foo(bar|baz)", "in").in_fixed(raw[7:], 50, 160).sum())
    text = synth.gen_c4(4099, 0.7)
    total += sum(fx.Pattern(synth.PATTERNS["c4"], "regex").regex_buffer(text))
    total += sum(fx.Pattern(b"foo(bar|baz)", "regex").regex_buffer(text[3:]))
    b3, o3 = synth.gen_c3(200)
    f, t = fx.Pattern(synth.PATTERNS["c3"], "regex").regex_batch(b3, o3)
    total += int(f.sum()) + int(t.sum())
    print("sanitize quick checksum", total)


def round2():
    """the kernels that were new or rewritten in round 2: K3f (span words, lane pool), K1c (compact table, both homes),
    K4 under its budget, K4L (literal), K5 (state-map scan, forced), K6 (NFA engine), the all-matches / count kernels,
    the sequential-candidate replay with its presence sweeps"""
    total = 0
    b3, o3 = synth.gen_c3(400)
    strings = [b"y" * 9000, b"", b" ", "あい ab".encode() * 3, b"\xe3\x81"] + [bytes(b3[o3[i]:o3[i + 1]]) for i in range(60)]
    bb, oo = pack(strings)
    for pat in [synth.PATTERNS["c3"], b"[a-z]+r", rb"\s\S+$"]:
        print("span batch", pat, flush=True)
        f, t = fx.Pattern(pat, "regex").regex_batch(bb, oo)
        total += int(f.sum()) + int(t.sum())
        total += int(fx.Pattern(pat, "regex").regex_count_batch(bb, oo).sum())
    print("compact table", flush=True)
    fb, n, stride = synth.gen_c5(300)
    fb = fb.copy()
    fb[70] = 0xC3
    for res in ("auto", "global"):
        total += int(fx.Pattern(synth.PATTERNS["c5"], "in", residency=res).in_fixed(fb, n, stride).sum())
    text = synth.gen_c4(60001, 0.7)
    for mode in ("1", "2"):
        print("buffer paths, FX_STATEMAP", mode, flush=True)
        os.environ["FX_STATEMAP"] = mode
        total += sum(fx.Pattern(synth.PATTERNS["c4"], "regex").regex_buffer(text[1:]))
        total += sum(fx.Pattern(rb"\w+@\w+", "regex").regex_buffer(text))
        total += sum(fx.Pattern(rb"ERROR.*timeout=\d+", "regex").regex_buffer(text[3:]))
        total += sum(fx.Pattern(b"[ab].*c", "regex").regex_buffer(np.frombuffer(b"a" * 40000 + b"c", dtype=np.uint8)))
    os.environ["FX_STATEMAP"] = "1"
    total += sum(fx.Pattern(b"ERROR", "regex").regex_buffer(text))
    total += sum(fx.Pattern(b"ab+c", "regex").regex_buffer(np.frombuffer(b"xx abbc " * 500, dtype=np.uint8)))
    total += sum(fx.Pattern(b"a.*b", "regex").regex_buffer(np.frombuffer(b"a" * 30000, dtype=np.uint8)))
    f, t, cnt = fx.Pattern(b"[A-Z]+", "regex").regex_buffer_all(text, capacity=64)
    total += cnt + int(f.sum())
    print("NFA engine", flush=True)
    os.environ["FX_STATE_CAP"] = "3"
    small, so = pack([b"foobar", b"xx foobaz", b"", b" ", "あい".encode(), b"\xff\xc3"])
    for op in ("in", "match", "regex"):
        p = fx.Pattern(b"foo(bar|baz)", op)
        assert p.info()["nfa_engine"] == 1
        if op == "regex":
            f, t = p.regex_batch(small, so)
            total += int(f.sum()) + sum(p.regex_buffer(small))
        else:
            total += int((p.in_batch(small, so) if op == "in" else p.match_batch(small, so)).sum())
    del os.environ["FX_STATE_CAP"]
    print("sanitize round-2 checksum", total)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        return quick()
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        return round2()
    rng = np.random.default_rng(3)
    buf, off = synth.gen_c2(300)
    strings = [b"", b" ", b"foobar", b"x" * 9000 + b"foobaz", b"\xc1\xa6oobar fooba!", b"f"] + \
              [bytes(rng.integers(0x20, 0x7F, size=int(k), dtype=np.uint8)) for k in rng.integers(0, 300, size=200)]
    b2, o2 = pack(strings)
    total = 0
    for pat in [b"foo(bar|baz)", b"^foo", b"[a-z]+r", rb"\d{3}-\d{4}", "[ぁ-ん]+a".encode()]:
        p = fx.Pattern(pat, "in")
        total += int(p.in_batch(buf, off).sum()) + int(p.in_batch(b2, o2).sum())
        total += int(fx.Pattern(pat, "match").match_batch(b2, o2).sum())
        f, t = fx.Pattern(pat, "regex").regex_batch(b2, o2)
        total += int(f.sum())
    fb, n, stride = synth.gen_c1(400)
    total += int(fx.Pattern(synth.PATTERNS["c1"], "match").match_fixed(fb, n, stride).sum())
    fb, n, stride = synth.gen_c5(50)
    total += int(fx.Pattern(synth.PATTERNS["c5"], "in").in_fixed(fb, n, stride).sum())
    raw = rng.integers(0x20, 0x7F, size=100 * 160 + 7, dtype=np.uint8)
    total += int(fx.Pattern(b"foo(bar|baz)", "in").in_fixed(raw[7:], 100, 160).sum())
    for nbytes in (4099, 30001):
        text = synth.gen_c4(nbytes, 0.7)
        total += sum(fx.Pattern(synth.PATTERNS["c4"], "regex").regex_buffer(text))
        total += sum(fx.Pattern(b"foo(bar|baz)", "regex").regex_buffer(text))
        total += sum(fx.Pattern(rb"ERROR.*timeout=\d+", "regex").regex_buffer(text[3:]))
        total += sum(fx.Pattern(rb"\w+@\w+", "regex").regex_buffer(text))
    print("sanitize smoke checksum", total)


if __name__ == "__main__":
    main()
