"""One-off extended GPU fuzz of the ragged batch kernels with LONG strings mixed in (not part of the -m gpu suite):
strings of 1-40 KB between short ones -- longer than a warp's / a CTA's staged tile, so the not-staged paths, the tile
cut-offs and the degenerate strings behind them are exercised for random patterns -- .in. / .match. / regex / counts
against the oracle."""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import forgex_b200 as fx  # noqa: E402
from tests import oracle_lib as O  # noqa: E402
from tests.test_host_tables import gen_pattern, gen_text  # noqa: E402


def pack(strings):
    off = np.zeros(len(strings) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in strings], out=off[1:])
    return np.frombuffer(b"".join(strings), dtype=np.uint8).copy(), off


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    tried = 0
    while time.time() - t0 < budget:
        rng = random.Random(66000 + seed)
        texts = []
        for _ in range(260):
            r = rng.random()
            if r < 0.08:
                texts.append(b"".join(gen_text(rng) for _ in range(rng.choice((200, 900, 2500, 6000)))))
            elif r < 0.14:
                texts.append(rng.choice((b"", b" ", b"  ")))
            else:
                texts.append(gen_text(rng))
        buf, off = pack(texts)
        os.environ["FX_SPARSE_MAX_FIRST"] = "128" if seed % 2 else "6"
        for _ in range(8):
            pat = gen_pattern(rng).encode()
            for op in ("in", "match", "regex"):
                p = fx.Pattern(pat, op)
                if p.status != 0:
                    continue
                c = O.Compiled(pat, 1 if op == "match" else 0)
                if op == "regex":
                    f, t = p.regex_batch(buf, off)
                    ef, et = c.regex_batch(buf, off)
                    assert np.array_equal(f, ef) and np.array_equal(t, et), (seed, pat, np.nonzero((f != ef) | (t != et))[0][:5])
                    if not p.info()["nfa_engine"]:
                        cnt = p.regex_count_batch(buf, off)
                        # the oracle's loop on the short strings only (the long ones are checked through regex_batch)
                        for i in rng.sample(range(len(texts)), 25):
                            if len(texts[i]) > 300:
                                continue
                            k, pos = 0, 0
                            arr = np.frombuffer(texts[i], dtype=np.uint8)
                            while True:
                                a, b = c.regex_buffer(np.ascontiguousarray(arr[pos:]))
                                if a <= 0 or b <= 0:
                                    break
                                k += 1
                                pos += b
                            assert int(cnt[i]) == k, (seed, pat, i, int(cnt[i]), k)
                else:
                    o = 1 if op == "match" else 0
                    got = p.in_batch(buf, off) if op == "in" else p.match_batch(buf, off)
                    exp = c.bool_batch(o, buf, off)
                    assert np.array_equal(got, exp), (seed, pat, op, np.nonzero(got != exp)[0][:5])
                tried += 1
        print("seed", seed, "bytes", len(buf), "pattern/op pairs", tried, flush=True)
        seed += 1
    print("extended long-string fuzz ok: %d pattern/op pairs in %.0f s" % (tried, time.time() - t0))


if __name__ == "__main__":
    main()
