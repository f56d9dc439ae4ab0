"""One-off extended GPU fuzz of the fixed-stride kernels (not part of the -m gpu suite): random patterns, `.in.` and
`.match.`, over fixed-stride batches of many strides and buffer alignments (K1's 8/16-byte vector forms, the byte form,
K2c with a stride, the gated path), against the oracle."""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import forgex_b200 as fx  # noqa: E402
from tests import oracle_lib as O  # noqa: E402
from tests.test_host_tables import gen_pattern, gen_text  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    tried = 0
    while time.time() - t0 < budget:
        rng = random.Random(88000 + seed)
        blob = b"".join(gen_text(rng) + rng.choice((b"", b" ", b"\n", b"a")) for _ in range(6000))
        os.environ["FX_SPARSE_MAX_FIRST"] = "128" if seed % 2 else "6"
        for _ in range(10):
            pat = gen_pattern(rng).encode()
            for op in ("in", "match"):
                p = fx.Pattern(pat, op)
                if p.status != 0:
                    continue
                o = 1 if op == "match" else 0
                c = O.Compiled(pat, o)
                for stride in (1, 2, 3, 4, 7, 8, 12, 16, 24, 32, 40, 64):
                    shift = rng.randrange(0, 16)
                    view = np.frombuffer(blob, dtype=np.uint8)[shift:]
                    n = min(len(view) // stride, 3000)
                    fb = np.ascontiguousarray(view[: n * stride])
                    # keep the chosen misalignment: a copy into a buffer with the same offset from a 16-byte boundary
                    raw = np.zeros(n * stride + 32, dtype=np.uint8)
                    base = (-raw.ctypes.data) % 16
                    dst = raw[base + shift: base + shift + n * stride]
                    dst[:] = fb
                    got = p.in_fixed(dst, n, stride) if op == "in" else p.match_fixed(dst, n, stride)
                    exp = c.bool_fixed(o, fb, n, stride)
                    assert np.array_equal(got, exp), (seed, pat, op, stride, shift, np.nonzero(got != exp)[0][:5])
                tried += 1
        print("seed", seed, "pattern/op pairs", tried, flush=True)
        seed += 1
    print("extended fixed-stride fuzz ok: %d pattern/op pairs x 12 strides in %.0f s" % (tried, time.time() - t0))


if __name__ == "__main__":
    main()
