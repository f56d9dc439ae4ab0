import numpy as np, sys
sys.path.insert(0, '.')
import forgex_b200 as fx
from tests import oracle_lib as O
from tests.test_gpu_parity import pack, oracle_bool
from tools import synth
rng = np.random.default_rng(5)
pieces = [b"a", b"z", b" ", "あ".encode(), "ん".encode(), "α".encode(), "　".encode(), b"\x80", b"\xbf", b"\xc3", b"\xe3\x81",
          b"\xf0\x9f\x98", b"\xff", b"\xc0\x80", b"\xc1\xa1", b"\xe0\x81\xa1", b"\xef\xbf\xbf", b"\xf4\x90\x80\x81", b"\n", b"\r\n", b"_", b"7"]
strings = [b"".join(pieces[i] for i in rng.integers(0, len(pieces), size=int(k))) for k in rng.integers(0, 24, size=4000)]
buf, off = pack(strings)
for pat in [synth.PATTERNS["c3"], b"[a-z]+", b".+", rb"\S+", "[ぁ-ん]+".encode(), b"a.", rb"[^a]{2,3}$", rb"^\w"]:
    q = fx.Pattern(pat, "in")
    got = q.in_batch(buf, off)
    exp = oracle_bool(pat, "in", buf, offsets=off)
    bad = np.nonzero(got != exp)[0]
    i = q.info()
    print(pat, 'sparse', i['sparse'], i['sparse_used'], 'nbad', len(bad))
    for k in bad[:5]:
        print('   ', k, strings[k], 'got', got[k], 'exp', exp[k])
