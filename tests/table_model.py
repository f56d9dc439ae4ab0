"""Python model of what the kernels do with the device tables (forgex_b200/csrc/fx_kernels.cuh).

Used by the CPU-only tests to check the HOST side of the product (pattern compiler + table builder)
against the oracle without a GPU: same tables, same walk, written independently in Python.
It lives under tests/ on purpose -- it is a checker, not a fallback: the product never imports it.
"""
import forgex_b200 as fx

W_ACC, W_INTER, W_STATE = 0x8000, 0x4000, 0x3FFF
SF_ACC, SF_END, SF_INTER, SF_MATCHED, SF_FAILACC1 = 1, 2, 4, 8, 16
NONE = -9999


def blank(b: bytes):
    return b.strip(b" ") == b""


class Anchored:
    """run_attempt / attempt_at / including_exact on a REGEX-mode (flag-bit) table"""

    def __init__(self, pattern_obj, use_direct):
        self.t = pattern_obj.tables()
        self.lit = pattern_obj.literals()
        self.use_direct = use_direct

    def nxt(self, state, byte):
        if self.use_direct:
            return int(self.t["direct"][state, byte])
        return int(self.t["table"][state, self.t["classmap"][byte]])

    def attempt(self, s, st, pos, last):
        flags = self.t["flags"]
        w = st
        seq = 0
        inter = False
        n = len(s)
        if (w & W_STATE) == 0:
            return last
        j = pos
        while j <= n:
            b = s[j] if j < n else 0
            if inter and (b & 0xC0) != 0x80:
                f = int(flags[w & W_STATE])
                for k in range(1, j - seq + 1):
                    if f & (SF_FAILACC1 << (k - 1)):
                        last = seq + k
                inter = False
            nw = self.nxt(w & W_STATE, b)
            if (nw & W_INTER) and not inter:
                seq = j
            inter = bool(nw & W_INTER)
            w = nw
            if w & W_ACC:
                last = j + 1
            if (w & W_STATE) == 0:
                break
            j += 1
        return last

    def attempt_at(self, s, start):
        t = self.t
        if start == 1:
            if t["start_nul"] == 0:
                return -1
            last = 0 if (t["flags"][t["start_nul"]] & SF_ACC) else -1
            return self.attempt(s, t["start_nul"], 0, last)
        return self.attempt(s, t["q0"], start - 2, -1)

    @staticmethod
    def char_len(s, pos):
        b = s[pos]
        if b < 0x80:
            return 1
        if (b >> 5) == 6:
            n = 2
        elif (b >> 4) == 14:
            n = 3
        elif (b >> 3) == 30:
            n = 4
        else:
            return 1
        if pos + n > len(s):
            return 1
        for k in range(1, n):
            if (s[pos + k] & 0xC0) != 0x80:
                return 1
        return n

    def including_exact(self, s):
        _, pre, suf = self.lit
        n = len(s)
        m = n + 2
        S = b"\0" + s + b"\0"
        pre_active, suf_active = not blank(pre), not blank(suf)

        def findex(lit, frm):     # framed_index: 1-based position in S at or after frm, 0 if none
            i = S.find(lit, frm - 1)
            return i + 1 if i >= 0 else 0

        def rindex(hay, lit):     # index(hay, lit, back=.true.)
            if len(lit) > len(hay):
                return 0
            i = hay.rfind(lit)
            return i + 1 if i >= 0 else 0
        brute = not pre_active
        first, frame_suf, offset, more = NONE, NONE, 0, False
        if not brute:
            idx = findex(pre, 1)
            frame_suf = rindex(S, suf) or NONE
            if idx > 0:
                if frame_suf != NONE:
                    if idx <= frame_suf:
                        first = idx
                else:
                    first = idx
                offset = idx + len(pre) - 1
                more = True
            if first == NONE:
                brute = True
        if brute:
            last = self.attempt_at(s, 1)
            if last >= 0:
                return 1, min(last, n)
            pos = 0
            while pos < n:
                last = self.attempt(s, self.t["q0"], pos, -1)
                if last >= 0:
                    return pos + 1, min(last, n)
                pos += self.char_len(s, pos)
            return 0, 0
        at_zero = first == 2
        start = 1 if at_zero else first
        text_suf = NONE
        if suf_active:
            text_suf = rindex(s, suf)
            if text_suf == 0:
                return 0, 0
        while start < m:
            if text_suf != NONE and text_suf < start:
                return 0, 0
            last = self.attempt_at(s, start)
            if last >= 0:
                return max(start - 1, 1), min(last, n)
            if at_zero:
                at_zero = False
                start = first
                continue
            if not more or not (offset < m):
                return 0, 0
            hit = findex(pre, offset + 1)
            if hit <= 0:
                return 0, 0
            start = hit
            offset = hit + len(pre) - 1
            if frame_suf != NONE and offset > frame_suf:
                more = False
        return 0, 0

    def regex(self, s: bytes):   # eval_regex
        all_ = self.lit[0]
        if not blank(all_):
            i = s.find(all_)
            return (i + 1, i + len(all_)) if i >= 0 else (0, 0)
        if len(s) == 0 or s == b" ":
            return (0, 0)
        f, t = self.including_exact(s)
        return (f, t) if f > 0 and t > 0 else (0, 0)


class Model:
    """one compiled pattern, evaluated the way the kernels evaluate it"""

    def __init__(self, pattern_obj, use_direct=True, gates=True):
        self.p = pattern_obj
        self.op = pattern_obj.op
        self.use_direct = use_direct
        self.gates = gates          # False: what the batch kernels compute for a pattern the host did NOT mark `gated`
        self.t = pattern_obj.tables()
        self.lit = pattern_obj.literals()
        self.prefix_mode = pattern_obj.info()["prefix_mode"]
        self.anch = None
        if self.op == 2:
            self.anch = Anchored(pattern_obj, use_direct)
        elif self.prefix_mode:
            self._anch_pat = fx.Pattern(pattern_obj.pattern, "regex")
            self.anch = Anchored(self._anch_pat, False)

    def nxt(self, state, byte):
        if self.use_direct:
            return int(self.t["direct"][state, byte])
        return int(self.t["table"][state, self.t["classmap"][byte]])

    def in_with_prefix(self, s):
        if len(s) == 0 or s == b" ":
            return self.t["q0_accepting"]
        f, t = self.anch.including_exact(s)
        return f > 0 and t > 0

    def boolean(self, s: bytes):   # k_bool_* incl. eval_bool_generic
        all_, pre, suf = self.lit
        t = self.t
        if not self.gates:
            pass                    # un-gated fast path: degenerate texts and the automaton walk only
        elif self.op == 1:
            if not blank(all_):
                return all_ in s
            if self.prefix_mode == 2:
                return self.in_with_prefix(s)
        else:
            if not blank(all_) and len(s) == len(all_):
                return s == all_
            lp, ls = len(pre), len(suf)
            if len(s) > 0 and lp > 0 and lp == len(s) and s == pre:
                return True
            if lp > len(s) or ls > len(s):
                return False
            if len(s) > 0:
                if not blank(pre) and s[:lp] != pre:
                    return False
                if not blank(suf) and s[len(s) - ls:] != suf:
                    return False
            else:
                if not blank(pre) and lp != 0:
                    return False
                if not blank(suf) and ls != 0:
                    return False
        if len(s) == 0 or (self.op == 1 and s == b" "):
            return t["q0_accepting"]
        st = t["start"]
        high = 0
        for b in s:
            high |= b
            st = self.nxt(st, b)
        r = bool(t["flags"][st] & (SF_END | SF_MATCHED))
        assert r == (st >= t["result_threshold"])   # the kernels decide with this compare
        if self.op == 1 and r and self.prefix_mode == 1 and (high & 0x80):
            r = self.in_with_prefix(s)
        return r

    def regex(self, s: bytes):
        return self.anch.regex(s)


class SparseIn:
    """k_in_sparse: `.in.` as "some candidate start wins" -- candidates are the leading NUL and the positions whose
    byte passes the sweep filter (ASCII ranges, optionally every byte >= 0xC0) and survives the first table step"""

    def __init__(self, pattern_obj):
        inf = pattern_obj.info()
        assert inf["sparse"]
        self.ranges = list(zip(inf["sparse_lo"], inf["sparse_hi"]))
        self.high = bool(inf["sparse_high"])
        self.prefix_mode = inf["prefix_mode"]
        self.q0_accepting = pattern_obj.tables()["q0_accepting"]
        self._anch_pat = fx.Pattern(pattern_obj.pattern, "regex")
        self.anch = Anchored(self._anch_pat, False)

    def sweep(self, b):
        return any(lo <= b <= hi for lo, hi in self.ranges) or (self.high and b >= 0xC0)

    def boolean(self, s: bytes):
        if len(s) == 0 or s == b" ":
            return self.q0_accepting
        a, t = self.anch, self.anch.t
        r = False
        if t["start_nul"] != 0:
            assert not (t["flags"][t["start_nul"]] & SF_ACC)
            r = a.attempt(s, t["start_nul"], 0, -1) >= 0
        for pos, b in enumerate(s):
            if r:
                break
            if not self.sweep(b):
                assert (a.nxt(t["q0"], b) & (W_STATE | W_ACC)) == 0   # the filter never drops a byte of F
                continue
            if (a.nxt(t["q0"], b) & (W_STATE | W_ACC)) == 0:
                continue
            r = a.attempt(s, t["q0"], pos, -1) >= 0
        if r and self.prefix_mode == 1 and any(b >= 0x80 for b in s):
            f, to = a.including_exact(s)
            r = f > 0 and to > 0
        return r


class BufferPrefix:
    """the long-buffer path for a pattern with a prefix literal (k_buffer_scan_sparse<..., PREFIX> + gated plain scan
    + k_buffer_finish): winner = the smallest candidate start whose attempt wins, candidates = the literal's
    occurrences (plus the leading NUL if the literal sits at the front), or every boundary if it occurs nowhere"""

    def __init__(self, pattern_obj):
        assert pattern_obj.info()["prefix_scan"]
        self.anch = Anchored(pattern_obj, False)
        self.pre = pattern_obj.literals()[1]

    def regex(self, s: bytes):
        a = self.anch
        if len(s) == 0 or s == b" ":
            return (0, 0)
        pre, n = self.pre, len(s)
        occ = [i for i in range(0, n - len(pre) + 1) if s[i:i + len(pre)] == pre]
        key = None
        if occ:
            starts = ([1] if occ[0] == 0 else []) + [i + 2 for i in occ]
        else:
            starts = [1]
            pos = 0
            while pos < n:
                starts.append(pos + 2)
                pos += a.char_len(s, pos)
        for st in starts:                     # ascending, so the first winner is the minimum
            if a.attempt_at(s, st) >= 0:
                key = st
                break
        if key is None:
            return (0, 0)
        last = a.attempt_at(s, key)
        f = max(key - 1, 1)
        t = min(last, n)
        return (f, t) if f > 0 and t > 0 else (0, 0)


W_SSTATE, W_RA = 0x0FFF, 0x3000


class SpanLinear:
    """span_linear / k_span_ragged (K3f): forward ordered-groups automaton in its span word format (every event of a
    step is in the word: ACC, and RA = a broken sequence's replay passed an accept RA-1 bytes back), then the reverse
    automaton over code-point classes with its two-level class map"""

    def __init__(self, pattern_obj):
        self.t = pattern_obj.span_tables()
        assert self.t is not None
        self.cuts = [int(x) for x in self.t["cuts"]]

    def cls(self, cp):
        t = self.t
        if cp < 0x10000 and t["rpage"] is not None:
            pg = int(t["rpage"][cp >> 6])
            c = pg if pg < 0x80 else int(t["rmixed"][((pg & 0x7F) << 6) | (cp & 63)])
            import bisect
            assert c == bisect.bisect_right(self.cuts, cp) - 1      # the two-level map equals the binary search
            return c
        import bisect
        return bisect.bisect_right(self.cuts, cp) - 1

    def forward(self, s: bytes, st=None, last=None, base=0):
        """walk s from state st; returns (state, last) with positions offset by base (used by the chunked model too)"""
        direct = self.t["direct"]
        if st is None:
            st = self.t["start"]
            last = 0 if self.t["start_acc"] else -1
        for j, b in enumerate(s):
            if st == 0:
                break
            nw = int(direct[st, b])
            if nw & W_RA:
                last = base + j + 1 - ((nw >> 12) & 3)
            if nw & W_ACC:
                last = base + j + 1
            st = nw & W_SSTATE
        return st, last

    def end_of_text(self, st, n, last):
        if st != 0:
            e = int(self.t["endinfo"][st])
            if e & 3:
                last = n + 1 - (e & 3)
            if e & 4:
                last = n + 1
        return last

    def backward(self, s: bytes, last):
        t = self.t
        n = len(s)
        rd = t["rdelta"]
        nul = self.cls(0)
        r = t["rstart"]
        pos = last
        if last > n:
            r = int(rd[r, nul]) & 0x7FFF
            pos = n
        best = -2
        while r != 0 and pos > 0:
            c = s[pos - 1]
            q = pos - 1
            cp = c
            if c >= 0x80:
                cp = 0xFFFF
                if (c & 0xC0) == 0x80 and pos >= 2:
                    d1 = s[pos - 2]
                    if (d1 & 0xE0) == 0xC0:
                        cp, q = ((d1 & 0x1F) << 6) | (c & 0x3F), pos - 2
                    elif (d1 & 0xC0) == 0x80 and pos >= 3:
                        d2 = s[pos - 3]
                        if (d2 & 0xF0) == 0xE0:
                            cp, q = ((d2 & 0x0F) << 12) | ((d1 & 0x3F) << 6) | (c & 0x3F), pos - 3
                        elif (d2 & 0xC0) == 0x80 and pos >= 4:
                            d3 = s[pos - 4]
                            if (d3 & 0xF8) == 0xF0:
                                cp, q = ((d3 & 0x07) << 18) | ((d2 & 0x3F) << 12) | ((d1 & 0x3F) << 6) | (c & 0x3F), pos - 4
            w = int(rd[r, self.cls(cp)])
            r = w & 0x7FFF
            if r == 0:
                break
            pos = q
            if w & 0x8000:
                best = pos
        if r != 0 and pos == 0:
            w = int(rd[r, nul])
            if (w & 0x7FFF) != 0 and (w & 0x8000):
                best = -1
        assert best != -2, "forward and reverse automata disagree"
        return 1 if best < 0 else best + 1

    def regex(self, s: bytes):
        n = len(s)
        if n == 0 or s == b" ":
            return (0, 0)
        st, last = self.forward(s)
        last = self.end_of_text(st, n, last)
        if last <= 0:
            return (0, 0)
        return (self.backward(s, last), min(last, n))


class StateMapScan:
    """the long-buffer state-map scan (k_statemap_regions + k_statemap_compose): the text is cut into regions, a region
    into sub-chunks; every sub-chunk is walked from a small CANDIDATE set of states that is guaranteed to hold the true
    incoming state (the image of ALL reachable states under the byte(s) in front of it), candidates that reach the same
    state merge, and the per-sub-chunk maps candidate -> (end state, last accept) are chained.  Model of the algorithm,
    not of the lane layout: sub-chunk and region sizes are parameters so that tiny texts exercise every boundary."""

    M = 4

    def __init__(self, pattern_obj, sub=8, region=40, lookback=6):
        self.sl = SpanLinear(pattern_obj)
        t = self.sl.t
        self.direct = t["direct"]
        self.sub, self.region, self.lookback = sub, region, lookback
        st = self.direct & W_SSTATE
        reach, work = {int(t["start"])}, [int(t["start"])]
        while work:
            s = work.pop()
            for x in set(int(v) for v in st[s]):
                if x and x not in reach:
                    reach.add(x)
                    work.append(x)
        self.reach = sorted(reach)
        self.img = []
        for b in range(256):
            im = sorted(set(int(v) for v in st[self.reach, b]) - {0})
            self.img.append(im if len(im) <= self.M else None)
        # synchronising bytes: after c and any one more byte at most one live state is left (sub-chunks start behind them)
        self.sync = []
        for b in range(256):
            im = self.img[b]
            ok = im is not None and len(im) >= 1
            if ok:
                for d in range(256):
                    if len(set(int(st[s, d]) for s in im) - {0}) > 1:
                        ok = False
                        break
            self.sync.append(ok)
        self.scan = max(1, sub // 2)
        self.stats = {"unknown": 0, "complex": 0, "rewalk": 0, "wide_regions": 0}

    def bnd(self, text, k):
        """where sub-chunk k starts: right behind the first synchronising byte near its nominal start k * sub"""
        if k <= 0:
            return 0
        nom = k * self.sub
        if nom >= len(text):
            return len(text)
        for q in range(nom - 1, min(len(text), nom - 1 + self.scan)):
            if self.sync[text[q]]:
                return q + 1
        return nom

    def step(self, s, b):
        return int(self.direct[s, b]) & W_SSTATE

    def candidates(self, text, b):
        if b == 0:
            return [int(self.sl.t["start"])]
        im = self.img[text[b - 1]]
        if im is not None:
            return list(im)
        if b >= 2 and self.img[text[b - 2]] is not None:
            out = []
            for s in self.img[text[b - 2]]:
                n = self.step(s, text[b - 1])
                if n and n not in out:
                    out.append(n)
            return out if len(out) <= self.M else None
        return None

    def sub_map(self, text, b, e):
        """candidate -> (end, last) for text[b:e], or None (unknown candidates / an accept while candidates still differ)"""
        cands = self.candidates(text, b)
        if cands is None:
            self.stats["unknown"] += 1
            return None
        st = list(cands)
        root = list(range(len(st)))          # which trajectory a candidate follows
        active = [s != 0 for s in st]
        last, owner = None, -1               # an accept is only recorded while ONE trajectory is active: `owner`
        for j in range(b, e):
            for k in range(len(st)):
                if not active[k]:
                    continue
                nw = int(self.direct[st[k], text[j]])
                if nw & (W_RA | W_ACC):
                    if sum(active) > 1:
                        self.stats["complex"] += 1
                        return None
                    owner = k
                    if nw & W_RA:
                        last = j + 1 - ((nw >> 12) & 3)
                    if nw & W_ACC:
                        last = j + 1
                st[k] = nw & W_SSTATE
                if st[k] == 0:
                    active[k] = False
            if (j - b) % 4 == 3:             # merge check at block ends
                for k in range(len(st)):
                    for m in range(k):
                        if active[k] and active[m] and st[k] == st[m]:
                            active[k] = False
                            for q in range(len(st)):
                                if root[q] == k:
                                    root[q] = m
        out = {}
        for k, c in enumerate(cands):
            r = root[k]
            out[c] = (st[r], last if r == owner else None)   # the accepts belong to the candidates that follow the owner
        return out

    def walk(self, text, b, e, s, last):
        for j in range(b, e):
            if s == 0:
                break
            nw = int(self.direct[s, text[j]])
            if nw & W_RA:
                last = j + 1 - ((nw >> 12) & 3)
            if nw & W_ACC:
                last = j + 1
            s = nw & W_SSTATE
        return s, last

    def region_map(self, text, r0, r1):
        """region candidates (full enumeration over a look-back window) -> (end, last)"""
        if r0 == 0:
            rc = [int(self.sl.t["start"])]
        else:
            lb = min(self.lookback, r0)
            rc = []
            for s in self.reach:
                for j in range(r0 - lb, r0):
                    s = self.step(s, text[j])
                    if s == 0:
                        break
                if s and s not in rc:
                    rc.append(s)
            if len(rc) > 32:
                self.stats["wide_regions"] += 1
                return None
        chain = [(c, None) for c in rc]
        for k in range(self._k0, self._k1):
            b, e = min(self.bnd(text, k), r1), min(self.bnd(text, k + 1), r1)
            if b >= e:
                continue
            m = self.sub_map(text, b, e)
            for i, (c, last) in enumerate(chain):
                if c == 0:
                    continue
                if m is not None and c in m:
                    ne, nl = m[c]
                    chain[i] = (ne, nl if nl is not None else last)
                else:
                    assert m is None, "the candidate set must hold every state that can arrive here"
                    self.stats["rewalk"] += 1
                    chain[i] = self.walk(text, b, e, c, last)
        return {c: chain[i] for i, c in enumerate(rc)}

    def last(self, text: bytes):
        """end of the leftmost-longest match as the forward walk defines it (None: the scan declined)"""
        n = len(text)
        s = int(self.sl.t["start"])
        last = 0 if self.sl.t["start_acc"] else -1
        spr = max(1, self.region // self.sub)          # sub-chunks per region
        reg = 0
        while s != 0 and reg * spr * self.sub < max(n, 1):
            self._k0, self._k1 = reg * spr, (reg + 1) * spr
            r0, r1 = self.bnd(text, self._k0), min(n, self.bnd(text, self._k1))
            m = self.region_map(text, r0, r1)
            if m is None or s not in m:
                return None
            s, l = m[s]
            if l is not None:
                last = l
            reg += 1
        return self.sl.end_of_text(s, n, last)


class NfaModel:
    """the NFA engine (k_nfa_bool / k_nfa_regex): the reference's per-character subset step on bit sets -- here Python
    integers -- with the same drivers as the table engine (attempt_at, brute force, prefix candidates, .match. walk)"""

    def __init__(self, pattern_obj):
        t = pattern_obj.nfa_tables()
        assert t is not None
        self.op = pattern_obj.op
        self.lit = pattern_obj.literals()
        nst, ncls, w = t["trans"].shape
        self.trans = [[int.from_bytes(t["trans"][s, c].tobytes(), "little") for c in range(ncls)] for s in range(nst)]
        self.q0 = int.from_bytes(t["q0"].tobytes(), "little")
        self.cuts = [int(x) for x in t["cuts"]]
        self.exit = t["exit"]
        self.q0_accepting = t["q0_accepting"]

    def cls(self, cp):
        import bisect
        return bisect.bisect_right(self.cuts, cp) - 1

    def step(self, cur, c):
        nxt, s = 0, 0
        while cur:
            if cur & 1:
                nxt |= self.trans[s][c]
            cur >>= 1
            s += 1
        return nxt

    def acc(self, cur):
        return (cur >> self.exit) & 1

    def symbol(self, s, j):
        n = Anchored.char_len(s, j)
        b = s[j]
        if b < 0x80:
            return self.cls(b), 1
        if n == 1:
            return self.cls(0xFFFF), 1
        cp = b & (0x1F if n == 2 else 0x0F if n == 3 else 0x07)
        for k in range(1, n):
            cp = (cp << 6) | (s[j + k] & 0x3F)
        return self.cls(cp), n

    def attempt_at(self, s, start):
        n, cur, last, j = len(s), self.q0, -1, start - 2
        if start == 1:
            cur = self.step(cur, self.cls(0))
            if not cur:
                return -1
            if self.acc(cur):
                last = 0
            j = 0
        while j <= n:
            c, nb = self.symbol(s, j) if j < n else (self.cls(0), 1)
            cur = self.step(cur, c)
            if not cur:
                break
            j += nb
            if self.acc(cur):
                last = j
        return last

    def match(self, s):
        if len(s) == 0:
            return self.q0_accepting
        cur = self.step(self.q0, self.cls(0)) or self.q0
        j = 0
        while j < len(s):
            c, nb = self.symbol(s, j)
            cur = self.step(cur, c)
            if not cur:
                return False
            j += nb
        return bool(self.acc(cur)) or bool(self.acc(self.step(cur, self.cls(0))))

    def including(self, s):
        """including_exact with this engine's attempts (the Anchored class holds the driver; only attempts differ)"""
        drv = Anchored.__new__(Anchored)
        drv.lit = self.lit
        drv.t = {"q0": None}
        eng = self
        drv.attempt_at = lambda text, start: eng.attempt_at(text, start)
        drv.attempt = lambda text, st, pos, last: eng.attempt_at(text, pos + 2)
        return drv.including_exact(s)

    def boolean(self, s: bytes):
        all_, pre, suf = self.lit
        if self.op == 1:
            if not blank(all_):
                return all_ in s
            if len(s) == 0 or s == b" ":
                return self.q0_accepting
            f, t = self.including(s)
            return f > 0 and t > 0
        if not blank(all_) and len(s) == len(all_):
            return s == all_
        lp, ls = len(pre), len(suf)
        if len(s) > 0 and lp > 0 and lp == len(s) and s == pre:
            return True
        if lp > len(s) or ls > len(s):
            return False
        if len(s) > 0:
            if not blank(pre) and s[:lp] != pre:
                return False
            if not blank(suf) and s[len(s) - ls:] != suf:
                return False
        else:
            if (not blank(pre) and lp != 0) or (not blank(suf) and ls != 0):
                return False
        return self.match(s)

    def regex(self, s: bytes):
        all_ = self.lit[0]
        if not blank(all_):
            i = s.find(all_)
            return (i + 1, i + len(all_)) if i >= 0 else (0, 0)
        if len(s) == 0 or s == b" ":
            return (0, 0)
        f, t = self.including(s)
        return (f, t) if f > 0 and t > 0 else (0, 0)
