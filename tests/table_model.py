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

    def __init__(self, pattern_obj, use_direct=True):
        self.p = pattern_obj
        self.op = pattern_obj.op
        self.use_direct = use_direct
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
        if self.op == 1:
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


class SpanLinear:
    """span_linear_smem (K3f): forward ordered-groups automaton, then the reverse automaton"""

    def __init__(self, pattern_obj):
        self.t = pattern_obj.span_tables()
        assert self.t is not None
        self.cuts = [int(x) for x in self.t["cuts"]]

    def cls(self, cp):
        import bisect
        return bisect.bisect_right(self.cuts, cp) - 1

    def regex(self, s: bytes):
        t = self.t
        n = len(s)
        if n == 0 or s == b" ":
            return (0, 0)
        direct, flags = t["direct"], t["flags"]
        w = t["start"]
        last = 0 if (flags[t["start"]] & SF_ACC) else -1
        seq, inter = 0, False
        j = 0
        dead = False
        while j < n:
            b = s[j]
            if inter and (b & 0xC0) != 0x80:
                f = int(flags[w & W_STATE])
                for k in range(1, j - seq + 1):
                    if f & (SF_FAILACC1 << (k - 1)):
                        last = seq + k
                inter = False
            nw = int(direct[w & W_STATE, b])
            if (nw & W_INTER) and not inter:
                seq = j
            inter = bool(nw & W_INTER)
            w = nw
            if w & W_ACC:
                last = j + 1
            if (w & W_STATE) == 0:
                dead = True
                break
            j += 1
        if not dead:
            f = int(flags[w & W_STATE])
            if inter:
                for k in range(1, n - seq + 1):
                    if f & (SF_FAILACC1 << (k - 1)):
                        last = seq + k
            if f & SF_END:
                last = n + 1
        if last <= 0:
            return (0, 0)
        rd, ok = t["rdelta"], t["rstartok"]
        nul = self.cls(0)
        r = t["rstart"]
        pos = last
        if last > n:
            r = int(rd[r, nul])
            pos = n
        best = -2
        while r != 0 and pos > 0:
            c = s[pos - 1]
            q = pos - 1
            cp = c
            if c >= 0x80:
                cp = 0xFFFF
                if (c & 0xC0) == 0x80:
                    acc, shift = c & 0x3F, 6
                    for back in range(2, 5):
                        if pos - back < 0:
                            break
                        d = s[pos - back]
                        if (d & 0xC0) == 0x80:
                            acc |= (d & 0x3F) << shift
                            shift += 6
                            continue
                        nn = 2 if (d >> 5) == 6 else 3 if (d >> 4) == 14 else 4 if (d >> 3) == 30 else 1
                        if nn == back:
                            lead = (d & 0x1F) if nn == 2 else (d & 0x0F) if nn == 3 else (d & 0x07)
                            cp = acc | (lead << shift)
                            q = pos - back
                        break
            r = int(rd[r, self.cls(cp)])
            if r == 0:
                break
            pos = q
            if ok[r]:
                best = pos
        if r != 0 and pos == 0:
            r = int(rd[r, nul])
            if r != 0 and ok[r]:
                best = -1
        assert best != -2, "forward and reverse automata disagree"
        return (1 if best < 0 else best + 1, min(last, n))
