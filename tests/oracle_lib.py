"""ctypes loader for the CPU oracle (oracle/forgex_oracle.cpp).

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libforgex_oracle.so")

INVALID_CHAR_INDEX = -9999
ERRSTOP_TREE_LIMIT = -1001
ERRSTOP_DFA_LIMIT = -1002


def build(force=False):
    src = os.path.join(ORACLE_DIR, "forgex_oracle.cpp")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.fxo_error_message.restype = C.c_char_p
        L.fxo_error_message.argtypes = [C.c_int]
        L.fxo_is_valid.argtypes = [C.c_char_p, C.c_long, C.POINTER(C.c_int)]
        L.fxo_literals.argtypes = [C.c_char_p, C.c_long] + [C.c_char_p, C.POINTER(C.c_long)] * 3 + [C.c_long]
        L.fxo_in.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_long]
        L.fxo_match.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_long]
        L.fxo_regex.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_long] + [C.POINTER(C.c_long)] * 3 + \
            [C.POINTER(C.c_int)]
        L.fxo_compile.restype = C.c_void_p
        L.fxo_compile.argtypes = [C.c_char_p, C.c_long, C.c_int, C.POINTER(C.c_int)]
        L.fxo_free.argtypes = [C.c_void_p]
        L.fxo_steps.restype = C.c_long
        L.fxo_steps.argtypes = [C.c_void_p]
        L.fxo_dfa_states.restype = C.c_long
        L.fxo_dfa_states.argtypes = [C.c_void_p]
        L.fxo_bool_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.fxo_bool_fixed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_long, C.c_void_p]
        L.fxo_regex_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        L.fxo_regex_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong),
                                       C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def error_message(code):
    return lib().fxo_error_message(code).decode()


def is_valid(pattern: bytes):
    st = C.c_int(0)
    r = lib().fxo_is_valid(pattern, len(pattern), C.byref(st))
    return r, st.value


def literals(pattern: bytes):
    cap = 1 << 16
    bufs = [C.create_string_buffer(cap) for _ in range(3)]
    lens = [C.c_long(0) for _ in range(3)]
    r = lib().fxo_literals(pattern, len(pattern), bufs[0], C.byref(lens[0]), bufs[1], C.byref(lens[1]),
                           bufs[2], C.byref(lens[2]), cap)
    if r != 0:
        return None
    return tuple(bufs[i].raw[:lens[i].value] for i in range(3))


def op_in(pattern: bytes, text: bytes):
    return lib().fxo_in(pattern, len(pattern), text, len(text))


def op_match(pattern: bytes, text: bytes):
    return lib().fxo_match(pattern, len(pattern), text, len(text))


def regex(pattern: bytes, text: bytes):
    """-> (res, length, from, to, status) like the reference's `regex` subroutine."""
    f, t, l, st = C.c_long(0), C.c_long(0), C.c_long(0), C.c_int(0)
    r = lib().fxo_regex(pattern, len(pattern), text, len(text), C.byref(f), C.byref(t), C.byref(l), C.byref(st))
    if r != 0:
        return None, 0, 0, 0, r
    res = text[f.value - 1:t.value] if f.value > 0 and t.value > 0 else b""
    return res, l.value, f.value, t.value, st.value


class Compiled:
    """Pattern compiled once (mode 0 = .in./regex preprocessing, 1 = .match. preprocessing)."""

    def __init__(self, pattern: bytes, mode: int):
        st = C.c_int(0)
        self.h = lib().fxo_compile(pattern, len(pattern), mode, C.byref(st))
        self.status = st.value
        self.mode = mode
        if not self.h:
            raise RuntimeError("oracle compile aborted with %d" % st.value)

    def close(self):
        if self.h:
            lib().fxo_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def steps(self):
        return lib().fxo_steps(self.h)

    def dfa_states(self):
        return lib().fxo_dfa_states(self.h)

    def bool_batch(self, op, buf, offsets):
        import numpy as np
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=np.uint8)
        r = lib().fxo_bool_batch(self.h, op, buf.ctypes.data, offsets.ctypes.data, n, out.ctypes.data)
        if r != 0:
            raise RuntimeError("oracle aborted with %d" % r)
        return out

    def bool_fixed(self, op, buf, n, stride):
        import numpy as np
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        out = np.zeros(n, dtype=np.uint8)
        r = lib().fxo_bool_fixed(self.h, op, buf.ctypes.data, n, stride, out.ctypes.data)
        if r != 0:
            raise RuntimeError("oracle aborted with %d" % r)
        return out

    def regex_batch(self, buf, offsets):
        import numpy as np
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        f = np.zeros(n, dtype=np.int64)
        t = np.zeros(n, dtype=np.int64)
        r = lib().fxo_regex_batch(self.h, buf.ctypes.data, offsets.ctypes.data, n, f.ctypes.data, t.ctypes.data)
        if r != 0:
            raise RuntimeError("oracle aborted with %d" % r)
        return f, t

    def regex_buffer(self, buf):
        import numpy as np
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        f, t = C.c_longlong(0), C.c_longlong(0)
        r = lib().fxo_regex_buffer(self.h, buf.ctypes.data, len(buf), C.byref(f), C.byref(t))
        if r != 0:
            raise RuntimeError("oracle aborted with %d" % r)
        return f.value, t.value
