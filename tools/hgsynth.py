"""ctypes wrapper of the fixture generator (tools/hg_synth.cpp) for tests and bench.py."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libhinge_synth.so")
COLS = ["aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "flags"]


class Params(C.Structure):
    _fields_ = [("genome_len", C.c_int64), ("coverage", C.c_double),
                ("read_mean", C.c_int32), ("read_sd", C.c_int32), ("read_min", C.c_int32), ("read_max", C.c_int32),
                ("n_families", C.c_int32), ("rep_min_len", C.c_int32), ("rep_max_len", C.c_int32),
                ("rep_min_copies", C.c_int32), ("rep_max_copies", C.c_int32),
                ("min_ovl", C.c_int32), ("jitter", C.c_int32), ("tspace", C.c_int32),
                ("qv_bad_frac", C.c_double), ("seed", C.c_uint64), ("frag_prob", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.hgs_create.restype = C.c_void_p
        L.hgs_create.argtypes = [C.POINTER(Params)]
        L.hgs_destroy.argtypes = [C.c_void_p]
        L.hgs_n_read.argtypes = [C.c_void_p]
        L.hgs_generate.restype = C.c_int64
        L.hgs_generate.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        for n, t in (("hgs_rlen", C.c_int32), ("hgs_qv_off", C.c_int64), ("hgs_qv", C.c_uint8),
                     ("hgs_trace_off", C.c_int64), ("hgs_trace", C.c_uint8)):
            getattr(L, n).restype = C.POINTER(t)
            getattr(L, n).argtypes = [C.c_void_p]
        L.hgs_col.restype = C.POINTER(C.c_int32)
        L.hgs_col.argtypes = [C.c_void_p, C.c_int32]
        L.hgs_trace_bytes.restype = C.c_int64
        L.hgs_trace_bytes.argtypes = [C.c_void_p]
        L.hgs_write_db.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32]
        L.hgs_write_las.argtypes = [C.c_void_p, C.c_char_p]
        L.hgs_default_params.argtypes = [C.POINTER(Params)]
        _lib = L
    return _lib


class Synth:
    def __init__(self, **kw):
        L = lib()
        self.p = Params()
        L.hgs_default_params(C.byref(self.p))
        for k, v in kw.items():
            setattr(self.p, k, v)
        self.h = L.hgs_create(C.byref(self.p))
        self.n_read = L.hgs_n_read(self.h)
        self.novl = 0

    def close(self):
        if self.h:
            lib().hgs_destroy(self.h)
            self.h = None

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,))

    @property
    def rlen(self):
        return self._arr(lib().hgs_rlen(self.h), self.n_read, np.int32)

    @property
    def qv_off(self):
        return self._arr(lib().hgs_qv_off(self.h), self.n_read + 1, np.int64)

    @property
    def qv(self):
        return self._arr(lib().hgs_qv(self.h), int(self.qv_off[-1]), np.uint8)

    def generate(self, a_lo=0, a_hi=None, want_trace=True, threads=8):
        a_hi = self.n_read if a_hi is None else a_hi
        self.novl = int(lib().hgs_generate(self.h, a_lo, a_hi, int(want_trace), threads))
        return self.novl

    def cols(self):
        """Views into the generator's buffers (valid until the next generate())."""
        return {name: self._arr(lib().hgs_col(self.h, i), self.novl, np.int32) for i, name in enumerate(COLS)}

    def trace(self):
        off = self._arr(lib().hgs_trace_off(self.h), self.novl + 1, np.int64)
        nb = int(lib().hgs_trace_bytes(self.h))
        return off, self._arr(lib().hgs_trace(self.h), nb, np.uint8)

    def write_db(self, directory, root, with_bps=True, with_qv=True):
        os.makedirs(directory, exist_ok=True)
        assert lib().hgs_write_db(self.h, directory.encode(), root.encode(), int(with_bps), int(with_qv)) == 0

    def write_las(self, path):
        assert lib().hgs_write_las(self.h, path.encode()) == 0
