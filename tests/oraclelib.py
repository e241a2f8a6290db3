"""ctypes wrapper of the oracle's array interface (oracle/oracle_capi.cpp).  TEST-ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libhinge_oracle.so")
INI = os.path.join(ROOT, "tests", "golden", "nominal.ini")

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64] + \
            [C.c_void_p] * 7 + [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_load_ini.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_filter.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_filter_results.restype = C.c_int64
        L.orc_filter_results.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        _lib = L
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Oracle:
    def __init__(self, rlen, qv_off, qv, tspace, cols, trace_off=None, trace=None, tbytes=1, ini=INI, threads=1):
        L = lib()
        self.n = len(rlen)
        keep = [np.ascontiguousarray(rlen, np.int32)]
        if qv_off is not None:
            keep += [np.ascontiguousarray(qv_off, np.int64), np.ascontiguousarray(qv, np.uint8)]
        else:
            keep += [None, None]
        cc = [np.ascontiguousarray(cols[k], np.int32) for k in
              ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "flags")]
        self.h = L.orc_create(self.n, _p(keep[0]), _p(keep[1]), _p(keep[2]), tspace, len(cc[0]),
                              *[_p(c) for c in cc], _p(trace_off), _p(trace), tbytes)
        assert L.orc_load_ini(self.h, ini.encode()) == 0
        L.orc_set_threads(self.h, int(threads))

    def filter(self):
        L = lib()
        L.orc_filter(self.h)
        n = self.n
        summary = np.zeros(4, np.int32)
        total = L.orc_filter_results(self.h, None, None, None, None, None, None, None, _p(summary))
        out = {"mask": np.zeros((n, 2), np.int32), "cmask": np.zeros((n, 2), np.int32),
               "flags": np.zeros(n, np.uint8), "anno_off": np.zeros(n + 1, np.int64),
               "anno_pos": np.zeros(total + 1, np.int32), "anno_type": np.zeros(total + 1, np.int32),
               "hinge_keep": np.zeros(total + 1, np.uint8)}
        L.orc_filter_results(self.h, _p(out["mask"]), _p(out["cmask"]), _p(out["flags"]), _p(out["anno_off"]),
                             _p(out["anno_pos"]), _p(out["anno_type"]), _p(out["hinge_keep"]), _p(summary))
        for k in ("anno_pos", "anno_type", "hinge_keep"):
            out[k] = out[k][:total]
        out["summary"] = summary
        return out

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None
