"""Synthetic TFCS traces (include/tfw_trace.h)."""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import lib, check

SEED_C1 = 0x7F5EED


def c1_params(**over):
    p = N.C1Params()
    lib.tfw_trace_c1_defaults(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _two_pass(fn, into=None):
    n = C.c_size_t()
    check(fn(None, 0, C.byref(n)), "trace size")
    if into is None:
        buf = np.zeros(n.value, dtype=np.uint8)
        ptr = buf.ctypes.data
    else:
        if into.nbytes < n.value:
            raise ValueError(f"trace needs {n.value} bytes, buffer has {into.nbytes}")
        buf, ptr = into.array[: n.value], into.ptr
    check(fn(C.c_void_p(ptr), n.value, C.byref(n)), "trace gen")
    return buf


def gen_c1(into=None, **over):
    """The 1k-call mixed trace of BASELINE config 1 (seed 0x7F5EED)."""
    p = c1_params(**over)
    return _two_pass(lambda out, cap, n: lib.tfw_trace_gen_c1(C.byref(p), out, cap, n), into)


def gen_bulk(nbuf, ncopies, bytes_each, seed=SEED_C1, nthreads=8, into=None):
    return _two_pass(lambda out, cap, n: lib.tfw_trace_gen_bulk(seed, nbuf, ncopies, bytes_each, nthreads, out, cap, n), into)


def bulk_size(nbuf, ncopies, bytes_each):
    n = C.c_size_t()
    check(lib.tfw_trace_gen_bulk(0, nbuf, ncopies, bytes_each, 1, None, 0, C.byref(n)), "trace size")
    return n.value


def gen_small(ncalls, bytes_each, seed=SEED_C1, into=None):
    return _two_pass(lambda out, cap, n: lib.tfw_trace_gen_small(seed, ncalls, bytes_each, out, cap, n), into)


def payload(seed, call_id, nbytes):
    out = np.empty(nbytes, dtype=np.uint8)
    lib.tfw_trace_payload(seed, call_id, C.c_void_p(out.ctypes.data), nbytes)
    return out


def native_replay(data, passes=3, device=0):
    """Native-CUDA comparator: returns (seconds_per_pass, payload_bytes, calls)."""
    arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data
    sec, pay, calls = C.c_double(), C.c_uint64(), C.c_uint64()
    check(lib.tfw_native_replay(device, C.c_void_p(arr.ctypes.data), arr.nbytes, passes, C.byref(sec), C.byref(pay), C.byref(calls)), "tfw_native_replay")
    return sec.value, pay.value, calls.value
