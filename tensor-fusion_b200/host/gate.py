"""Device-resident ERL token bucket (include/tfw_gate.h)."""
import ctypes as C

from . import _native as N
from ._native import lib, check


class Gate:
    def __init__(self, device=0, shm_path=None, device_index=0):
        h = C.c_void_p()
        check(lib.tfw_gate_create(device, shm_path.encode() if shm_path else None, device_index, C.byref(h)), "tfw_gate_create")
        self.h = h

    def close(self):
        if self.h:
            lib.tfw_gate_destroy(self.h)
            self.h = None

    def try_acquire(self, cost):
        before, adm = C.c_double(), C.c_int()
        check(lib.tfw_gate_try(self.h, cost, C.byref(before), C.byref(adm)), "tfw_gate_try")
        return before.value, bool(adm.value)

    def enqueue(self, cost, stream):
        check(lib.tfw_gate_enqueue(self.h, cost, C.c_void_p(stream)), "tfw_gate_enqueue")

    def refill(self, amount):
        before = C.c_double()
        check(lib.tfw_gate_refill(self.h, amount, C.byref(before)), "tfw_gate_refill")
        return before.value

    def set_capacity(self, v): check(lib.tfw_gate_set_capacity(self.h, v), "tfw_gate_set_capacity")
    def set_tokens(self, v): check(lib.tfw_gate_set_tokens(self.h, v), "tfw_gate_set_tokens")

    def state(self):
        s = N.GateState()
        check(lib.tfw_gate_get_state(self.h, C.byref(s)), "tfw_gate_get_state")
        return {n: getattr(s, n) for n, _ in s._fields_}

    def run_sequence(self, ops):
        """ops: list of (kind, amount); returns the value found before each op."""
        arr = (N.GateOp * len(ops))()
        for i, (k, a) in enumerate(ops):
            arr[i].kind, arr[i].amount = k, a
        out = (C.c_double * len(ops))()
        check(lib.tfw_gate_run_sequence(self.h, arr, len(ops), out), "tfw_gate_run_sequence")
        return list(out)

    def contend(self, nthreads, per_thread, cost):
        a = C.c_uint64()
        check(lib.tfw_gate_contend(self.h, nthreads, per_thread, cost, C.byref(a)), "tfw_gate_contend")
        return a.value
