"""ctypes binding of libtfw_b200.so (include/tfw_worker.h, tfw_gate.h, tfw_trace.h).

This is the Python stand-in for the cgo stub a Go maintainer would write
(INTEGRATION.md); it only forwards to the C-ABI.  There is no Python or CPU
implementation behind it: if the library is missing the import fails, and if
no CUDA device is present every data-path call raises ``NoDeviceError``.
"""
import ctypes as C
import os

from . import PACKAGE_DIR

LIB_PATH = os.path.join(PACKAGE_DIR, "lib", "libtfw_b200.so")

TFW_OK, TFW_ERR_INVALID, TFW_ERR_NOT_FOUND, TFW_ERR_NOT_SUPPORTED, TFW_ERR_EXHAUSTED = 0, 1, 2, 3, 4
TFW_ERR_FAILED, TFW_ERR_INTERNAL, TFW_ERR_PROTOCOL, TFW_ERR_NO_DEVICE = 5, 6, 7, 8
STATUS_NAMES = {
    0: "OK", 1: "INVALID", 2: "NOT_FOUND", 3: "NOT_SUPPORTED", 4: "EXHAUSTED",
    5: "FAILED", 6: "INTERNAL", 7: "PROTOCOL", 8: "NO_DEVICE",
}

TFW_F_MOVER_TMA, TFW_F_MOVER_LDG, TFW_F_NO_ZERO_FILL, TFW_F_NO_LIMITER, TFW_F_GATE_FAIL_CLOSED = 0x1, 0x2, 0x4, 0x8, 0x10


class TfwError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: {STATUS_NAMES.get(status, status)}{(' - ' + detail) if detail else ''}")


class NoDeviceError(TfwError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("chunk_bytes", C.c_uint64),
        ("num_slots", C.c_uint32), ("flags", C.c_uint32), ("vram_limit_bytes", C.c_uint64),
        ("shm_path", C.c_char_p), ("shm_device_index", C.c_uint32), ("mover_ctas_per_sm", C.c_uint32),
        ("tiering", C.c_void_p), ("sm_percent_limit", C.c_uint32), ("reserved0", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "frames", "payload_bytes", "d2h_bytes", "d2d_bytes", "fill_bytes", "h2d_dma_bytes", "mover_launches",
        "gate_launches", "client_launches", "batches_hazard", "vram_bytes", "vram_peak_bytes", "live_buffers",
        "other_launches", "h2d_ref_bytes", "d2h_ref_bytes", "user_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class ResponseSink(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("ring", C.c_void_p), ("ring_bytes", C.c_uint64),
                ("head", C.POINTER(C.c_uint64)), ("tail", C.POINTER(C.c_uint64))]


class MoveDesc(C.Structure):
    _fields_ = [("dst", C.c_uint64), ("src", C.c_uint64), ("len", C.c_uint64), ("tile0", C.c_uint32), ("fill", C.c_uint32)]


class GateState(C.Structure):
    _fields_ = [("tokens", C.c_double), ("capacity", C.c_double), ("refill_rate", C.c_double),
                ("admitted", C.c_uint64), ("denied", C.c_uint64), ("blocked_gates", C.c_uint64),
                ("wait_ns", C.c_uint64), ("bridged_tokens_milli", C.c_uint64), ("timeouts", C.c_uint64)]


class GateOp(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("pad", C.c_uint32), ("amount", C.c_double)]


class C1Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("ncalls", C.c_uint32), ("max_live", C.c_uint32),
                ("max_buffer_bytes", C.c_uint64), ("max_payload_bytes", C.c_uint64),
                ("unaligned_percent", C.c_uint32), ("error_permille", C.c_uint32),
                ("launch_cost", C.c_uint32), ("reserved", C.c_uint32)]


class VspaceConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("home_device", C.c_int32), ("va_bytes", C.c_uint64),
                ("region_bytes", C.c_uint64), ("home_budget_bytes", C.c_uint64), ("peer_budget_bytes", C.c_uint64),
                ("host_budget_bytes", C.c_uint64), ("peer_devices", C.c_int32 * 15), ("n_peers", C.c_uint32),
                ("flags", C.c_uint32), ("prefetch_ahead", C.c_uint32)]


class VspaceStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "regions_home", "regions_peer", "regions_host", "evict_bytes_peer", "prefetch_bytes_peer", "evict_bytes_host",
        "prefetch_bytes_host", "mover_launches", "remaps", "policy_evictions", "policy_prefetches", "policy_hits",
        "policy_hits_inflight", "policy_prefetch_ahead", "stall_ns", "phys_created", "phys_destroyed", "vmm_ns")]


class MigrateResult(C.Structure):
    _fields_ = [("bytes", C.c_uint64), ("copy_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32), ("pad", C.c_uint32)]


_P = C.c_void_p
_SIGS = {
    "tfw_abi_version": (C.c_uint32, []),
    "tfw_worker_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "tfw_worker_destroy": (C.c_int, [_P]),
    "tfw_last_error": (C.c_char_p, [_P]),
    "tfw_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "tfw_host_free": (C.c_int, [_P]),
    "tfw_host_register": (C.c_int, [_P, C.c_size_t]),
    "tfw_host_unregister": (C.c_int, [_P]),
    "tfw_submit": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tfw_set_arena_prefix": (C.c_int, [_P, C.c_char_p]),
    "tfw_set_response_sink": (C.c_int, [_P, C.POINTER(ResponseSink)]),
    "tfw_flush": (C.c_int, [_P]),
    "tfw_worker_freeze": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "tfw_worker_resume": (C.c_int, [_P]),
    "tfw_worker_poll_control": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "tfw_worker_auto_freeze": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "tfw_worker_auto_resume": (C.c_int, [_P]),
    "tfw_worker_set_sm_limit": (C.c_int, [_P, C.c_uint32]),
    "tfw_worker_set_vram_limit": (C.c_int, [_P, C.c_uint64]),
    "tfw_fence_query": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int)]),
    "tfw_fence": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "tfw_fence_wait": (C.c_int, [_P, C.c_uint64]),
    "tfw_poll_responses": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tfw_trace_load": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(_P)]),
    "tfw_trace_replay": (C.c_int, [_P, _P]),
    "tfw_trace_info": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tfw_trace_free": (C.c_int, [_P, _P]),
    "tfw_trace_buffer_info": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tfw_buffer_info": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tfw_buffer_read": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P, C.c_uint64]),
    "tfw_buffer_digest": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint64)]),
    "tfw_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "tfw_exec_stream": (_P, [_P]),
    "tfw_move_batch": (C.c_int, [_P, C.POINTER(MoveDesc), C.c_uint32, C.POINTER(C.c_float)]),
    "tfw_dev_alloc": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "tfw_dev_free": (C.c_int, [_P, C.c_uint64]),
    "tfw_dev_write": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "tfw_dev_read": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64]),
    "tfw_dev_digest": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    # gate
    "tfw_gate_create": (C.c_int, [C.c_int, C.c_char_p, C.c_uint32, C.POINTER(_P)]),
    "tfw_gate_destroy": (C.c_int, [_P]),
    "tfw_gate_set_policy": (C.c_int, [_P, C.c_int, C.c_double]),
    "tfw_gate_try": (C.c_int, [_P, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "tfw_gate_enqueue": (C.c_int, [_P, C.c_double, _P]),
    "tfw_gate_refill": (C.c_int, [_P, C.c_double, C.POINTER(C.c_double)]),
    "tfw_gate_set_capacity": (C.c_int, [_P, C.c_double]),
    "tfw_gate_set_tokens": (C.c_int, [_P, C.c_double]),
    "tfw_gate_get_state": (C.c_int, [_P, C.POINTER(GateState)]),
    "tfw_worker_gate_state": (C.c_int, [_P, C.POINTER(GateState)]),
    "tfw_gate_run_sequence": (C.c_int, [_P, C.POINTER(GateOp), C.c_uint32, C.POINTER(C.c_double)]),
    "tfw_gate_contend": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_double, C.POINTER(C.c_uint64)]),
    # VRAM tiering
    "tfw_vspace_create": (C.c_int, [C.POINTER(VspaceConfig), C.POINTER(_P)]),
    "tfw_vspace_destroy": (C.c_int, [_P]),
    "tfw_vspace_last_error": (C.c_char_p, [_P]),
    "tfw_vspace_info": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "tfw_vspace_populate": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_int32]),
    "tfw_vspace_migrate": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.c_uint32, C.POINTER(MigrateResult)]),
    "tfw_vspace_residency": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "tfw_vspace_access": (C.c_int, [_P, C.c_uint32]),
    "tfw_vspace_bind_stream": (C.c_int, [_P, _P]),
    "tfw_vspace_quiesce": (C.c_int, [_P]),
    "tfw_vspace_sweep": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "tfw_vspace_unpopulate": (C.c_int, [_P, C.c_uint32]),
    "tfw_vspace_pin": (C.c_int, [_P, C.c_uint32, C.c_int]),
    "tfw_vspace_get_stats": (C.c_int, [_P, C.POINTER(VspaceStats)]),
    "tfw_vspace_fill_pattern": (C.c_int, [_P, C.c_uint32, C.c_uint64]),
    "tfw_vspace_digest": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint64)]),
    "tfw_vspace_read": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P, C.c_uint64]),
    "tfw_vspace_write": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P, C.c_uint64]),
    # trace generators
    "tfw_trace_c1_defaults": (None, [C.POINTER(C1Params)]),
    "tfw_trace_gen_c1": (C.c_int, [C.POINTER(C1Params), _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tfw_trace_gen_bulk": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tfw_trace_gen_small": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint64, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tfw_trace_payload": (None, [C.c_uint64, C.c_uint32, _P, C.c_uint64]),
    "tfw_native_replay": (C.c_int, [C.c_int, _P, C.c_size_t, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tfw_native_copy": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]),
}

DECLARED_SYMBOLS = tuple(_SIGS)


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
            "there is no Python/CPU fallback for the vGPU worker data path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


def check(status, where, worker=None):
    if status == TFW_OK:
        return
    detail = ""
    if worker:
        detail = (lib.tfw_last_error(worker) or b"").decode()
    if status == TFW_ERR_NO_DEVICE:
        raise NoDeviceError(status, where, "no CUDA device: the B200 worker has no CPU fallback")
    raise TfwError(status, where, detail)
