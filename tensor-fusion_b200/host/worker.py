"""Host-side mirror of the worker process: drives the C-ABI of libtfw_b200.so.

Mirrors the role of the reference's worker container
(``./tensor-fusion-worker -p 8000``, internal/utils/compose.go:1304-1325): it
receives the forwarded-CUDA byte stream and hands it to the GPU.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import lib, check


class PinnedBuffer:
    """Page-locked host memory the DMA engine reads in place (tfw_host_alloc)."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib.tfw_host_alloc(nbytes, C.byref(p)), "tfw_host_alloc")
        self.ptr = p.value
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            lib.tfw_host_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Trace:
    def __init__(self, worker, handle, keepalive):
        self.worker, self.handle, self._keep = worker, handle, keepalive

    def info(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib.tfw_trace_info(self.handle, C.byref(a), C.byref(b), C.byref(c)), "tfw_trace_info")
        return {"payload_bytes": a.value, "mover_launches": b.value, "algorithmic_bytes": c.value}

    def replay(self):
        check(lib.tfw_trace_replay(self.worker.h, self.handle), "tfw_trace_replay", self.worker.h)

    def buffer_info(self, handle):
        s, p = C.c_uint64(), C.c_uint64()
        check(lib.tfw_trace_buffer_info(self.handle, handle, C.byref(s), C.byref(p)), "tfw_trace_buffer_info")
        return s.value, p.value

    def free(self):
        if self.handle:
            lib.tfw_trace_free(self.worker.h, self.handle)
            self.handle = None


class Worker:
    def __init__(self, device=0, chunk_bytes=0, num_slots=0, flags=0, vram_limit=0, shm_path=None,
                 shm_device_index=0, ctas_per_sm=0, tiering=None, sm_percent=0):
        """tiering: dict(va_bytes, region_bytes, home_budget, peer_budget=0, host_budget=0, peers=(), prefetch_ahead=0) puts the
        client buffers into a tiered vGPU address space (include/tfw_vram.h)."""
        cfg = N.Config()
        cfg.struct_size = C.sizeof(N.Config)
        cfg.device = device
        cfg.chunk_bytes = chunk_bytes
        cfg.num_slots = num_slots
        cfg.flags = flags
        cfg.vram_limit_bytes = vram_limit
        cfg.shm_path = shm_path.encode() if shm_path else None
        cfg.shm_device_index = shm_device_index
        cfg.mover_ctas_per_sm = ctas_per_sm
        cfg.sm_percent_limit = sm_percent
        if tiering:
            vc = N.VspaceConfig()
            vc.struct_size = C.sizeof(N.VspaceConfig)
            vc.home_device = device
            vc.va_bytes, vc.region_bytes = tiering["va_bytes"], tiering["region_bytes"]
            vc.home_budget_bytes = tiering["home_budget"]
            vc.peer_budget_bytes = tiering.get("peer_budget", 0)
            vc.host_budget_bytes = tiering.get("host_budget", 0)
            peers = tiering.get("peers", ())
            for i, p in enumerate(peers):
                vc.peer_devices[i] = p
            vc.n_peers = len(peers)
            vc.flags = tiering.get("flags", 0)
            vc.prefetch_ahead = tiering.get("prefetch_ahead", 0)
            self._vc = vc
            cfg.tiering = C.cast(C.pointer(vc), C.c_void_p)
        h = C.c_void_p()
        check(lib.tfw_worker_create(C.byref(cfg), C.byref(h)), "tfw_worker_create")
        self.h = h

    def close(self):
        if self.h:
            lib.tfw_worker_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- streaming path ----------------------------------------------------
    def submit(self, data):
        """Feed wire bytes (bytes / numpy uint8 / PinnedBuffer slice). Returns bytes consumed."""
        if isinstance(data, (bytes, bytearray)):
            arr = np.frombuffer(data, dtype=np.uint8)
        else:
            arr = data
        consumed = C.c_size_t()
        self._keep = arr
        check(lib.tfw_submit(self.h, C.c_void_p(arr.ctypes.data), arr.nbytes, C.byref(consumed)), "tfw_submit", self.h)
        return consumed.value

    def submit_ptr(self, ptr, nbytes):
        consumed = C.c_size_t()
        check(lib.tfw_submit(self.h, C.c_void_p(ptr), nbytes, C.byref(consumed)), "tfw_submit", self.h)
        return consumed.value

    def freeze(self):
        moved = C.c_uint64()
        check(lib.tfw_worker_freeze(self.h, C.byref(moved)), "tfw_worker_freeze", self.h)
        return moved.value

    def resume(self):
        check(lib.tfw_worker_resume(self.h), "tfw_worker_resume", self.h)

    def flush(self):
        check(lib.tfw_flush(self.h), "tfw_flush", self.h)

    def poll(self, cap=1 << 26):
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t()
        check(lib.tfw_poll_responses(self.h, C.c_void_p(out.ctypes.data), cap, C.byref(n)), "tfw_poll_responses", self.h)
        return out[: n.value].tobytes()

    def run(self, data):
        """submit + flush + drain all responses (convenience for tests)."""
        n = self.submit(data)
        self.flush()
        resp = b""
        while True:
            r = self.poll()
            if not r:
                break
            resp += r
        return n, resp

    # -- resident trace ----------------------------------------------------
    def load_trace(self, data):
        arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data
        t = C.c_void_p()
        check(lib.tfw_trace_load(self.h, C.c_void_p(arr.ctypes.data), arr.nbytes, C.byref(t)), "tfw_trace_load", self.h)
        return Trace(self, t, arr)

    # -- introspection -------------------------------------------------------
    def buffer_info(self, handle):
        s, p = C.c_uint64(), C.c_uint64()
        check(lib.tfw_buffer_info(self.h, handle, C.byref(s), C.byref(p)), "tfw_buffer_info", self.h)
        return s.value, p.value

    def read(self, handle, off=0, n=None):
        if n is None:
            n = self.buffer_info(handle)[0] - off
        out = np.empty(n, dtype=np.uint8)
        check(lib.tfw_buffer_read(self.h, handle, off, C.c_void_p(out.ctypes.data), n), "tfw_buffer_read", self.h)
        return out

    def digest(self, handle):
        d = C.c_uint64()
        check(lib.tfw_buffer_digest(self.h, handle, C.byref(d)), "tfw_buffer_digest", self.h)
        return d.value

    def gate_state(self):
        """Counters of the vGPU's device-resident token bucket ({} if it has no limiter)."""
        g = N.GateState()
        rc = lib.tfw_worker_gate_state(self.h, C.byref(g))
        if rc == N.TFW_ERR_NOT_FOUND:
            return {}
        check(rc, "tfw_worker_gate_state", self.h)
        return {n: getattr(g, n) for n, _ in g._fields_}

    def stats(self):
        s = N.Stats()
        check(lib.tfw_get_stats(self.h, C.byref(s)), "tfw_get_stats", self.h)
        return s.as_dict()

    # -- raw device scratch + the mover on its own ----------------------------
    def dev_alloc(self, n):
        p = C.c_uint64()
        check(lib.tfw_dev_alloc(self.h, n, C.byref(p)), "tfw_dev_alloc", self.h)
        return p.value

    def dev_free(self, p):
        check(lib.tfw_dev_free(self.h, p), "tfw_dev_free", self.h)

    def dev_write(self, p, arr):
        arr = np.ascontiguousarray(arr)
        check(lib.tfw_dev_write(self.h, p, C.c_void_p(arr.ctypes.data), arr.nbytes), "tfw_dev_write", self.h)

    def dev_read(self, p, n):
        out = np.empty(n, dtype=np.uint8)
        check(lib.tfw_dev_read(self.h, p, C.c_void_p(out.ctypes.data), n), "tfw_dev_read", self.h)
        return out

    def dev_digest(self, p, n):
        d = C.c_uint64()
        check(lib.tfw_dev_digest(self.h, p, n, C.byref(d)), "tfw_dev_digest", self.h)
        return d.value

    def move_batch(self, descs, timed=False):
        """descs: list of (dst, src, len, fill_byte); src == 0 means fill."""
        arr = (N.MoveDesc * len(descs))()
        for i, (dst, src, ln, fill) in enumerate(descs):
            arr[i].dst, arr[i].src, arr[i].len, arr[i].fill = dst, src, ln, (fill & 0xFF) * 0x01010101
        ms = C.c_float()
        check(lib.tfw_move_batch(self.h, arr, len(descs), C.byref(ms) if timed else None), "tfw_move_batch", self.h)
        return ms.value if timed else None
