"""vGPU VRAM tiering (include/tfw_vram.h): one address space over home HBM, peer HBM and host DRAM."""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import lib, TfwError, NoDeviceError

NONE, HOME, PEER, HOST = 0, 1, 2, 3
COPY_ENGINE = 0x1
PUSH_EVICT = 0x4
SENDER_DRIVEN = 0x8
PEER_IN_PLACE = 0x10
HOME_DRIVEN = 0x20
REMAP_LATE = 0x40
FIXED_FRAMES = 0x80


class VSpace:
    def __init__(self, home=0, va_bytes=1 << 30, region_bytes=64 << 20, home_budget=0, peer_budget=0, host_budget=0,
                 peers=(), flags=0, prefetch_ahead=0):
        cfg = N.VspaceConfig()
        cfg.struct_size = C.sizeof(N.VspaceConfig)
        cfg.home_device, cfg.va_bytes, cfg.region_bytes = home, va_bytes, region_bytes
        cfg.home_budget_bytes, cfg.peer_budget_bytes, cfg.host_budget_bytes = home_budget, peer_budget, host_budget
        for i, p in enumerate(peers):
            cfg.peer_devices[i] = p
        cfg.n_peers, cfg.flags, cfg.prefetch_ahead = len(peers), flags, prefetch_ahead
        h = C.c_void_p()
        rc = lib.tfw_vspace_create(C.byref(cfg), C.byref(h))
        if rc == N.TFW_ERR_NO_DEVICE:
            raise NoDeviceError(rc, "tfw_vspace_create", "no CUDA device: VRAM tiering has no CPU fallback")
        if rc:
            raise TfwError(rc, "tfw_vspace_create")
        self.h = h
        b, r, n = C.c_uint64(), C.c_uint64(), C.c_uint32()
        self._ck(lib.tfw_vspace_info(h, C.byref(b), C.byref(r), C.byref(n)), "info")
        self.base, self.region_bytes, self.n_regions = b.value, r.value, n.value

    def _ck(self, rc, where):
        if rc:
            raise TfwError(rc, "tfw_vspace_" + where, (lib.tfw_vspace_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            lib.tfw_vspace_destroy(self.h)
            self.h = None

    def __enter__(self): return self
    def __exit__(self, *a): self.close()

    def populate(self, region, tier, peer_slot=-1): self._ck(lib.tfw_vspace_populate(self.h, region, tier, peer_slot), "populate")

    def migrate(self, regions, tiers, peer_slots=None):
        n = len(regions)
        r = (C.c_uint32 * n)(*regions)
        t = (C.c_uint8 * n)(*tiers)
        s = (C.c_int32 * n)(*(peer_slots if peer_slots is not None else [-1] * n))
        res = N.MigrateResult()
        self._ck(lib.tfw_vspace_migrate(self.h, r, t, s, n, C.byref(res)), "migrate")
        return {"bytes": res.bytes, "copy_ms": res.copy_ms, "total_ms": res.total_ms, "launches": res.launches}

    def residency(self, region):
        t, d = C.c_uint32(), C.c_int32()
        self._ck(lib.tfw_vspace_residency(self.h, region, C.byref(t), C.byref(d)), "residency")
        return t.value, d.value

    def access(self, region): self._ck(lib.tfw_vspace_access(self.h, region), "access")

    def bind_stream(self, cuda_stream): self._ck(lib.tfw_vspace_bind_stream(self.h, C.c_void_p(cuda_stream)), "bind_stream")

    def quiesce(self): self._ck(lib.tfw_vspace_quiesce(self.h), "quiesce")

    def sweep(self, first, count):
        """The policy path in one native loop: access + a digest kernel per region; returns (digests, seconds)."""
        d = (C.c_uint64 * count)()
        s = C.c_double()
        self._ck(lib.tfw_vspace_sweep(self.h, first, count, d, C.byref(s)), "sweep")
        return list(d), s.value

    def stats(self):
        s = N.VspaceStats()
        self._ck(lib.tfw_vspace_get_stats(self.h, C.byref(s)), "get_stats")
        return {n: int(getattr(s, n)) for n, _ in s._fields_}

    def fill_pattern(self, region, seed): self._ck(lib.tfw_vspace_fill_pattern(self.h, region, seed), "fill_pattern")

    def digest(self, region):
        d = C.c_uint64()
        self._ck(lib.tfw_vspace_digest(self.h, region, C.byref(d)), "digest")
        return d.value

    def read(self, region, off, n):
        out = np.empty(n, dtype=np.uint8)
        self._ck(lib.tfw_vspace_read(self.h, region, off, C.c_void_p(out.ctypes.data), n), "read")
        return out

    def write(self, region, off, arr):
        arr = np.ascontiguousarray(arr)
        self._ck(lib.tfw_vspace_write(self.h, region, off, C.c_void_p(arr.ctypes.data), arr.nbytes), "write")
