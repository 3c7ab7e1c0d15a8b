"""libtfc_client.so over the shared-memory rings (include/tfw_shm_ring.h), CPU only: a Python thread plays
the worker's half of the protocol (attach, consume the client->worker ring, answer through the
worker->client ring, session hand-over) with rings small enough that every cursor wraps many times."""
import ctypes as C
import mmap
import os
import threading
import time

import numpy as np
import pytest

import conftest
from tensor_fusion_b200 import shm_ring as R
from tensor_fusion_b200 import wire

LIB = os.path.join(conftest.ROOT, "tensor-fusion_b200", "lib", "libtfc_client.so")


def client_lib():
    lib = C.CDLL(LIB)
    lib.tfc_connect.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.tfc_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_memcpy_h2d.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
    lib.tfc_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_memset.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_uint64]
    lib.tfc_sync.argtypes = [C.c_void_p]
    lib.tfc_close.argtypes = [C.c_void_p]
    return lib


class FakeWorker(threading.Thread):
    """The worker's side of the rings; executes MALLOC / H2D / MEMSET / D2H / SYNC on numpy buffers."""

    def __init__(self, path, total, sessions=1, vram_quota=None, stall=0.0):
        super().__init__(daemon=True)
        self.vram_quota = vram_quota
        self.stall = stall            # seconds the consumer sleeps at the start of every session
        self.f = open(path, "w+b")
        self.f.truncate(total)
        self.mm = mmap.mmap(self.f.fileno(), total)
        self.h = R.TfsrHeader.from_buffer(self.mm)
        c2w_off, c2w_size, w2c_off, w2c_size = R.layout(total)
        self.h.version, self.h.total_bytes = R.TFSR_VERSION, total
        self.h.c2w_off, self.h.c2w_size, self.h.w2c_off, self.h.w2c_size = c2w_off, c2w_size, w2c_off, w2c_size
        self.h.worker_pid, self.h.session = os.getpid(), 1
        self.h.magic = R.TFSR_MAGIC
        self.h.worker_ready = 1
        self.sessions, self.bytes_in, self.frames, self.stop = sessions, 0, 0, False
        self.max_inflight = 0
        self.launches = []
        self.path = path
        self.arenas, self.modules, self.functions, self.user_launches = {}, {}, {}, []
        self.payload_bytes_in = 0     # bytes that travelled as frame payloads (by-reference copies add nothing)

    # parameter layouts of the kernels of tools/user_kernels.cu, as cuFuncGetParamInfo reports them
    KERNELS = {b"saxpy_u32": [(0, 8), (8, 8), (16, 4), (20, 4)], b"vec_add_struct": [(0, 32)]}

    def _deref(self, bufs, ptr):
        assert ptr >> 62 == 1, hex(ptr)
        return bufs[(ptr >> 40) & 0x3FFFFF], ptr & ((1 << 40) - 1)

    def _run_user_kernel(self, name, params, bufs):
        import struct
        if name == b"saxpy_u32":
            px, py, a, n = struct.unpack_from("<QQII", params)
            (bx, ox), (by, oy) = self._deref(bufs, px), self._deref(bufs, py)
            x = bx[ox:ox + 4 * n].view(np.uint32)
            y = by[oy:oy + 4 * n].view(np.uint32)
            y[:] = (np.uint32(a) * x + y)
        elif name == b"vec_add_struct":
            n, bias, pa, pb, po = struct.unpack_from("<IIQQQ", params)
            (ba, oa), (bb, ob), (bo, oo) = self._deref(bufs, pa), self._deref(bufs, pb), self._deref(bufs, po)
            bo[oo:oo + 4 * n].view(np.uint32)[:] = ba[oa:oa + 4 * n].view(np.uint32) + bb[ob:ob + 4 * n].view(np.uint32) + np.uint32(bias)

    def _send(self, data):
        h, off, size = self.h, self.h.w2c_off, self.h.w2c_size
        view = memoryview(data)
        while len(view):
            free = size - (h.w2c_head - h.w2c_tail)
            if not free:
                if self.stop:
                    return
                time.sleep(0.0002)
                continue
            pos = h.w2c_head % size
            k = min(len(view), free, size - pos)
            self.mm[off + pos:off + pos + k] = view[:k]
            h.w2c_head += k
            view = view[k:]

    def run(self):
        h = self.h
        for _ in range(self.sessions):
            while h.client_pid == 0 and not self.stop:
                time.sleep(0.0005)
            session, bufs, stream = h.session, {}, bytearray()
            time.sleep(self.stall)
            while not self.stop:
                head = h.c2w_head
                avail = head - h.c2w_tail
                self.max_inflight = max(self.max_inflight, avail)
                if avail:
                    pos = h.c2w_tail % h.c2w_size
                    k = min(avail, h.c2w_size - pos)
                    stream += self.mm[h.c2w_off + pos:h.c2w_off + pos + k]
                    h.c2w_tail += k
                    self.bytes_in += k
                    stream = self._execute(stream, bufs)
                elif h.client_closed >= session:
                    break
                else:
                    time.sleep(0.0002)
            h.worker_closed = session
            h.session = session + 1
            h.client_pid = 0

    def _execute(self, stream, bufs):
        while len(stream) >= 64:
            hdr = wire.unpack_header(bytes(stream[:64]))
            pay = wire.pad16(hdr["length"]) if hdr["opcode"] in wire.PAYLOAD_OPS else 0
            if len(stream) < 64 + pay:
                break
            self.frames += 1
            self.payload_bytes_in += pay
            op = hdr["opcode"]
            err = lambda code: self._send(wire.frame(wire.OP_RESP_ERROR, call_id=hdr["call_id"], arg0=code, arg1=op))
            if op == wire.OP_MALLOC:
                if self.vram_quota is not None and sum(len(b) for b in bufs.values()) + hdr["length"] > self.vram_quota:
                    self._send(wire.frame(wire.OP_RESP_ERROR, call_id=hdr["call_id"], arg0=4, arg1=op))   # TFW_ERR_EXHAUSTED
                else:
                    bufs[hdr["h0"]] = np.zeros(hdr["length"], dtype=np.uint8)
            elif op == wire.OP_FREE:
                if bufs.pop(hdr["h0"], None) is None:
                    self._send(wire.frame(wire.OP_RESP_ERROR, call_id=hdr["call_id"], arg0=2, arg1=op))
            elif op == wire.OP_D2D:
                d, sbuf = bufs[hdr["h0"]], bufs[hdr["h1"]]
                d[hdr["off0"]:hdr["off0"] + hdr["length"]] = sbuf[hdr["off1"]:hdr["off1"] + hdr["length"]].copy()
            elif op == wire.OP_LAUNCH:
                self.launches.append((hdr["arg0"], hdr["arg1"], hdr["arg2"], hdr["arg3"]))
                if hdr["length"]:
                    r = bufs[hdr["h0"]][hdr["off0"]:hdr["off0"] + hdr["length"]]
                    if hdr["arg0"] == wire.K_ADD_U8:
                        r += np.uint8(hdr["off1"] & 0xff)
                    elif hdr["arg0"] == wire.K_XOR_IDX:
                        r ^= ((np.arange(len(r), dtype=np.uint64) * np.uint64(hdr["off1"])) >> np.uint64(3)).astype(np.uint8)
            elif op == wire.OP_H2D:
                bufs[hdr["h0"]][hdr["off0"]:hdr["off0"] + hdr["length"]] = np.frombuffer(bytes(stream[64:64 + hdr["length"]]), dtype=np.uint8)
            elif op == wire.OP_MEMSET:
                bufs[hdr["h0"]][hdr["off0"]:hdr["off0"] + hdr["length"]] = hdr["arg0"]
            elif op == wire.OP_D2H:
                if hdr["h0"] not in bufs:
                    self._send(wire.frame(wire.OP_RESP_ERROR, call_id=hdr["call_id"], arg0=2, arg1=op))
                else:
                    data = bufs[hdr["h0"]][hdr["off0"]:hdr["off0"] + hdr["length"]].tobytes()
                    self._send(wire.frame(wire.OP_RESP_D2H, call_id=hdr["call_id"], h0=hdr["h0"], off0=hdr["off0"], length=len(data), payload=data))
            elif op == wire.OP_SYNC:
                self._send(wire.frame(wire.OP_RESP_SYNC, call_id=hdr["call_id"]))
            elif op == wire.OP_UPGRADE_SHM:
                self.upgrade_offers = getattr(self, "upgrade_offers", 0) + 1
                err(3)                                   # this stand-in stays on the socket
            elif op == wire.OP_HOST_REGISTER:
                try:
                    f = open(f"{self.path}.a{hdr['h0']}", "r+b")
                    if os.fstat(f.fileno()).st_size < hdr["length"] or hdr["h0"] in self.arenas:
                        raise OSError
                    self.arenas[hdr["h0"]] = np.frombuffer(mmap.mmap(f.fileno(), hdr["length"]), dtype=np.uint8)
                except OSError:
                    err(2)
            elif op == wire.OP_HOST_UNREGISTER:
                if self.arenas.pop(hdr["h0"], None) is None:
                    err(2)
            elif op in (wire.OP_H2D_REF, wire.OP_D2H_REF):
                a, b = self.arenas.get(hdr["h1"]), bufs.get(hdr["h0"])
                if a is None or b is None:
                    err(2)
                elif hdr["off1"] + hdr["length"] > len(a) or hdr["off0"] + hdr["length"] > len(b):
                    err(1)
                elif op == wire.OP_H2D_REF:
                    b[hdr["off0"]:hdr["off0"] + hdr["length"]] = a[hdr["off1"]:hdr["off1"] + hdr["length"]]
                else:
                    a[hdr["off1"]:hdr["off1"] + hdr["length"]] = b[hdr["off0"]:hdr["off0"] + hdr["length"]]
                    if hdr["flags"] & wire.F_ACK:
                        self._send(wire.frame(wire.OP_RESP_ACK, call_id=hdr["call_id"]))
            elif op == wire.OP_MODULE_LOAD:
                image = bytes(stream[64:64 + hdr["length"]])
                if image.startswith(b"this is not"):
                    err(1)
                else:
                    self.modules[hdr["h0"]] = image
                    self.loaded_images = getattr(self, "loaded_images", []) + [image]
            elif op == wire.OP_MODULE_UNLOAD:
                if self.modules.pop(hdr["h0"], None) is None:
                    err(2)
            elif op == wire.OP_MODULE_GET_FUNCTION:
                name = bytes(stream[64:64 + hdr["length"]])
                if hdr["h0"] not in self.modules or name not in self.KERNELS:
                    err(2)
                else:
                    lay = self.KERNELS[name]
                    self.functions[hdr["h1"]] = name
                    table = b"".join(int(o).to_bytes(4, "little") + int(sz).to_bytes(4, "little") for o, sz in lay)
                    self._send(wire.frame(wire.OP_RESP_FUNCTION, call_id=hdr["call_id"], h0=hdr["h0"], h1=hdr["h1"], length=len(table), arg0=len(lay),
                                          arg1=max(o + sz for o, sz in lay), payload=table))
            elif op == wire.OP_LAUNCH_USER:
                import struct
                blob = bytes(stream[64:64 + hdr["length"]])
                geo = struct.unpack_from("<8I", blob)
                name = self.functions.get(hdr["h1"])
                if name is None:
                    err(2)
                else:
                    self.user_launches.append((name, geo[:3], geo[3:6], geo[6], geo[7], hdr["arg3"]))
                    self._run_user_kernel(name, blob[32:], bufs)
            stream = stream[64 + pay:]
        return stream


@pytest.fixture
def shm_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("TFC_SHM_DIR", str(tmp_path))
    return tmp_path


def test_client_streams_through_small_rings_and_hands_over_the_session(shm_dir):
    lib = client_lib()
    total = 1 << 20                                    # rings of 765 KiB / 255 KiB: everything below wraps
    w = FakeWorker(str(shm_dir / "tf_shm"), total, sessions=2)
    w.start()
    rng = np.random.default_rng(3)
    for session in (1, 2):
        c = C.c_void_p()
        assert lib.tfc_connect(b"shmem+tf_shm+1+1", C.byref(c)) == 0
        a, b = C.c_uint32(), C.c_uint32()
        n = 3_000_017                                  # 4x the upstream ring, 12x the downstream ring
        assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_malloc(c, 5000, C.byref(b)) == 0
        src = rng.integers(0, 256, n, dtype=np.uint8)
        assert lib.tfc_memcpy_h2d(c, a, 0, src.ctypes.data, n) == 0
        want_b = np.zeros(5000, dtype=np.uint8)
        for i in range(300):                           # many small frames: headers land on every 16-byte phase of the wrap
            piece = rng.integers(0, 256, 1 + i % 47, dtype=np.uint8)
            off = (i * 13) % 4900
            assert lib.tfc_memcpy_h2d(c, b, off, piece.ctypes.data, len(piece)) == 0
            want_b[off:off + len(piece)] = piece
        assert lib.tfc_memset(c, a, 10, 0x5A, 1000) == 0
        src[10:1010] = 0x5A
        got = np.empty(n, dtype=np.uint8)
        assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0 and np.array_equal(got, src)
        gb = np.empty(5000, dtype=np.uint8)
        assert lib.tfc_memcpy_d2h(c, gb.ctypes.data, b, 0, 5000) == 0 and np.array_equal(gb, want_b)
        assert lib.tfc_memcpy_d2h(c, gb.ctypes.data, 77, 0, 16) == 2      # RESP_ERROR comes back as the call's result
        assert lib.tfc_sync(c) == 0
        lib.tfc_close(c)
        deadline = time.time() + 5
        while w.h.session != session + 1 and time.time() < deadline:
            time.sleep(0.001)
        assert w.h.session == session + 1 and w.h.client_pid == 0 and w.h.worker_closed == session
    w.join(timeout=10)
    assert not w.is_alive()
    assert w.h.c2w_head == w.h.c2w_tail == w.bytes_in > 2 * 3_000_000 and w.h.w2c_head == w.h.w2c_tail
    assert w.max_inflight <= w.h.c2w_size              # the producer never overran the consumer


def test_connect_waits_for_a_ready_header_and_gives_up(shm_dir, monkeypatch):
    lib = client_lib()
    monkeypatch.setenv("TFC_CONNECT_TIMEOUT_MS", "300")
    c = C.c_void_p()
    t0 = time.time()
    assert lib.tfc_connect(b"shmem+nobody+1+1", C.byref(c)) == 5 and 0.25 < time.time() - t0 < 5     # no file
    (shm_dir / "touched").write_bytes(b"")                                                               # `touch`ed by the operator, worker not up
    assert lib.tfc_connect(b"shmem+touched+1+1", C.byref(c)) == 5
    assert lib.tfc_connect(b"shmem+../evil+1+1", C.byref(c)) == 1
    assert lib.tfc_connect(b"shmem+touched+1+2", C.byref(c)) == 3      # initVersion 2: a ring layout this client does not speak
    # the worker shows up while the client is waiting
    monkeypatch.setenv("TFC_CONNECT_TIMEOUT_MS", "5000")
    res = []
    th = threading.Thread(target=lambda: res.append(lib.tfc_connect(b"shmem+late+1+1", C.byref(c))), daemon=True)
    th.start()
    time.sleep(0.2)
    w = FakeWorker(str(shm_dir / "late"), 1 << 20)
    w.start()
    th.join(timeout=10)
    assert res == [0]
    # a second client cannot attach while the session is taken
    monkeypatch.setenv("TFC_CONNECT_TIMEOUT_MS", "200")
    c2 = C.c_void_p()
    assert lib.tfc_connect(b"shmem+late+1+1", C.byref(c2)) == 5
    lib.tfc_close(c)
    w.join(timeout=10)


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_large_copies_split_over_the_copy_pool(shm_dir, monkeypatch, threads):
    """Pieces of 2 MiB and more are copied by TFC_COPY_THREADS threads (into and out of the rings)."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import test_cpu_client_shm as T
lib = T.client_lib()
w = T.FakeWorker(os.path.join(os.environ["TFC_SHM_DIR"], "big"), 16 << 20)
w.start()
c = C.c_void_p()
assert lib.tfc_connect(b"shmem+big+16+1", C.byref(c)) == 0
n = 9_000_001
a = C.c_uint32()
assert lib.tfc_malloc(c, n, C.byref(a)) == 0
src = np.random.default_rng(8).integers(0, 256, n, dtype=np.uint8)
for off in (0, 3):                                   # 16-byte aligned and not
    assert lib.tfc_memcpy_h2d(c, a, off, src.ctypes.data, n - off) == 0
    got = np.empty(n, dtype=np.uint8)
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0
    want = src if off == 0 else np.concatenate([src[:3], src[:n - 3]])
    assert np.array_equal(got, want), off
assert lib.tfc_sync(c) == 0
lib.tfc_close(c)
w.join(timeout=10)
print("ok")
''' % (conftest.ROOT, conftest.ROOT)
    env = dict(os.environ, TFC_COPY_THREADS=threads, TFC_SHM_DIR=str(shm_dir))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]


def test_small_frames_never_overrun_a_stalled_consumer_in_a_later_session(shm_dir):
    """Cursors keep counting across sessions; the producer's cached view of the consumer's cursor must start from
    the real one.  The consumer sleeps while the client pushes three ring-fulls of small frames."""
    lib = client_lib()
    w = FakeWorker(str(shm_dir / "tf_shm"), 1 << 20, sessions=3, stall=0.3)
    w.start()
    rng = np.random.default_rng(17)
    for session in range(3):
        c = C.c_void_p()
        assert lib.tfc_connect(b"shmem+tf_shm+1+1", C.byref(c)) == 0
        a = C.c_uint32()
        n = 600 * 4096
        assert lib.tfc_malloc(c, n, C.byref(a)) == 0
        src = rng.integers(0, 256, n, dtype=np.uint8)
        for i in range(600):                      # 600 x (64 + 4096) bytes = 2.4 MiB through a 765 KiB ring
            assert lib.tfc_memcpy_h2d(c, a, i * 4096, src[i * 4096:].ctypes.data, 4096) == 0
        got = np.empty(n, dtype=np.uint8)
        assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0
        assert np.array_equal(got, src), session
        lib.tfc_close(c)
    w.join(timeout=20)
    assert not w.is_alive() and w.max_inflight <= w.h.c2w_size


def test_attached_client_holds_the_liveness_lock_until_it_leaves_or_dies(shm_dir):
    """PIDs mean nothing across the containers of a pod; an open-file-description lock on the ring file does
    (include/tfw_shm_ring.h).  The worker's probe (tools/ring_lock_probe.c uses the same inline functions) sees the
    lock while a client is attached, and sees it gone after tfc_close -- or after the client is killed."""
    import signal
    import subprocess
    probe = os.path.join(conftest.ROOT, "build", "mock", "ring_lock_probe")
    if not os.path.exists(probe):
        subprocess.run(["make", "-s", "build/mock/ring_lock_probe"], cwd=conftest.ROOT, check=True)
    path = str(shm_dir / "tf_shm")

    def seen():
        return subprocess.run([probe, path], capture_output=True, text=True, timeout=20).stdout.strip()

    lib = client_lib()
    w = FakeWorker(path, 1 << 20, sessions=1)
    w.start()
    assert seen() == "free"
    c = C.c_void_p()
    assert lib.tfc_connect(b"shmem+tf_shm+1+1", C.byref(c)) == 0
    assert w.h.client_lock_session == w.h.session == 1 and seen() == "held"
    assert lib.tfc_sync(c) == 0
    lib.tfc_close(c)
    assert seen() == "free"
    w.join(timeout=10)
    # a client that is killed: the kernel drops the lock with the process
    p = subprocess.Popen([probe, path, "hold"], stdout=subprocess.PIPE, text=True)
    assert p.stdout.readline().strip() == "holding" and seen() == "held"
    p.send_signal(signal.SIGKILL)
    p.wait(timeout=10)
    assert seen() == "free"


def test_handle_ids_are_recycled(shm_dir):
    """TFCS has 65 536 buffer ids; a client that allocates and frees for ever must not run out, and an id whose
    MALLOC the worker refused (quota) goes back into the pool as well."""
    lib = client_lib()
    lib.tfc_free.argtypes = [C.c_void_p, C.c_uint32]
    w = FakeWorker(str(shm_dir / "tf_shm"), 1 << 20, vram_quota=1 << 20)
    w.start()
    c = C.c_void_p()
    assert lib.tfc_connect(b"shmem+tf_shm+1+1", C.byref(c)) == 0
    h = C.c_uint32()
    seen = set()
    for i in range(3000):
        assert lib.tfc_malloc(c, 4096, C.byref(h)) == 0
        seen.add(h.value)
        assert lib.tfc_free(c, h) == 0
    assert lib.tfc_sync(c) == 0 and len(seen) <= 2                # the same id over and over
    keep = C.c_uint32()
    assert lib.tfc_malloc(c, 4096, C.byref(keep)) == 0
    big = C.c_uint32()
    assert lib.tfc_malloc(c, 2 << 20, C.byref(big)) == 0           # over the quota: refused by the worker ...
    assert lib.tfc_sync(c) == 4                                    # ... TFW_ERR_EXHAUSTED surfaces at the sync
    again = C.c_uint32()
    assert lib.tfc_malloc(c, 4096, C.byref(again)) == 0
    assert again.value == big.value and again.value != keep.value  # the refused id is reused, the live one is not
    data = np.arange(4096, dtype=np.uint8)
    got = np.empty(4096, dtype=np.uint8)
    assert lib.tfc_memcpy_h2d(c, again, 0, data.ctypes.data, 4096) == 0
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, again, 0, 4096) == 0 and np.array_equal(got, data)
    assert lib.tfc_sync(c) == 0
    lib.tfc_close(c)
    w.join(timeout=10)


@pytest.mark.parametrize("payload", [64, 4000, 300000])
def test_c_client_against_the_c_stand_in_at_full_speed(shm_dir, payload):
    """No interpreter on either side: tools/transport_lab.c (client library) against tools/null_worker.c (a consumer
    that parses headers, skips payloads, answers SYNC) through a 2 MiB file, so the cursors lap the rings
    hundreds of times at memory speed."""
    import subprocess
    mock = os.path.join(conftest.ROOT, "build", "mock")
    if not (os.path.exists(os.path.join(mock, "transport_lab")) and os.path.exists(os.path.join(mock, "null_worker"))):
        subprocess.run(["make", "-s", "build/mock/transport_lab", "build/mock/null_worker"], cwd=conftest.ROOT, check=True)
    path = str(shm_dir / "r")
    w = subprocess.Popen([os.path.join(mock, "null_worker"), path, "2", "1"], stdout=subprocess.PIPE, text=True)
    assert "serving" in w.stdout.readline()
    r = subprocess.run([os.path.join(mock, "transport_lab"), "shmem+r+2+1", "30000", str(payload)], capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, TFC_SHM_DIR=str(shm_dir)))
    assert r.returncode == 0, r.stderr
    out = __import__("json").loads(r.stdout)
    assert out["calls"] == 30000 and out["h2d_us_per_call"] > 0
    assert w.wait(timeout=20) == 0      # a header that did not parse would have ended it with "bad magic" (4)


class FakeTcpWorker(FakeWorker):
    """The same stand-in behind a TCP listener: the client's native+<ip>+<port>+<name> path on the CPU."""

    def __init__(self, vram_quota=None):
        threading.Thread.__init__(self, daemon=True)
        import socket
        self.vram_quota, self.launches, self.frames, self.bytes_in, self.stop = vram_quota, [], 0, 0, False
        self.path, self.payload_bytes_in = None, 0
        self.arenas, self.modules, self.functions, self.user_launches = {}, {}, {}, []
        self.ls = socket.socket()
        self.ls.bind(("127.0.0.1", 0))
        self.ls.listen(1)
        self.port = self.ls.getsockname()[1]

    def _send(self, data):
        self.conn.sendall(data)

    def run(self):
        self.conn, _ = self.ls.accept()
        bufs, stream = {}, bytearray()
        while True:
            b = self.conn.recv(1 << 20)
            if not b:
                break
            self.bytes_in += len(b)
            stream += b
            stream = self._execute(stream, bufs)
        self.conn.close()
        self.ls.close()


def test_tcp_transport_against_the_stand_in_worker():
    """native+<ip>+<port>+<name>-<rv> (tensorfusionconnection_controller.go:136-138) without a GPU: coalesced small
    frames, a payload sent in place, blocking reads, errors of blocking and of fire-and-forget calls."""
    lib = client_lib()
    lib.tfc_free.argtypes = [C.c_void_p, C.c_uint32]
    lib.tfc_last_error_code.argtypes = [C.c_void_p]
    w = FakeTcpWorker(vram_quota=64 << 20)
    w.start()
    c = C.c_void_p()
    assert lib.tfc_connect(f"native+127.0.0.1+{w.port}+tf-worker-abc-123".encode(), C.byref(c)) == 0
    rng = np.random.default_rng(2)
    n = 5_000_011
    a, b = C.c_uint32(), C.c_uint32()
    assert lib.tfc_malloc(c, n, C.byref(a)) == 0 and lib.tfc_malloc(c, 9000, C.byref(b)) == 0
    src = rng.integers(0, 256, n, dtype=np.uint8)
    assert lib.tfc_memcpy_h2d(c, a, 0, src.ctypes.data, n) == 0
    want_b = np.zeros(9000, dtype=np.uint8)
    for i in range(500):
        piece = rng.integers(0, 256, 1 + i % 33, dtype=np.uint8)
        assert lib.tfc_memcpy_h2d(c, b, i * 17, piece.ctypes.data, len(piece)) == 0
        want_b[i * 17:i * 17 + len(piece)] = piece
    assert lib.tfc_memset(c, a, 3, 0x11, 100) == 0
    src[3:103] = 0x11
    got = np.empty(n, dtype=np.uint8)
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, a, 0, n) == 0 and np.array_equal(got, src)
    gb = np.empty(9000, dtype=np.uint8)
    assert lib.tfc_memcpy_d2h(c, gb.ctypes.data, b, 0, 9000) == 0 and np.array_equal(gb, want_b)
    assert lib.tfc_memcpy_d2h(c, gb.ctypes.data, 4242, 0, 16) == 2          # unknown handle: the call's own result
    big = C.c_uint32()
    assert lib.tfc_malloc(c, 128 << 20, C.byref(big)) == 0                   # over the quota, fire-and-forget ...
    assert lib.tfc_sync(c) == 4 and lib.tfc_last_error_code(c) == 4          # ... reported by the next sync
    assert lib.tfc_sync(c) == 0
    assert lib.tfc_free(c, a) == 0 and lib.tfc_free(c, b) == 0 and lib.tfc_sync(c) == 0
    lib.tfc_close(c)
    w.join(timeout=10)
    assert not w.is_alive() and w.frames > 500
    # a worker on the loopback address is offered shared-memory rings first (TFCS_OP_UPGRADE_SHM); refused, the session
    # stays on the socket and the ring file the client had created is gone again
    assert w.upgrade_offers == 1


def test_connect_rejects_malformed_urls_and_dead_endpoints():
    lib = client_lib()
    c = C.c_void_p()
    for url in (b"", b"native+", b"native+127.0.0.1", b"nonsense", b"native+not-an-ip+80+x"):
        assert lib.tfc_connect(url, C.byref(c)) in (1, 5), url
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()                                                                # nobody listens there any more
    assert lib.tfc_connect(f"native+127.0.0.1+{port}+x".encode(), C.byref(c)) == 5
    assert lib.tfc_connect(None, C.byref(c)) == 1


def test_random_operation_sequences_through_tiny_rings(shm_dir):
    """Property test: any sequence of h2d / memset / d2h with random sizes (around the 256 KiB streaming
    threshold and the ring size), offsets and alignments leaves the buffers equal to a numpy model."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    lib = client_lib()
    counter = [0]
    size_st = st.one_of(st.integers(1, 300), st.integers(260_000, 266_000), st.integers(380_000, 400_000), st.integers(700_000, 1_200_000))
    op_st = st.tuples(st.sampled_from(["h2d", "memset", "d2h", "sync"]), size_st, st.integers(0, 1 << 30), st.integers(0, 255))

    @settings(max_examples=20, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.lists(op_st, min_size=1, max_size=25), st.integers(0, 2**32 - 1))
    def run(ops, seed):
        counter[0] += 1
        name = f"fz{counter[0]}"
        w = FakeWorker(str(shm_dir / name), 1 << 20)
        w.start()
        c = C.c_void_p()
        assert lib.tfc_connect(f"shmem+{name}+1+1".encode(), C.byref(c)) == 0
        n = 1_500_000
        h = C.c_uint32()
        assert lib.tfc_malloc(c, n, C.byref(h)) == 0
        model = np.zeros(n, dtype=np.uint8)
        rng = np.random.default_rng(seed)
        for kind, size, off_seed, val in ops:
            size = min(size, n)
            off = off_seed % (n - size + 1)
            if kind == "h2d":
                data = rng.integers(0, 256, size, dtype=np.uint8)
                assert lib.tfc_memcpy_h2d(c, h, off, data.ctypes.data, size) == 0
                model[off:off + size] = data
            elif kind == "memset":
                assert lib.tfc_memset(c, h, off, val, size) == 0
                model[off:off + size] = val
            elif kind == "d2h":
                got = np.empty(size, dtype=np.uint8)
                assert lib.tfc_memcpy_d2h(c, got.ctypes.data, h, off, size) == 0
                assert np.array_equal(got, model[off:off + size])
            else:
                assert lib.tfc_sync(c) == 0
        got = np.empty(n, dtype=np.uint8)
        assert lib.tfc_memcpy_d2h(c, got.ctypes.data, h, 0, n) == 0 and np.array_equal(got, model)
        lib.tfc_close(c)
        w.join(timeout=10)
        assert not w.is_alive()

    run()


def more_sigs(lib):
    lib.tfc_host_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.tfc_host_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.tfc_memcpy_d2h_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]
    lib.tfc_module_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
    lib.tfc_module_unload.argtypes = [C.c_void_p, C.c_uint32]
    lib.tfc_module_get_function.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.tfc_launch_user.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    return lib


def test_page_locked_client_memory_is_copied_by_reference(shm_dir):
    """tfc_host_alloc memory lives in an arena file next to the ring; copies from / to it carry no payload."""
    lib = more_sigs(client_lib())
    w = FakeWorker(str(shm_dir / "tf_shm"), 1 << 20)
    w.start()
    c = C.c_void_p()
    assert lib.tfc_connect(b"shmem+tf_shm+1+1", C.byref(c)) == 0
    n = 3 << 20                                             # three times the whole ring file
    p, q, small = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.tfc_host_alloc(c, n, C.byref(p)) == 0 and lib.tfc_host_alloc(c, n, C.byref(q)) == 0
    assert lib.tfc_host_alloc(c, 100, C.byref(small)) == 0
    assert os.path.exists(shm_dir / "tf_shm.a1")            # arena 1 = 64 MiB holds all three
    assert not os.path.exists(shm_dir / "tf_shm.a2")
    src = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n,))
    dst = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint8)), (n,))
    src[:] = np.random.default_rng(5).integers(0, 256, n, dtype=np.uint8)
    a = C.c_uint32()
    assert lib.tfc_malloc(c, n, C.byref(a)) == 0
    before = w.payload_bytes_in
    assert lib.tfc_memcpy_h2d(c, a, 0, p, n) == 0                              # by reference
    assert lib.tfc_memcpy_h2d(c, a, 5, C.c_void_p(p.value + 1000), 777) == 0   # interior pointer, odd alignment
    want = src.copy()
    want[5:5 + 777] = src[1000:1777]
    assert lib.tfc_memcpy_d2h(c, q, a, 0, n) == 0 and np.array_equal(dst, want)   # synchronous form: acknowledged
    dst[:] = 0
    assert lib.tfc_memcpy_d2h_async(c, q, a, 0, n) == 0 and lib.tfc_sync(c) == 0 and np.array_equal(dst, want)
    assert w.payload_bytes_in == before                                           # no payload crossed the ring
    page = np.empty(n, dtype=np.uint8)
    assert lib.tfc_memcpy_d2h_async(c, page.ctypes.data, a, 0, n) == 1            # pageable memory has no asynchronous form
    assert lib.tfc_memcpy_d2h(c, page.ctypes.data, a, 0, n) == 0 and np.array_equal(page, want)
    # a copy that runs past the buffer is the worker's error, reported by the next sync
    assert lib.tfc_memcpy_h2d(c, a, n - 10, p, 100) == 0 and lib.tfc_sync(c) == 1
    for ptr in (p, small):
        assert lib.tfc_host_free(c, ptr) == 0
    assert os.path.exists(shm_dir / "tf_shm.a1")            # q still lives in it
    assert lib.tfc_host_free(c, q) == 0 and not os.path.exists(shm_dir / "tf_shm.a1") and 1 not in w.arenas
    assert lib.tfc_host_free(c, q) == 2
    big = C.c_void_p()
    assert lib.tfc_host_alloc(c, 65 << 20, C.byref(big)) == 0 and os.path.getsize(shm_dir / "tf_shm.a1") == 65 << 20
    lib.tfc_close(c)                                         # closing gives the pages back
    assert not os.path.exists(shm_dir / "tf_shm.a1")
    w.join(timeout=10)


def test_user_modules_and_launches_travel_as_frames(shm_dir):
    lib = more_sigs(client_lib())
    w = FakeWorker(str(shm_dir / "tf_shm"), 1 << 20)
    w.start()
    c = C.c_void_p()
    assert lib.tfc_connect(b"shmem+tf_shm+1+1", C.byref(c)) == 0
    image = bytes(np.random.default_rng(2).integers(0, 256, 2_500_001, dtype=np.uint8))   # larger than the ring, odd length
    m, bad = C.c_uint32(), C.c_uint32()
    assert lib.tfc_module_load(c, image, len(image), C.byref(m)) == 0 and w.modules[m.value] == image
    assert lib.tfc_module_load(c, b"this is not a code image", 24, C.byref(bad)) == 1     # refused by the worker, reported by the call
    f, nparams, pbytes = C.c_uint32(), C.c_uint32(), C.c_uint32()
    offs, sizes = (C.c_uint32 * 8)(), (C.c_uint32 * 8)()
    assert lib.tfc_module_get_function(c, m, b"saxpy_u32", C.byref(f), C.byref(nparams), offs, sizes, 8, C.byref(pbytes)) == 0
    assert nparams.value == 4 and list(offs[:4]) == [0, 8, 16, 20] and list(sizes[:4]) == [8, 8, 4, 4] and pbytes.value == 24
    g = C.c_uint32()
    assert lib.tfc_module_get_function(c, m, b"nope", C.byref(g), None, None, None, 0, None) == 2
    n = 10_001
    x, y = C.c_uint32(), C.c_uint32()
    assert lib.tfc_malloc(c, 4 * n, C.byref(x)) == 0 and lib.tfc_malloc(c, 4 * n + 64, C.byref(y)) == 0
    hx = np.arange(n, dtype=np.uint32) * np.uint32(2654435761)
    hy = np.arange(n, dtype=np.uint32) ^ np.uint32(0xABCDEF)
    assert lib.tfc_memcpy_h2d(c, x, 0, hx.ctypes.data, 4 * n) == 0 and lib.tfc_memcpy_h2d(c, y, 64, hy.ctypes.data, 4 * n) == 0
    import struct
    params = struct.pack("<QQII", wire.tagged_ptr(x.value), wire.tagged_ptr(y.value, 64), 7, n)
    grid, block = (C.c_uint32 * 3)(40, 1, 1), (C.c_uint32 * 3)(256, 1, 1)
    assert lib.tfc_launch_user(c, f, grid, block, 0, params, len(params), 320) == 0
    got = np.empty(n, dtype=np.uint32)
    assert lib.tfc_memcpy_d2h(c, got.ctypes.data, y, 64, 4 * n) == 0
    assert np.array_equal(got, np.uint32(7) * hx + hy)
    assert w.user_launches == [(b"saxpy_u32", (40, 1, 1), (256, 1, 1), 0, 24, 320)]
    assert lib.tfc_module_unload(c, m) == 0 and lib.tfc_sync(c) == 0 and not w.modules
    lib.tfc_close(c)
    w.join(timeout=10)
