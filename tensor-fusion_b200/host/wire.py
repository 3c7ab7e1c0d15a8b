"""TFCS wire format helpers (include/tfw_wire.h) for tests and tools."""
import ctypes as C
import struct

MAGIC = 0x53434654
VERSION = 1
HDR_BYTES = 64

OP_NOP, OP_MALLOC, OP_FREE, OP_H2D, OP_D2H, OP_D2D, OP_MEMSET, OP_LAUNCH, OP_SYNC = range(9)
OP_HOST_REGISTER, OP_HOST_UNREGISTER, OP_H2D_REF, OP_D2H_REF = 9, 10, 11, 12
OP_MODULE_LOAD, OP_MODULE_UNLOAD, OP_MODULE_GET_FUNCTION, OP_LAUNCH_USER, OP_UPGRADE_SHM = 13, 14, 15, 16, 17
OP_RESP_D2H, OP_RESP_SYNC, OP_RESP_ACK, OP_RESP_FUNCTION, OP_RESP_ERROR = 0x84, 0x88, 0x8C, 0x8F, 0xFF
PAYLOAD_OPS = (OP_H2D, OP_RESP_D2H, OP_MODULE_LOAD, OP_MODULE_GET_FUNCTION, OP_LAUNCH_USER, OP_RESP_FUNCTION, OP_UPGRADE_SHM)
F_ACK = 1
PTR_TAG = 1 << 62
K_NOOP, K_SPIN, K_ADD_U8, K_XOR_IDX = range(4)


def tagged_ptr(handle, offset=0):
    """The device pointer a remote-mode application holds for (handle, offset) (TFCS_PTR_TAG, include/tfw_wire.h)."""
    return PTR_TAG | (handle << 40) | offset


def launch_payload(grid, block, params, shared_bytes=0):
    """tfcs_launch_params + parameter block of a LAUNCH_USER frame."""
    g = tuple(grid) + (1,) * (3 - len(grid))
    b = tuple(block) + (1,) * (3 - len(block))
    return struct.pack("<8I", *g, *b, shared_bytes, len(params)) + bytes(params)

_HDR = struct.Struct("<IHHIIIIQQQIIII")
assert _HDR.size == HDR_BYTES


def pad16(n):
    return (n + 15) & ~15


def frame(op, call_id=0, h0=0, h1=0, off0=0, off1=0, length=0, arg0=0, arg1=0, arg2=0, arg3=0, payload=b"", flags=0):
    hdr = _HDR.pack(MAGIC, VERSION, op, call_id, flags, h0, h1, off0, off1, length, arg0, arg1, arg2, arg3)
    if op in PAYLOAD_OPS:
        assert len(payload) == length
        return hdr + bytes(payload) + b"\0" * (pad16(length) - length)
    return hdr


class Builder:
    """Tiny trace recorder: b.malloc(1, 4096); b.h2d(1, 0, data); ... ; bytes(b)"""

    def __init__(self):
        self.parts = []
        self.call_id = 0

    def _add(self, op, **kw):
        self.parts.append(frame(op, call_id=self.call_id, **kw))
        self.call_id += 1
        return self

    def malloc(self, h, size): return self._add(OP_MALLOC, h0=h, length=size)
    def free(self, h): return self._add(OP_FREE, h0=h)
    def h2d(self, h, off, data): return self._add(OP_H2D, h0=h, off0=off, length=len(data), payload=data)
    def d2h(self, h, off, n): return self._add(OP_D2H, h0=h, off0=off, length=n)
    def d2d(self, dst, doff, src, soff, n): return self._add(OP_D2D, h0=dst, off0=doff, h1=src, off1=soff, length=n)
    def memset(self, h, off, n, value): return self._add(OP_MEMSET, h0=h, off0=off, length=n, arg0=value)
    def launch(self, kernel, grid=1, block=32, h=0, off=0, n=0, scalar=0, cost=0):
        return self._add(OP_LAUNCH, arg0=kernel, arg1=grid, arg2=block, arg3=cost, h0=h, off0=off, length=n, off1=scalar)
    def sync(self): return self._add(OP_SYNC)
    def host_register(self, arena, size): return self._add(OP_HOST_REGISTER, h0=arena, length=size)
    def host_unregister(self, arena): return self._add(OP_HOST_UNREGISTER, h0=arena)
    def h2d_ref(self, h, off, arena, aoff, n): return self._add(OP_H2D_REF, h0=h, off0=off, h1=arena, off1=aoff, length=n)
    def d2h_ref(self, h, off, arena, aoff, n, ack=False):
        return self._add(OP_D2H_REF, h0=h, off0=off, h1=arena, off1=aoff, length=n, flags=F_ACK if ack else 0)
    def module_load(self, module, image): return self._add(OP_MODULE_LOAD, h0=module, length=len(image), payload=image)
    def module_unload(self, module): return self._add(OP_MODULE_UNLOAD, h0=module)
    def get_function(self, module, function, name):
        return self._add(OP_MODULE_GET_FUNCTION, h0=module, h1=function, length=len(name), payload=name)
    def launch_user(self, function, grid, block, params, shared_bytes=0, cost=0):
        pay = launch_payload(grid, block, params, shared_bytes)
        return self._add(OP_LAUNCH_USER, h1=function, arg3=cost, length=len(pay), payload=pay)
    def raw(self, b):
        self.parts.append(bytes(b))
        return self

    def __bytes__(self):
        return b"".join(self.parts)


_NAMES = ("magic", "version", "opcode", "call_id", "flags", "h0", "h1", "off0", "off1", "length", "arg0", "arg1", "arg2", "arg3")


def unpack_header(b):
    """The 64-byte frame header as a dict."""
    return dict(zip(_NAMES, _HDR.unpack_from(b, 0)))


def parse_frames(buf):
    """Yield (hdr_dict, payload_bytes) for every frame in ``buf``."""
    buf = memoryview(buf)
    pos = 0
    names = ("magic", "version", "opcode", "call_id", "flags", "h0", "h1", "off0", "off1", "length",
             "arg0", "arg1", "arg2", "arg3")
    while pos + HDR_BYTES <= len(buf):
        vals = _HDR.unpack_from(buf, pos)
        h = dict(zip(names, vals))
        pos += HDR_BYTES
        pay = b""
        if h["opcode"] in PAYLOAD_OPS:
            pay = bytes(buf[pos:pos + h["length"]])
            pos += pad16(h["length"])
        yield h, pay
