"""TFCS wire format helpers (include/tfw_wire.h) for tests and tools."""
import ctypes as C
import struct

MAGIC = 0x53434654
VERSION = 1
HDR_BYTES = 64

OP_NOP, OP_MALLOC, OP_FREE, OP_H2D, OP_D2H, OP_D2D, OP_MEMSET, OP_LAUNCH, OP_SYNC = range(9)
OP_RESP_D2H, OP_RESP_SYNC, OP_RESP_ERROR = 0x84, 0x88, 0xFF
K_NOOP, K_SPIN, K_ADD_U8, K_XOR_IDX = range(4)

_HDR = struct.Struct("<IHHIIIIQQQIIII")
assert _HDR.size == HDR_BYTES


def pad16(n):
    return (n + 15) & ~15


def frame(op, call_id=0, h0=0, h1=0, off0=0, off1=0, length=0, arg0=0, arg1=0, arg2=0, arg3=0, payload=b""):
    hdr = _HDR.pack(MAGIC, VERSION, op, call_id, 0, h0, h1, off0, off1, length, arg0, arg1, arg2, arg3)
    if op in (OP_H2D, OP_RESP_D2H):
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
        if h["opcode"] in (OP_H2D, OP_RESP_D2H):
            pay = bytes(buf[pos:pos + h["length"]])
            pos += pad16(h["length"])
        yield h, pay
